#!/bin/bash
# round 2, visit c: v9 with the full-batch wait policy and lighter fences; ncu capture of v9 on config 2
mkdir -p gpurun_out
timeout 120 python tools/dbg_v9.py > gpurun_out/dbg_r02_c.log 2>&1; echo "dbg rc=$?"; tail -3 gpurun_out/dbg_r02_c.log
if ! grep -q "DBG_V9 OK" gpurun_out/dbg_r02_c.log; then echo "v9 broken: stop"; exit 1; fi
one() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_r02_c_$name.json 2> gpurun_out/ab_r02_c_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_r02_c_$name.json'))
    print('$name', round(d['value'] / 1e6, 1), 'M photons/s resident', round(d['e2e']['value'] / 1e6, 1), 'e2e')
except Exception as e:
    print('$name', 'FAILED', e)
PY
}
one v9 B200RT_KERNEL=9
for v in B E A0 A2 AW0 E0; do one v9_$v B200RT_KERNEL=9 ER3T_B200_LIB=$PWD/tools/variants/lib$v.so; done
B200RT_KERNEL=9 timeout 300 ncu --set full --clock-control none --import-source on -k regex:transport_v9 --launch-skip 1 -c 1 -f -o gpurun_out/transport_r02_c \
    python bench.py --steps 1 --warmup 1 --photons 3e6 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_r02_c.log 2>&1
ls -la gpurun_out/transport_r02_c.ncu-rep
