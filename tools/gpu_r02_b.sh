#!/bin/bash
# round 2, visit b: role-specialised kernel after the fixes (uniform decisions, balanced setmaxnreg)
mkdir -p gpurun_out
for v in N A; do
  if [ $v == A ]; then unset ER3T_B200_LIB; else export ER3T_B200_LIB=$PWD/tools/variants/lib$v.so; fi
  timeout 120 python tools/dbg_v9.py > gpurun_out/dbg_r02_b_$v.log 2>&1; rc=$?
  echo "dbg $v rc=$rc"; tail -4 gpurun_out/dbg_r02_b_$v.log
  if [ $rc != 0 ]; then
    timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python tools/dbg_v9.py 20000 > gpurun_out/memcheck_r02_b_$v.log 2>&1
    grep -E "Invalid|at .*transport|b200rt.cu|transport_v9.cuh|ERROR SUMMARY" gpurun_out/memcheck_r02_b_$v.log | head -30
  fi
done
unset ER3T_B200_LIB
if ! grep -q "DBG_V9 OK" gpurun_out/dbg_r02_b_A.log; then echo "v9 broken: stop"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_b.log
tail -5 gpurun_out/pytest_r02_b.log
one() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_r02_b_$name.json 2> gpurun_out/ab_r02_b_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_r02_b_$name.json'))
    print('$name', round(d['value'] / 1e6, 1), 'M photons/s resident', round(d['e2e']['value'] / 1e6, 1), 'e2e')
except Exception as e:
    print('$name', 'FAILED', e)
PY
}
one v8 B200RT_KERNEL=8
one v9 B200RT_KERNEL=9
for v in B C D E F; do one v9_$v B200RT_KERNEL=9 ER3T_B200_LIB=$PWD/tools/variants/lib$v.so; done
B200RT_KERNEL=9 timeout 500 python tools/bench_configs.py --reps 2 --out gpurun_out/configs_r02_b_v9.json > gpurun_out/configs_r02_b_v9.log 2>&1
python - <<PY
import json
for k in ('v9',):
    try:
        for r in json.load(open('gpurun_out/configs_r02_b_%s.json' % k)):
            print(k, r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'balance %.1e' % r['max_abs_balance'])
    except Exception as e:
        print(k, 'FAILED', e)
PY
