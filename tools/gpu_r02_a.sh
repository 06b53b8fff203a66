#!/bin/bash
# round 2, visit a: first run of the role-specialised kernel (v9): parity tests, A/B against v8, launch-shape variants
mkdir -p gpurun_out
export B200RT_KERNEL=9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_a.log
tail -5 gpurun_out/pytest_r02_a.log
one() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_r02_a_$name.json 2> gpurun_out/ab_r02_a_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_r02_a_$name.json'))
    print('$name', round(d['value'] / 1e6, 1), 'M photons/s resident', round(d['e2e']['value'] / 1e6, 1), 'e2e')
except Exception as e:
    print('$name', 'FAILED', e)
PY
}
one v8 B200RT_KERNEL=8
one v9 B200RT_KERNEL=9
one v9_p1024 B200RT_KERNEL=9 B200RT_V9_POOL=1024
one v9_p2048 B200RT_KERNEL=9 B200RT_V9_POOL=2048
for v in B C D E; do one v9_$v B200RT_KERNEL=9 ER3T_B200_LIB=$PWD/tools/variants/lib$v.so; done
B200RT_KERNEL=9 timeout 600 python tools/bench_configs.py --reps 2 --out gpurun_out/configs_r02_a_v9.json > gpurun_out/configs_r02_a_v9.log 2>&1
B200RT_KERNEL=8 timeout 600 python tools/bench_configs.py --reps 2 --out gpurun_out/configs_r02_a_v8.json > gpurun_out/configs_r02_a_v8.log 2>&1
python - <<PY
import json
for k in ('v8', 'v9'):
    try:
        for r in json.load(open('gpurun_out/configs_r02_a_%s.json' % k)):
            print(k, r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'balance %.1e' % r['max_abs_balance'])
    except Exception as e:
        print(k, 'FAILED', e)
PY
