#!/bin/bash
# round 2, visit d: full GPU suite (incl. the 1e9-photon tests against the deterministic solver), full bench line, configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_r02_d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_d.log
grep -E "1-D:|hom-3D:|passed|failed|Error|assert" gpurun_out/pytest_r02_d.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err; tail -3 gpurun_out/bench_r02_d.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_d.json'))
    print('value %.1f M/s  e2e %.1f M/s (%.0f ms/step, h2d %.0f MB)' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'] / 1e6))
    print('accuracy', json.dumps(d['accuracy']))
    print('issue', d['roofline']['issue'])
except Exception as e:
    print('bench FAILED', e)
PY
timeout 500 python tools/bench_configs.py --reps 2 --out gpurun_out/configs_r02_d.json > gpurun_out/configs_r02_d.log 2>&1
python - <<PY
import json
try:
    for r in json.load(open('gpurun_out/configs_r02_d.json')):
        print(r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'upload %.0f ms' % r['upload_ms'], 'balance %.1e' % r['max_abs_balance'])
except Exception as e:
    print('configs FAILED', e)
PY
