#!/bin/bash
# the round's closing measurement on the committed tree: full GPU suite, every config and named variant, the bench line
# usage (under gpurun): bash tools/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -10 gpurun_out/pytest_$TAG.log
timeout 600 python tools/bench_configs.py --reps 2 --configs C1,C1H,C2,C2R,C3,C3V1,C3V9,C4,C5,C5S --out gpurun_out/configs_$TAG.json > gpurun_out/configs_$TAG.log 2>&1
python - <<PY
import json
try:
    for r in json.load(open('gpurun_out/configs_$TAG.json')):
        print(r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'B_alg/photon %.0f' % r['bytes_alg_per_photon'], 'frac %.3f' % r['roofline_frac_hbm'], 'balance %.1e' % r['max_abs_balance'])
except Exception as e:
    print('configs FAILED', e)
PY
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$TAG.json'))
    print('value %.1f M/s  e2e %.1f M/s (%.1f ms/step)  issue frac %.3f  hbm frac %.4f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms_per_step'], d['roofline']['issue']['frac'], d['roofline']['frac']))
except Exception as e:
    print('bench FAILED', e)
PY
