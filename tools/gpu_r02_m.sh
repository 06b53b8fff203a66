#!/bin/bash
# round 2, visit m: tree with the tight 1-D step compiled into the plane-parallel / few-column per-level kernels only (T);
# full GPU suite, every config and variant at its named shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=6 > gpurun_out/pytest_r02_m.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_m.log
grep -E "1-D:|hom-3D:|ref vs COT|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_m.log | head -30
timeout 600 python tools/bench_configs.py --reps 2 --configs C1,C1H,C2,C2R,C3,C3V1,C3V9,C4,C5,C5S --out gpurun_out/configs_r02_m.json > gpurun_out/configs_r02_m.log 2>&1
python - <<PY
import json
try:
    for r in json.load(open('gpurun_out/configs_r02_m.json')):
        print(r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'B_alg/photon %.0f' % r['bytes_alg_per_photon'], 'frac %.3f' % r['roofline_frac_hbm'], 'balance %.1e' % r['max_abs_balance'])
except Exception as e:
    print('configs FAILED', e)
PY
