#!/bin/bash
# round 2, visit e: full GPU suite on the repaired tree (warp-aggregated private tallies, guide tables), bench line with the
# zero-copy e2e leg, end-to-end breakdown, all configs + named variants, private-tally A/B on the plane-parallel configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=12 > gpurun_out/pytest_r02_e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_e.log
grep -E "1-D:|hom-3D:|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_e.log | head -40
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_e.json 2> gpurun_out/bench_r02_e.err; tail -3 gpurun_out/bench_r02_e.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_e.json'))
    print('value %.1f M/s  e2e %.1f M/s (%.0f ms/step, h2d %.0f MB)' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'] / 1e6))
    print('accuracy', json.dumps(d['accuracy']))
except Exception as e:
    print('bench FAILED', e)
PY
PHOT=1e7 timeout 200 python tools/e2e_breakdown.py 2>&1 | tail -4 | tee gpurun_out/e2e_breakdown_r02_e.txt
timeout 600 python tools/bench_configs.py --reps 2 --configs C1,C1H,C2,C2R,C3,C3V1,C3V9,C4,C5,C5S --out gpurun_out/configs_r02_e.json > gpurun_out/configs_r02_e.log 2>&1
python - <<PY
import json
try:
    for r in json.load(open('gpurun_out/configs_r02_e.json')):
        print(r['config'], round(r['mphotons_per_s'], 1), 'M photons/s', 'upload %.0f ms' % r['upload_ms'], 'balance %.1e' % r['max_abs_balance'], 'tallies/photon %.1f' % r['per_photon']['n_tally'])
except Exception as e:
    print('configs FAILED', e)
PY
for m in -1 2; do
timeout 200 python tools/bench_configs.py --reps 2 --configs C1,C1H --smem-tally $m --out gpurun_out/configs_r02_e_tal$m.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('smem_tally $m', d['config'], round(d['mphotons_per_s'], 1), 'M photons/s')"
done
