#!/bin/bash
# A/B of prebuilt library variants (tools/variants/lib<NAME>.so, selected with $ER3T_B200_LIB) on one box.
#   bash tools/gpu_variants.sh bench O A B ...      # C2 bench, resident leg
#   bash tools/gpu_variants.sh C3,C4 O A B ...      # tools/bench_configs.py on the named configs
MODE=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  export ER3T_B200_LIB=$PWD/tools/variants/lib$v.so
  if [ "$MODE" == "bench" ]; then
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-accuracy --e2e-steps 1 > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
    python -c "import json; d=json.load(open('gpurun_out/variant_$v.json')); print('$v', round(d['value']/1e6,1), 'M photons/s resident', round(d['e2e']['value']/1e6,1), 'e2e')"
  else
    python tools/bench_configs.py --configs $MODE --reps 2 --out gpurun_out/variant_cfg_$v.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v', d['config'], round(d['mphotons_per_s'], 1), 'M photons/s')"
  fi
done
