#!/bin/bash
# A/B of prebuilt library variants on the C2 bench (resident leg only): bash tools/gpu_variants.sh O A B ...
mkdir -p gpurun_out
for v in "$@"; do
  ER3T_B200_LIB=$PWD/tools/variants/lib$v.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/variant_$v.json'))
print('$v', round(d['value']/1e6,1), 'M photons/s', d['clocks'])
PY
done
