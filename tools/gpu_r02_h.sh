#!/bin/bash
# round 2, visit h: code-size experiments on the config-2 kernel (F = tree, N1 = without the oblique local-estimate march,
# N3 = without the __syncwarp after queue pops) + the repaired reflectance-vs-COT test
mkdir -p gpurun_out
bash tools/gpu_variants.sh bench F N1 N3 2>&1 | tee gpurun_out/ab_r02_h.txt
timeout 300 python -m pytest tests/test_gpu_deterministic.py -m gpu -x -q -s -k "cot" > gpurun_out/pytest_r02_h.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_h.log
grep -E "ref vs COT|passed|failed|Error|assert" gpurun_out/pytest_r02_h.log | head
