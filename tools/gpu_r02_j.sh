#!/bin/bash
# round 2, visit j: one tight 1-D layer step per loop pass in the per-level kernels (Q = this tree; F2 = the tree before the change)
mkdir -p gpurun_out
bash tools/gpu_variants.sh C1,C1H,C5,C5S F2 Q 2>&1 | tee gpurun_out/ab_r02_j.txt
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=6 > gpurun_out/pytest_r02_j.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_j.log
grep -E "1-D:|hom-3D:|ref vs COT|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_j.log | head -30
