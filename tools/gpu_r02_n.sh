#!/bin/bash
# round 2, visit n: photons inside 1-D layers in their own queue and phase (per-level kernels): U = this tree, T = commit 12f2726
mkdir -p gpurun_out
bash tools/gpu_variants.sh C1,C1H,C5,C5S,C2 T U 2>&1 | tee gpurun_out/ab_r02_n.txt
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=6 > gpurun_out/pytest_r02_n.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_n.log
grep -E "1-D:|hom-3D:|ref vs COT|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_n.log | head -30
