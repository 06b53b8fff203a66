#!/bin/bash
# round 2, visit i: tight 1-D layer loop in the per-level kernels (P = this tree; F2 = the tree before it), lean group look-up in
# the UZ kernels (L = P + -DRT_UZ_LEAN); full GPU suite on P
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/pytest_r02_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_i.log
grep -E "1-D:|hom-3D:|ref vs COT|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_i.log | head -30
bash tools/gpu_variants.sh C1,C1H,C5,C5S F2 P 2>&1 | tee gpurun_out/ab_r02_i.txt
bash tools/gpu_variants.sh bench P L 2>&1 | tee -a gpurun_out/ab_r02_i.txt
ER3T_B200_LIB=$PWD/tools/variants/libL.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r02_i_L.log 2>&1; echo "pytest L exit $?" >> gpurun_out/pytest_r02_i_L.log
tail -3 gpurun_out/pytest_r02_i_L.log
