#!/bin/bash
# round 2, visit k: warp-uniform tight 1-D step (R = this tree, F2 = before the per-level changes), per-COT diagnosis of the LUT curve
mkdir -p gpurun_out
bash tools/gpu_variants.sh C1,C1H,C5,C5S F2 R 2>&1 | tee gpurun_out/ab_r02_k.txt
timeout 300 python tools/diag_cot.py 2>&1 | tail -40 | tee gpurun_out/diag_cot_r02_k.txt
ER3T_B200_LIB=$PWD/tools/variants/libF2.so timeout 300 python tools/diag_cot.py 2>&1 | tail -40 | tee gpurun_out/diag_cot_r02_k_F2.txt
