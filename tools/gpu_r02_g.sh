#!/bin/bash
# round 2, visit g: UZ-templated kernel (F = this tree) and the split-chain local-estimate march (S = F + -DRT_LE_SPLIT)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "not 1e9" --durations=8 > gpurun_out/pytest_r02_g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_g.log
grep -E "ref vs COT|passed|failed|Error|assert|^[0-9.]+s " gpurun_out/pytest_r02_g.log | head -30
ER3T_B200_LIB=$PWD/tools/variants/libS.so timeout 300 python -m pytest tests -m gpu -x -q -k "oblique or camera or c3 or cyclic or sensor" > gpurun_out/pytest_r02_g_S.log 2>&1; echo "pytest S exit $?" >> gpurun_out/pytest_r02_g_S.log
tail -3 gpurun_out/pytest_r02_g_S.log
bash tools/gpu_variants.sh bench F S 2>&1 | tee gpurun_out/ab_r02_g.txt
bash tools/gpu_variants.sh C3,C3V9,C4,C5,C1 F S 2>&1 | tee -a gpurun_out/ab_r02_g.txt
