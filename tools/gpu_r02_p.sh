#!/bin/bash
# round 2, visit p: tally context hoisted out of the per-level flight loop (V) against the tree of commit 12f2726 (T)
mkdir -p gpurun_out
bash tools/gpu_variants.sh C1,C1H,C5,C5S T V 2>&1 | tee gpurun_out/ab_r02_p.txt
ER3T_B200_LIB=$PWD/tools/variants/libV.so timeout 600 python -m pytest tests -m gpu -x -q -k "flux or c1 or c5 or heating or tallies or plane_parallel or 1e9 or ipa or shards" > gpurun_out/pytest_r02_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_p.log
tail -3 gpurun_out/pytest_r02_p.log
