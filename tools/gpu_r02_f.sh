#!/bin/bash
# round 2, visit f: did the tally refactor slow the config-2 kernel?  A/B of prebuilt libraries on one box:
#   O = kernel of commit ebb7d38 (before the refactor), A = this tree, B = A without the __syncwarp after queue pops,
#   C = 224 threads x 3 blocks (21 warps / SM), D = 160 x 4, E = A without the unequal-layer path in the flight loop
export B200RT_KERNEL=8
bash tools/gpu_variants.sh bench O A B C D E 2>&1 | tee gpurun_out/ab_r02_f.txt
bash tools/gpu_variants.sh C3,C4,C5,C1 O A 2>&1 | tee -a gpurun_out/ab_r02_f.txt
