#!/bin/bash
# round 2, visit q: the round's measurement set on the tree of commit 2ddc60a -- GPU suite, bench line, ncu launch list, DRAM
# traffic at the bench size, full captures (config 2, 3, 5), compute-sanitizer
bash tools/gpu_round.sh r02_q
bash tools/gpu_ncu_cfg.sh C3 0.03 r02_q
bash tools/gpu_ncu_cfg.sh C5 0.3 r02_q
bash tools/gpu_memcheck.sh r02_q
ls -la gpurun_out | head -60
