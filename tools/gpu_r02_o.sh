#!/bin/bash
# round 2, visit o: re-tune the run-time launch / grid parameters on the round-2 kernel (config 2, 2e7 photons per point)
mkdir -p gpurun_out
timeout 400 python tools/sweep_pool.py -:96,320,2,16,12 -:96,320,2,8,12 -:96,320,2,12,12 -:96,320,2,24,12 -:96,320,2,32,12 -:96,320,2,16,8 -:96,320,2,16,16 -:96,320,2,16,20 \
    -:96,320,2,24,16 -:128,320,2,16,12 -:128,320,2,24,16 -:64,320,2,16,12 -:96,256,2,16,12 -:96,288,2,16,12 2>&1 | grep -v Warning | tee gpurun_out/sweep_pool_r02_o.txt
timeout 400 python tools/sweep_sv.py 2,2,3 2,2,2 2,2,4 3,3,4 2,2,3,8,8,1 2,2,3,2,2,1 3,3,3 1,1,2 4,4,6 2,2,5 2>&1 | grep -v Warning | tee gpurun_out/sweep_sv_r02_o.txt
