#!/bin/bash
# full ncu capture of the transport kernel on one of the other configs: bash tools/gpu_ncu_cfg.sh C3 0.03 TAG
CFG=$1; PF=$2; TAG=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:transport_kernel --launch-skip 1 -c 1 -f -o gpurun_out/transport_${CFG}_$TAG \
    python tools/bench_configs.py --configs $CFG --reps 1 --photon-factor $PF --out gpurun_out/cfg_under_ncu.json > gpurun_out/ncu_${CFG}_$TAG.log 2>&1
tail -3 gpurun_out/ncu_${CFG}_$TAG.log
