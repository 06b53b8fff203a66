#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, DRAM traffic at the bench size, one full capture of the transport kernel.
# usage (under gpurun): bash tools/gpu_round.sh TAG [notest]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --photons 1e7 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# DRAM traffic of ONE transport launch at the bench size (3e8 photons): roofline.traffic
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:transport_kernel --launch-skip 2 -c 1 --csv \
    --log-file gpurun_out/traffic_$TAG.csv python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/traffic_run_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:transport_kernel --launch-skip 1 -c 1 -f -o gpurun_out/transport_$TAG \
    python bench.py --steps 1 --warmup 1 --photons 3e6 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
