TAG=${1:-r01_i}; CFGS=${2:-C1,C2,C3,C4,C5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python tools/bench_configs.py --configs $CFGS --out gpurun_out/configs_$TAG.json > gpurun_out/configs_$TAG.log 2>&1; cut -c1-330 gpurun_out/configs_$TAG.log | tail -8
