#!/bin/bash
# N GPUs of one box (gpurun --gpus N; usage: bash tools/gpu_2gpu.sh [N=2]): the 2-rank NCCL parity test and the bench line
# with its strong-scaling and config-4 sweep legs
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_r02_${N}gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r02_${N}gpu.log
tail -3 gpurun_out/pytest_r02_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_${N}gpu.json'))
    print('weak %.1f M/s, e2e %.1f M/s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6))
    print('strong', json.dumps(d['strong']))
    print('c4_sweep', json.dumps(d['c4_sweep']))
except Exception as e:
    print('bench FAILED', e)
PY
