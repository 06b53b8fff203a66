mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01_g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r01_g.log
tail -5 gpurun_out/pytest_r01_g.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_g.json 2> gpurun_out/bench_r01_g.err; cat gpurun_out/bench_r01_g.json
python bench.py --steps 3 --warmup 3 --empty-runs -1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_r01_g_noruns.json 2>&1; cat gpurun_out/bench_r01_g_noruns.json
timeout 600 python tools/bench_configs.py --out gpurun_out/configs_r01_g.json > gpurun_out/configs_r01_g.log 2>&1; tail -8 gpurun_out/configs_r01_g.log
