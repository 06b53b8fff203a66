mkdir -p gpurun_out
bash tools/gpu_variants.sh C1,C1H,C5,C5S W A2 A3 2>&1 | tee gpurun_out/ab_r02_u.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "flux or c1 or c5 or heating or tallies or plane_parallel or 1e9 or ipa or shards or lut or cot or variants" 2>&1 | tail -3
