set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pack" --tb=short -p no:cacheprovider 2>&1 | tail -n 4 | cut -c1-300
timeout 60 python scripts/pack_bench.py 2>&1 | tail -n 1
SRB200_PACK_STAGED=0 timeout 60 python scripts/pack_bench.py 2>&1 | tail -n 1
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_rcan.err | tee gpurun_out/bench_rcan_staged.json | cut -c1-230
