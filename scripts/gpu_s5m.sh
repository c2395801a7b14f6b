set -x
timeout 60 python scripts/pack_bench.py 2>&1 | tail -n 1
timeout 100 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -x -q -m gpu -k "pack or trainstep" --tb=short -p no:cacheprovider 2>&1 | tail -n 3 | cut -c1-300
