set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tiled_gpu.py -x -q --tb=short -p no:cacheprovider -k "umma or tiled or dgrad" > gpurun_out/t_wide.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t_wide.log | cut -c1-300
timeout 600 python scripts/bench_infer4k.py --steps 5 --warmup 2 > gpurun_out/infer4k_1gpu.json 2> gpurun_out/infer4k_1gpu.err; echo "rc=$?"; tail -n 3 gpurun_out/infer4k_1gpu.err; cat gpurun_out/infer4k_1gpu.json
SRB200_NO_WIDE=1 timeout 600 python scripts/bench_infer4k.py --steps 5 --warmup 2 2>/dev/null | cut -c1-300
