set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_kernels_gpu.py -x -q --tb=short -p no:cacheprovider > gpurun_out/t_chain.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/t_chain.log | cut -c1-300
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; echo "rc=$?"; grep -v "^trace" gpurun_out/chain_bench.txt | tail -n 10 | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-300 gpurun_out/bench_rcan.json
