set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q --tb=short -p no:cacheprovider > gpurun_out/t_chain.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/t_chain.log | cut -c1-300
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/chain_bench.txt | cut -c1-330
