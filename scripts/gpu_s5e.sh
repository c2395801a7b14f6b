set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 12 gpurun_out/smoke.log | cut -c1-400
timeout 600 python -m pytest tests/test_runner_gpu.py -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_runner.log 2>&1; echo "runner tests rc=$?"; tail -n 25 gpurun_out/t_runner.log | cut -c1-300
