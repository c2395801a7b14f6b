# full status: GPU tests, default bench line (timed), train-only bench with the L2-flag chain for comparison
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/status_tests.log 2>&1
tail -5 gpurun_out/status_tests.log
( time timeout 900 python bench.py ) > gpurun_out/status_bench_default.log 2>&1
tail -c 6000 gpurun_out/status_bench_default.log
( time SRB200_CHAIN_CLUSTER=0 timeout 600 python bench.py --workload train --steps 20 --warmup 5 ) > gpurun_out/status_bench_flags.log 2>&1
tail -c 3000 gpurun_out/status_bench_flags.log
