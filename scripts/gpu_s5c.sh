# session-5 call C: deferred-gate chain ops — unit tests first, then the suite, A/B benches, chain trace
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -m gpu --tb=short -p no:cacheprovider -k "deferred" > gpurun_out/t_gate.log 2>&1; echo "gate tests rc=$?"; tail -n 30 gpurun_out/t_gate.log | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/t_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_rcan.err | tee gpurun_out/bench_rcan.json | cut -c1-230; tail -n 3 gpurun_out/bench_rcan.err
SRB200_CHAIN_GATE=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_rcan_nogate.err | tee gpurun_out/bench_rcan_nogate.json | cut -c1-230
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; grep -v "^trace: CTA0" gpurun_out/chain_bench.txt | tail -n 12 | cut -c1-330
