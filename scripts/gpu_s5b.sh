# session-5 call B: tests (default + single-window weight gradient), A/B benches, chain trace
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/t_gpu.log | cut -c1-300
SRB200_WGRAD_ONEWIN=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -q -m gpu -k "wgrad or trainer or train" --tb=line -p no:cacheprovider > gpurun_out/t_onewin.log 2>&1; echo "onewin tests rc=$?"; tail -n 8 gpurun_out/t_onewin.log | cut -c1-300
for m in rcan rdn edsr; do
timeout 600 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | cut -c1-230
SRB200_WGRAD_ONEWIN=1 timeout 600 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_${m}_onewin.err | tee gpurun_out/bench_${m}_onewin.json | cut -c1-230
done
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; grep -v "^trace: CTA0" gpurun_out/chain_bench.txt | tail -n 12 | cut -c1-330
