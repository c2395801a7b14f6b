set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t_gpu.log | cut -c1-300
timeout 600 python scripts/bench_infer4k.py --steps 5 --warmup 2 2>/dev/null | cut -c1-300
for m in rdn edsr rcan; do timeout 900 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | cut -c1-260; done
SRB200_NO_WIDE=1 timeout 900 python bench.py --model rdn --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
