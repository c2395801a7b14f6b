set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-300
for m in rcan edsr rdn; do timeout 600 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | cut -c1-230; done
