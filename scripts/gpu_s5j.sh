# per-tile flag protocol (SRB200_CHAIN_TILEFLAGS=1): parity tests, then A/B benches
set -x
mkdir -p gpurun_out
SRB200_CHAIN_TILEFLAGS=1 timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_models_gpu.py tests/test_trainer_gpu.py -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_tileflags.log 2>&1; echo "tileflags tests rc=$?"; tail -n 8 gpurun_out/t_tileflags.log | cut -c1-300
for m in rcan edsr; do
SRB200_CHAIN_TILEFLAGS=1 timeout 300 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_${m}_tf.err | tee gpurun_out/bench_${m}_tf.json | cut -c1-230
timeout 300 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | cut -c1-230
done
