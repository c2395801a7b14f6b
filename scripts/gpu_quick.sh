# quick GPU round: kernel + model parity tests, kernel microbench, RCAN/EDSR bench (no ncu)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -x > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 12 gpurun_out/t_kernels.log | cut -c1-300
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_trainer_gpu.py tests/test_tiled_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/t_models.log 2>&1; echo "models rc=$?"; tail -n 12 gpurun_out/t_models.log | cut -c1-300
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; tail -n 16 gpurun_out/kernel_bench.txt
timeout 600 python bench.py --model edsr --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr.json 2> gpurun_out/bench_edsr.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_edsr.err; cut -c1-330 gpurun_out/bench_edsr.json
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-330 gpurun_out/bench_rcan.json
