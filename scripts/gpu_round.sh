# one GPU round: kernel tests (wgrad umma), trainer tests, kernel microbench, bench, ncu launch list
python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider -k "boundary" > gpurun_out/t_cpu.log 2>&1 || { tail -5 gpurun_out/t_cpu.log; exit 3; }
set -x
mkdir -p gpurun_out
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/t_kernels.log
tail -n 25 gpurun_out/t_kernels.log | cut -c1-250
timeout 900 python -m pytest tests/test_models_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/t_models.log 2>&1; echo "rc=$?" >> gpurun_out/t_models.log
tail -n 25 gpurun_out/t_models.log | cut -c1-250
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_tiled_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/t_trainer.log 2>&1; echo "rc=$?" >> gpurun_out/t_trainer.log
tail -n 25 gpurun_out/t_trainer.log | cut -c1-250
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt | tail -n 15
timeout 600 python bench.py --model edsr --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr.json 2> gpurun_out/bench_edsr.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_edsr.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err
cut -c1-1200 gpurun_out/bench_edsr.json; cut -c1-1200 gpurun_out/bench_rcan.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 6000 -c 2200 --csv --log-file gpurun_out/launches_rcan.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_rcan.csv gpurun_out/launches_rcan_summary.txt | head -n 25
