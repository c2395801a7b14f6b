set -x
mkdir -p gpurun_out
timeout 300 python scripts/probe.py > gpurun_out/probe.txt 2>&1; cat gpurun_out/probe.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q --tb=short -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log; tail -n 8 gpurun_out/t_all.log | cut -c1-250
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; cat gpurun_out/kernel_bench.txt | tail -n 15
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-400 gpurun_out/bench_rcan.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_resident -s 30 -c 2 -o gpurun_out/prof_conv_resident python scripts/kernel_bench.py > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
