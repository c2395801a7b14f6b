# HEAD validation round: full -m gpu test suite, RCAN bench (+cpu baseline), reference arm, launch list, one ncu --set full capture
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/t_gpu.log | cut -c1-300
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1; tail -n 16 gpurun_out/kernel_bench.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-600 gpurun_out/bench_rcan.json
timeout 600 python bench.py --model edsr --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr.json 2> gpurun_out/bench_edsr.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_edsr.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 6000 -c 2300 --csv --log-file gpurun_out/launches_rcan.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_rcan.csv gpurun_out/launches_rcan_summary.txt | head -n 25
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_c64 -s 30 -c 2 -o gpurun_out/prof_conv_c64 python scripts/kernel_bench.py > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
