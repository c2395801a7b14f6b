# chain-path round: full -m gpu tests, chain microbench, RCAN + EDSR bench, launch list
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/t_gpu.log | cut -c1-300
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; tail -n 14 gpurun_out/chain_bench.txt | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-400 gpurun_out/bench_rcan.json
SRB200_NO_CHAIN=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rcan_nochain.json 2> gpurun_out/bench_rcan_nochain.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_rcan_nochain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 400 -c 800 --csv --log-file gpurun_out/launches_rcan.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_rcan.csv gpurun_out/launches_rcan_summary.txt | head -n 25
