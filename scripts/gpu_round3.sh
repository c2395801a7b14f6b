# round: full -m gpu tests, chain microbench, RCAN + EDSR bench, clean launch list of one replayed step, ncu --set full of chain launches
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/t_gpu.log | cut -c1-300
timeout 300 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; grep -v "^trace: CTA0" gpurun_out/chain_bench.txt | tail -n 18 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-300 gpurun_out/bench_rcan.json
timeout 600 python bench.py --model edsr --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr.json 2> gpurun_out/bench_edsr.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_edsr.err; cut -c1-300 gpurun_out/bench_edsr.json
SRB200_NO_CHAIN=1 timeout 600 python bench.py --model edsr --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr_nochain.json 2> /dev/null; cut -c1-200 gpurun_out/bench_edsr_nochain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
python - <<'PY'
# keep the launches of the LAST replayed step: the last adam_kernel launch closes a step, the previous one opens it
import csv
lines = [l for l in open('gpurun_out/launches_all.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'adam_kernel' in r.get('Kernel Name', '')]
print('launches captured', len(rows), 'adam launches', len(idx))
if len(idx) >= 2:
    a, b = idx[-2] + 1, idx[-1] + 1
    # the timed step of bench.py --steps 1 is the last full adam-to-adam span before the roofline microbench
    step = rows[a:b]
    with open('gpurun_out/launches_rcan.csv', 'w') as f:
        w = csv.DictWriter(f, fieldnames=rows[0].keys()); w.writeheader(); w.writerows(step)
PY
python scripts/summarize_launches.py gpurun_out/launches_rcan.csv gpurun_out/launches_rcan_summary.txt | head -n 24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 42 -c 2 -o gpurun_out/prof_chain_rcan python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_chain_full.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/ncu_chain_full.log
rm -f gpurun_out/launches_all.csv
