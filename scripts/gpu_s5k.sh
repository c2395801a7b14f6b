# final validation of the round-1 HEAD: tests, smoke, bench lines, launch list, chain trace
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_rcan.err; cut -c1-300 gpurun_out/bench_rcan.json
for m in edsr rdn; do timeout 300 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | cut -c1-230; done
m=rcan
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/launches_all.csv python bench.py --model $m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$m.log 2>&1; echo "rc=$?"
python - $m <<'PY'
import csv, sys
m = sys.argv[1]
lines = [l for l in open('gpurun_out/launches_all.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'adam_kernel' in r.get('Kernel Name', '')]
print('launches captured', len(rows), 'adam launches', len(idx))
if len(idx) >= 2:
    a, b = idx[-2] + 1, idx[-1] + 1
    with open(f'gpurun_out/launches_{m}.csv', 'w') as f:
        w = csv.DictWriter(f, fieldnames=rows[0].keys()); w.writeheader(); w.writerows(rows[a:b])
PY
python scripts/summarize_launches.py gpurun_out/launches_$m.csv gpurun_out/launches_${m}_summary.txt > /dev/null; head -n 12 gpurun_out/launches_${m}_summary.txt | cut -c1-200
rm -f gpurun_out/launches_all.csv
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; grep -v "^trace: CTA0" gpurun_out/chain_bench.txt | tail -n 12 | cut -c1-330
