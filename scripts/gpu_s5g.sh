set -x
mkdir -p gpurun_out
m=edsr
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/launches_all.csv python bench.py --model $m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$m.log 2>&1; echo "rc=$?"
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
python scripts/summarize_launches.py gpurun_out/launches_$m.csv gpurun_out/launches_${m}_summary.txt > /dev/null; cat gpurun_out/launches_${m}_summary.txt | cut -c1-200
python - <<'PY'
import csv
rows = list(csv.DictReader(open('gpurun_out/launches_edsr.csv')))
for r in rows:
    print(r.get('Kernel Name','')[:70], r.get('Metric Value'), r.get('Metric Unit'), r.get('Grid Size'), sep=' | ')
PY
rm -f gpurun_out/launches_all.csv
