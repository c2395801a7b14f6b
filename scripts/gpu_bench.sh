# usage: bash scripts/gpu_bench.sh  (on the GPU box, from the repo root)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_models_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/t_models.log 2>&1; echo "rc=$?" >> gpurun_out/t_models.log
tail -n 30 gpurun_out/t_models.log | cut -c1-300
timeout 600 python bench.py --model edsr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_edsr.json 2> gpurun_out/bench_edsr.err; echo "rc=$?"; tail -n 5 gpurun_out/bench_edsr.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_rcan.json 2> gpurun_out/bench_rcan.err; echo "rc=$?"; tail -n 5 gpurun_out/bench_rcan.err
cat gpurun_out/bench_edsr.json gpurun_out/bench_rcan.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 6000 --csv --log-file gpurun_out/launches_rcan.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows = []
try:
    with open('gpurun_out/launches_rcan.csv') as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.DictReader(lines)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        name = row.get('Kernel Name', '')[:90]
        try: v = float(row.get('Metric Value', '0').replace(',', ''))
        except ValueError: continue
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open('gpurun_out/launches_rcan_summary.txt', 'w') as out:
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            line = f"{v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}% n={v[0]:5d} avg={v[1]/v[0]/1e3:7.2f} us  {k}"
            print(line); out.write(line + "\n")
except Exception as e:
    print("summary failed", e)
PY
