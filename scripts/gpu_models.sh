mkdir -p gpurun_out
for m in edsr rdn; do
  timeout 400 python bench.py --model $m --workload train --no-extras 2>/dev/null | tail -1 > gpurun_out/r02_bench_$m.json
  python - $m <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02_bench_{sys.argv[1]}.json").read())
r=d.get("roofline") or {}
print(sys.argv[1], round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(r.get("achieved",0),1), round(r.get("frac",0),3), "step", d.get("roofline_step",{}).get("achieved"))
PY
done
