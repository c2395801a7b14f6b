import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row.get("Kernel Name", "")[:100]
    try:
        v = float(row.get("Metric Value", "0").replace(",", ""))
    except ValueError:
        continue
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values()) or 1.0
with open(dst, "w") as out:
    hdr = f"total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches"
    print(hdr)
    out.write(hdr + "\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}% n={v[0]:5d} avg={v[1]/v[0]/1e3:8.2f} us  {k}"
        print(line)
        out.write(line + "\n")
