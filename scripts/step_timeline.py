"""Kernel timeline of one captured RCAN training step (CUPTI through torch.profiler): start / duration / stream of every
kernel, so that what overlaps what (backward chain launches vs weight-gradient launches on the side stream) is visible.
  python scripts/step_timeline.py [out.txt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import models  # noqa: E402
from srb200.trainer import TrainStep  # noqa: E402

torch.manual_seed(0)
m = models.RCAN(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4)
m.compute_dtype = "bf16"
m = m.to("cuda:0")
ts = TrainStep(m, (16, 3, 48, 48), 4, lr=1e-4)
x = torch.rand(16, 3, 48, 48).cuda()
h = torch.rand(16, 3, 192, 192).cuda()
ts.prepare()
for _ in range(3):
    ts.step(x, h)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(x, h)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
lines = []
short = lambda n: n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0][:36]  # noqa: E731
streams = {}
for e in evs:
    st = getattr(e, "stream", None)
    if st is None:
        st = e.device_resource_id if hasattr(e, "device_resource_id") else 0
    sid = streams.setdefault(st, len(streams))
    lines.append((e.time_range.start - t0, e.time_range.end - t0, sid, short(e.name)))
out = []
out.append(f"{len(lines)} device activities, step span {lines[-1][1] if lines else 0:.0f} us, streams: {len(streams)}")
for s, e, sid, n in lines:
    if e - s >= 20 or "chain" in n or "wgrad" in n:
        out.append(f"{s:9.1f} {e:9.1f} {e - s:8.1f} us  s{sid}  {n}")
# overlap summary: time covered by chain kernels, by wgrad kernels, by both
def cover(pred):
    iv = sorted((s, e) for s, e, _, n in lines if pred(n))
    merged = []
    for s, e in iv:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    return merged
def total(iv):
    return sum(e - s for s, e in iv)
def inter(a, b):
    t, i, j = 0.0, 0, 0
    while i < len(a) and j < len(b):
        lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
        if hi > lo:
            t += hi - lo
        if a[i][1] < b[j][1]:
            i += 1
        else:
            j += 1
    return t
ch = cover(lambda n: "chain" in n)
wg = cover(lambda n: "wgrad" in n)
out.append(f"chain kernels cover {total(ch):.0f} us, weight-gradient kernels cover {total(wg):.0f} us, both at once {inter(ch, wg):.0f} us")
text = "\n".join(out)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text + "\n")
