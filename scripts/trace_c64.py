"""Per-CTA timeline of the conv_c64 kernel (srb_debug_set_trace): a dependent chain of 3x3 64->64
convs on [16,48,48,64] is replayed from a CUDA graph; the LAST launch's event clocks are printed
(SM cycles relative to each CTA's start; min / median / max over CTAs)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

EV = ["t0_ns", "prologue", "depwait", "w_full", "a0_full", "mma0_issued", "acc0_full", "epi0_done", "acc1_full", "epi1_done",
      "stored", "end", "t1_ns", "smid"]
dev = torch.device("cuda", 0)
N, H, W, Cc = 16, 48, 48, 64
bf = torch.bfloat16
variant = sys.argv[1] if len(sys.argv) > 1 else "relu"
chain = 24
xs = [torch.randn(N, H, W, Cc, device=dev).to(bf) * 0.1 for _ in range(chain + 1)]
zs = torch.randn(N, H, W, Cc, device=dev).to(bf)
w = torch.randn(64, 64, 3, 3, device=dev) * 0.04
b = torch.zeros(64, device=dev)
packs = ops.PackedWeights()
pool = torch.zeros(N, 64, device=dev)
lib = L.load()
ctx = C.c_void_p(L.ctx(0))
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


def one(i):
    kw = dict(relu=True) if variant == "relu" else dict(res=(zs, 0)) if variant == "res" else \
        dict(mask=(zs, 0), colsum=pool[0], colsum_groups=1, mode=L.PACK_DGRAD)
    bias = None if variant == "mask" else b
    ops.conv(xs[i], 0, 64, packs, w, bias, xs[i + 1], 0, 64, 3, **kw)


for i in range(3):
    one(i)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    L.check(lib.srb_debug_set_trace(ctx, C.c_void_p(trace.data_ptr())))
    with torch.cuda.graph(g):
        for i in range(chain):
            one(i)
    L.check(lib.srb_debug_set_trace(ctx, None))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
print(f"variant={variant}: {e0.elapsed_time(e1) * 1e3 / chain:.2f} us per dependent layer (chain of {chain}, graph replay)")
t = trace.view(148, 16).cpu()
ns = (t[:, 12] - t[:, 0]).double()
print(f"last launch: CTA lifetime {ns.min():.0f} / {ns.median():.0f} / {ns.max():.0f} ns (min/med/max); "
      f"first start -> last end {(t[:, 12].max() - t[:, 0].min()).item()} ns; start skew {(t[:, 0].max() - t[:, 0].min()).item()} ns")
for k in range(1, 12):
    col = t[:, k].double()
    col = col[col > 0]
    if len(col):
        print(f"  {EV[k]:12s} {col.min():8.0f} {col.median():8.0f} {col.max():8.0f} cycles")
