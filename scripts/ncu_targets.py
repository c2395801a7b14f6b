"""Short single-kernel workloads for `ncu --set full` captures (one GPU):
  python scripts/ncu_targets.py wide128      one 3x3 256->256 conv (+bias+ReLU) on [1,540,960,256]: conv_wide_kernel<128>, 4K inference
  python scripts/ncu_targets.py wgrad [sms]  the 41 weight gradients of one RCAN ResidualGroup on [16,48,48,64] in one batched call
  python scripts/ncu_targets.py group        one RCAN ResidualGroup forward (conv_chain_kernel) + backward (chain_cluster_kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16
what = sys.argv[1]
torch.manual_seed(0)
if what == "wide128":
    x = torch.randn(1, 540, 960, 256, device=DEV).to(bf)
    y = torch.empty_like(x)
    w = (torch.randn(256, 256, 3, 3, device=DEV) * 0.02).contiguous()
    b = torch.zeros(256, device=DEV)
    pk = ops.PackedWeights()
    for _ in range(3):
        ops.conv(x, 0, 256, pk, w, b, y, 0, 256, 3, relu=True)
    torch.cuda.synchronize()
elif what == "wgrad":
    sms = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n = 41
    xs = [torch.randn(16, 48, 48, 64, device=DEV).to(bf) for _ in range(n)]
    gs = [(torch.randn(16, 48, 48, 64, device=DEV) * 0.01).to(bf) for _ in range(n)]
    dws = [torch.zeros(64, 64, 3, 3, device=DEV) for _ in range(n)]
    ctx = C.c_void_p(L.ctx(0))
    for _ in range(3):
        L.load().srb_set_wgrad_sm_budget(ctx, sms)
        with ops.deferred_wgrads(max_items=4096):
            for x, g, dw in zip(xs, gs, dws):
                ops.conv_wgrad(x, 0, 64, g, 0, 64, 3, dw, None, accumulate=True)
        L.load().srb_set_wgrad_sm_budget(ctx, 0)
    torch.cuda.synchronize()
else:
    import models
    from srb200.trainer import FlatParams
    grp = models.rcan.ResidualGroup(64, 3, 16, 1, 20).to(DEV)
    flat = FlatParams(grp)
    flat.begin_step(zero=True)
    x = torch.randn(16, 48, 48, 64, device=DEV).to(bf).requires_grad_(True)
    g = (torch.randn(16, 48, 48, 64, device=DEV) * 0.01).to(bf)
    for _ in range(3):
        with ops.deferred_wgrads() as q:
            grp(x).backward(g)
            q.items.clear()
    torch.cuda.synchronize()
print("done", what)
