import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch
import torch.nn.functional as F
from srb200 import lib as L, ops
dev = torch.device("cuda", 0)
bf = torch.bfloat16
for (n, h, w) in [(5, 33, 7), (1, 16, 7), (1, 33, 8), (5, 16, 8), (3, 48, 8), (1, 17, 8), (2, 32, 8)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, 64, generator=g).to(dev).to(bf)
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(dev)
    b = (torch.randn(64, generator=g) * 0.1).to(dev)
    pk = ops.PackedWeights()
    y1 = torch.empty_like(x)
    ops.conv(x, 0, 64, pk, wt, b, y1, 0, 64, 3, relu=True)
    bank = ops.FilterBank().get([(wt, pk)], L.PACK_FWD)
    A = torch.zeros((1, n, h, w, 64), dtype=bf, device=dev)
    ch = ops.Chain(n, h, w, dev); ch.space(0, A); ch.space(1, x.view(1, n, h, w, 64))
    ch.conv(ops.Chain.ref(1, 0), ops.Chain.ref(0, 0), 0, b, relu=True)
    ch.run(bank)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(bf).float(), b, padding=1)).permute(0, 2, 3, 1)
    d_layer = (y1.float() - ref).abs()
    d_chain = (A[0].float() - ref).abs()
    bad = (A[0] != y1).nonzero()
    print(f"shape {(n,h,w)}: layer-vs-torch max {d_layer.max().item():.4f}  chain-vs-torch max {d_chain.max().item():.4f}  "
          f"mismatching elems {bad.shape[0]}  first {bad[:3].tolist()} last {bad[-3:].tolist()}")
