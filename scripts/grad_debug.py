"""Per-parameter gradient errors of a golden case against the oracle (debugging aid): python scripts/wdsr_debug.py <case> <fp32|bf16>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import models
from golden_util import Golden
from oracle import sr_oracle
name, mode = sys.argv[1], sys.argv[2]
g = Golden(name)
m = getattr(models, g.cls)(**g.kwargs)
m.load_state_dict({k: torch.from_numpy(v) for k, v in g.state_dict().items()})
m.compute_dtype = mode
m = m.cuda()
x, hr = g.inputs()
xd = torch.from_numpy(x).cuda().requires_grad_(True)
out = m.training_step({"lr": xd, "hr": torch.from_numpy(hr).cuda()}, 0)
out["loss"].backward()
torch.cuda.synchronize()
_, _, ref = sr_oracle.forward_backward(g.cls, x, hr, g.state_dict(), **g.oracle_cfg())
for k, p in m.named_parameters():
    if p.requires_grad:
        a, b = p.grad.double().cpu(), ref[k]
        e = float((a - b).norm() / b.norm().clamp_min(1e-30))
        print(f"{k:28s} {tuple(p.shape)!s:20s} rel {e:.3e}  |g| {float(a.norm()):.3e} ref {float(b.norm()):.3e} finite={bool(torch.isfinite(a).all())}")
a, b = xd.grad.double().cpu(), ref["input"]
print("input rel", float((a - b).norm() / b.norm()))
