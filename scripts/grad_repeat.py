"""Run forward+backward of a golden case repeatedly in one process and report which gradients change between repetitions
(debugging aid for timing-dependent errors): python scripts/grad_repeat.py <case> <fp32|bf16> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import models
from golden_util import Golden
name, mode = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g = Golden(name)
m = getattr(models, g.cls)(**g.kwargs)
m.load_state_dict({k: torch.from_numpy(v) for k, v in g.state_dict().items()})
m.compute_dtype = mode
m = m.cuda()
x, hr = g.inputs()
base = None
for r in range(reps):
    for p in m.parameters():
        p.grad = None
    xd = torch.from_numpy(x).cuda().requires_grad_(True)
    out = m.training_step({"lr": xd, "hr": torch.from_numpy(hr).cuda()}, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    cur = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.requires_grad}
    cur["input"] = xd.grad.detach().clone()
    if base is None:
        base = cur
        continue
    bad = []
    for k in cur:
        d = float((cur[k] - base[k]).norm() / base[k].norm().clamp_min(1e-30))
        if d > 1e-5:
            bad.append((k, d))
    print(f"rep {r}: loss {out['loss'].item():.8f}  tensors off by > 1e-5: {len(bad)}", [(k, f'{d:.1e}') for k, d in bad[-4:]])
