"""Loss trace of the first RCAN training steps, to compare runs (same seed): python scripts/overlap_determinism.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import models  # noqa: E402
from srb200.trainer import TrainStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
if "LOCAL_RANK" in os.environ:      # under torchrun: every rank gets the SAME batches, so the averaged gradient is the 1-GPU one
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}"))
torch.manual_seed(0)
m = models.RCAN(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4)
m.compute_dtype = "bf16"
m = m.to(f"cuda:{torch.cuda.current_device()}")
ts = TrainStep(m, (16, 3, 48, 48), 4, lr=1e-4)
g = torch.Generator().manual_seed(1)
xs = [torch.rand(16, 3, 48, 48, generator=g).cuda() for _ in range(2)]
hs = [torch.rand(16, 3, 192, 192, generator=g).cuda() for _ in range(2)]
ts.prepare()
out = []
for i in range(steps):
    out.append(float(ts.step(xs[i % 2], hs[i % 2]).item()))
print(f"rank {os.environ.get('RANK', 0)} losses", " ".join(f"{v:.6f}" for v in out))
gn = float(ts.flat.grad.double().norm())
print(f"grad norm {gn:.6e}  overlap sections {ts.overlap.sections_run if ts.overlap else 0}")
if "LOCAL_RANK" in os.environ:      # (tearing NCCL down under a live CUDA graph can hang: leave at once)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
