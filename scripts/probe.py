"""tcgen05.mma throughput probe (srb_probe_umma): SM cycles per 128xNx16 bf16 MMA, SS operands."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L  # noqa: E402

lib = L.load()
ctx = C.c_void_p(L.ctx(0))
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for blocks in (1, 148):
    for mn in (0, 1):
        for n in (64, 128, 256):
            for distinct in (1, 4):
                iters = 4096
                L.check(lib.srb_probe_umma(ctx, n, iters, mn, distinct, blocks, C.c_void_p(out.data_ptr()), st))
                torch.cuda.synchronize()
                cyc = out[:blocks].double().mean().item() / iters
                macs = 128 * n * 16
                print(f"blocks={blocks:3d} {'MN' if mn else 'K '}-major N={n:3d} distinct={distinct}: {cyc:7.1f} cyc/MMA "
                      f"-> {macs / cyc:7.0f} MAC/cyc/SM  ({100 * macs / cyc / 4096:5.1f}% of 4096)")
