"""Times srb_pack_table on the RCAN step's filter set (822 3x3 64->64 weights x forward/DGRAD copies) with CUDA events."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

dev = torch.device("cuda", 0)
n_w = 822
w = torch.randn(n_w, 64, 64, 3, 3, device=dev)
nbytes = ops.CHAIN_LAYER_BYTES
dst = torch.zeros(2 * n_w * nbytes, dtype=torch.uint8, device=dev)
rows = []
for i in range(n_w):
    for m, mode in enumerate((L.PACK_FWD, L.PACK_DGRAD)):
        rows.append((w[i].data_ptr(), dst.data_ptr() + (2 * i + m) * nbytes, 64, 64, 3, L.PACK_UMMA, mode, 0))
dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("ksize", "<i4"),
               ("packing", "<i4"), ("mode", "<i4"), ("shuffle", "<i4")])
table = torch.from_numpy(np.array(rows, dtype=dt).view(np.uint8).copy()).to(dev)
lib = L.load()


def run():
    L.check(lib.srb_pack_table(C.c_void_p(L.ctx(0)), C.c_void_p(table.data_ptr()), len(rows), int(os.environ.get("PACK_GRID_ELEMS", 64 * 64 * 9)),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "srb_pack_table")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(10):
    flush.zero_()                      # evict the weights from L2 (the step reads them after 7.9 ms of other traffic)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
mb = (n_w * 64 * 64 * 9 * 4 * 2 + 2 * n_w * nbytes) / 1e6
t = sorted(ts)[len(ts) // 2]
print(f"pack_table: {t:.1f} us median of 10 for {len(rows)} items, "
      f"{mb:.0f} MB -> {mb / t * 1e-3 * 1e3:.2f} TB/s   checksum {int(dst.view(torch.int16).sum().item())}")
