"""tcgen05.mma throughput probe (scripts/probes/umma_rate_probe.cu): SM cycles per 128xNx16 bf16 MMA, SS operands."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L  # noqa: E402

# The probe is NOT part of libsrb200.so: it is built here on first use, against the product library's context / PTX helpers.
CSRC = os.path.join(ROOT, "sr-pytorch-lightning_b200", "csrc")
SO = os.path.join(ROOT, "scripts", "probes", "libumma_rate_probe.so")
if not os.path.isfile(SO):
    import subprocess
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                    "-I", CSRC, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "scripts", "probes", "umma_rate_probe.cu"),
                    os.path.join(CSRC, "libsrb200.so"), "-o", SO], check=True)
L.load()
lib = C.CDLL(SO)
lib.srb_probe_umma.restype = C.c_int
lib.srb_probe_umma.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
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
