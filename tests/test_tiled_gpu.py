"""Row-strip tiled EDSR inference with halo exchange (srb200/tiled.py) against the untiled forward of the same model:
bit-identical (same kernels, same per-pixel summation order), for ragged strip heights, both compute modes, and halos of
1-3 rows (an exchange every 1-3 layers).  With two or more GPUs: one strip per process, halo rows pushed through NVLink peer
memory by csrc/halo.cu, eager and as a CUDA graph."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("kw,shape,parts", [
    (dict(n_feats=64, n_resblocks=3, res_scale=1.0, scale_factor=4), (1, 3, 70, 40), 3),
    (dict(n_feats=256, n_resblocks=2, res_scale=0.1, scale_factor=4), (1, 3, 37, 24), 4),
    (dict(n_feats=64, n_resblocks=2, res_scale=1.0, scale_factor=2), (1, 3, 33, 16), 8),
])
@pytest.mark.parametrize("halo", [1, 2, 3])
def test_local_strips_bit_identical(kw, shape, parts, mode, halo):
    import models
    from srb200.tiled import LocalExchange, TiledEDSR
    torch.manual_seed(1)
    m = models.EDSR(**kw)
    m.compute_dtype = mode
    m = m.cuda()
    x = torch.rand(*shape).cuda()
    with torch.no_grad():
        full = m.forward(x)
        tiled = TiledEDSR(m, LocalExchange(parts), halo=halo).forward_gathered(x)
    assert tiled.shape == full.shape
    assert torch.equal(tiled, full), (tiled - full).abs().max().item()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_exchange_two_processes_bit_identical():
    """torchrun x2: PeerExchange (CUDA-IPC strip buffers, srb_halo_exchange) vs the untiled forward on each rank's own GPU."""
    env = dict(os.environ, TORCH_NCCL_ASYNC_ERROR_HANDLING="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "tiled_peer_check.py"), "small", "5"],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "bit-identical" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
