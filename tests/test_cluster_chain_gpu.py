"""Per-sample cluster chain kernel (conv_cluster.cu, taken by srb_conv_chain for H % 16 == 0, W in {24, 48}) against
the one-launch-per-layer kernels (tests/test_kernels_gpu.py pins those to torch fp32 references).  Plain conv ops run
the same MMA sequence and epilogue arithmetic, so they must agree BIT FOR BIT on every cluster shape (1 to 8 CTAs per
sample: no neighbour, side / vertical / diagonal neighbours); the fused CALayer evaluates its gate in a different
summation order and is compared within bf16 rounding."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def _dbg():
    import cluster_debug
    return cluster_debug


@pytest.mark.parametrize("shape", [(1, 16, 24), (1, 16, 48), (1, 32, 24), (3, 32, 48), (2, 48, 48), (16, 48, 48), (2, 64, 48), (20, 48, 48)])
def test_cluster_convs_bit_exact(shape):
    """relu conv -> residual conv (x 0.5, residual read in place from shared memory) -> masked conv (TMA operand tile)
    with column sums: bit-identical to three srb_conv launches; cluster sizes 1, 2 (side), 2 (vertical), 4, 6, 6, 8."""
    assert _dbg().stage_convs(shape, 3)


def test_cluster_long_dependency_chain_bit_exact():
    """24 dependent relu convs on the bench shape, three repetitions: a halo or window consumed before its producer's
    stores are visible (DSMEM ordering bug) shows up as a mismatch."""
    assert _dbg().stage_long(reps=3)


@pytest.mark.parametrize("shape", [(2, 16, 24), (3, 32, 48), (16, 48, 48)])
def test_cluster_ca_forward(shape):
    assert _dbg().stage_ca(shape)


@pytest.mark.parametrize("shape,blocks", [((2, 16, 24), 2), ((2, 16, 48), 2), ((3, 32, 48), 2), ((16, 48, 48), 3), ((2, 64, 48), 2)])
def test_cluster_rcab_chain_forward(shape, blocks):
    """conv1+ReLU -> conv2+CALayer+skip chains of several RCABs followed by a plain conv, on cluster shapes with every border
    combination, against the per-layer kernels."""
    assert _dbg().stage_rcab(shape, blocks)


@pytest.mark.parametrize("shape", [(2, 16, 24), (3, 32, 48), (16, 48, 48)])
def test_cluster_ca_backward_fused(shape):
    """dgrad + residual with the CALayer backward fused in, followed by a masked conv that consumes dt through the halos;
    the residual of a later op comes back from TMEM."""
    assert _dbg().stage_cabwd(shape)


@pytest.mark.parametrize("model", ["edsr", "rcan"])
def test_cluster_model_matches_flag_kernel_and_layer_path(model):
    """Whole model on 48x48 patches, forward + L1 + backward: cluster kernel vs the L2-flag chain kernel
    (SRB200_CHAIN_CLUSTER=0) vs the per-layer path (SRB200_NO_CHAIN=1).  Separate process: the stage flips env switches."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "cluster_debug.py"), model], capture_output=True, text=True,
                       timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_cluster_is_taken_for_the_bench_shape_and_not_for_ragged_ones():
    from srb200 import lib as L, ops
    import ctypes as C
    for (n, h, w), want in (((16, 48, 48), True), ((4, 24, 24), False), ((1, 48, 40), False), ((1, 144, 48), False)):
        x = torch.zeros((1, n, h, w, 64), dtype=torch.bfloat16, device="cuda:0")
        A = torch.zeros((1, n, h, w, 64), dtype=torch.bfloat16, device="cuda:0")
        ch = ops.Chain(n, h, w, x.device)
        ch.space(0, A)
        ch.space(1, x)
        ch.conv(ops.Chain.ref(1, 0), ops.Chain.ref(0, 0), 0, None, relu=True)
        ch.run(torch.zeros(ops.CHAIN_LAYER_BYTES, dtype=torch.uint8, device="cuda:0"))
        torch.cuda.synchronize()
        assert ch.used_cluster is want, (n, h, w)
