"""Pins oracle/sr_oracle.py (the CPU restatement) against
 (a) the golden vectors produced by the unmodified reference classes (oracle/make_golden.py),
 (b) the plain-numpy primitive restatements in oracle/np_prims.py,
 (c) the live reference classes when /root/reference exists (build container only)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_util import Golden, golden_names, rel_l2, rel_max
from oracle import np_prims, sr_oracle
from oracle.ref_import import import_reference_models, reference_available

FAST = [n for n in golden_names() if n not in ("rcan_full_x4",)]


@pytest.mark.parametrize("name", golden_names())
def test_oracle_forward_matches_golden(name):
    g = Golden(name)
    sd = {k: torch.from_numpy(v).double() for k, v in g.state_dict().items()}
    x, _ = g.inputs()
    with torch.no_grad():
        sr = sr_oracle.FORWARDS[g.cls](torch.from_numpy(x).double(), sd, **g.oracle_cfg())
    assert sr.shape == g.sr.shape
    # golden is stored in fp32: 1e-6 is rounding of the stored values
    assert rel_max(sr.numpy(), g.sr) < 1e-6


@pytest.mark.parametrize("name", FAST)
def test_oracle_backward_matches_golden(name):
    g = Golden(name)
    x, hr = g.inputs()
    sr, loss, grads = sr_oracle.forward_backward(g.cls, x, hr, g.state_dict(), **g.oracle_cfg())
    assert abs(loss.item() - g.loss) < 1e-10
    assert rel_l2(grads["input"].numpy(), g.z["grad/input"]) < 1e-6
    for i, k in enumerate(g.grad_names):
        gk = grads[k].numpy()
        assert abs(np.sqrt((gk * gk).sum()) - g.grad_norm[i]) <= 1e-9 * max(1.0, g.grad_norm[i]), k
        pr = (gk * g.probe(k, gk.shape)).sum()
        assert abs(pr - g.grad_proj[i]) <= 1e-8 * max(abs(g.grad_proj[i]), g.grad_norm[i]), k
    for k, v in g.full_grads().items():
        if k == "input":
            continue
        assert rel_l2(grads[k].numpy(), v) < 1e-6, k


def test_numpy_primitives_match_torch():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 5, 7, 6))
    w = rng.standard_normal((4, 5, 3, 3))
    b = rng.standard_normal(4)
    y = np_prims.conv2d(x, w, b, 1)
    yt = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1)
    assert np.abs(y - yt.numpy()).max() < 1e-12
    # backward
    xt = torch.from_numpy(x).requires_grad_(True)
    wt = torch.from_numpy(w).requires_grad_(True)
    bt = torch.from_numpy(b).requires_grad_(True)
    gy = rng.standard_normal(y.shape)
    F.conv2d(xt, wt, bt, padding=1).backward(torch.from_numpy(gy))
    dx, dw, db = np_prims.conv2d_backward(x, w, gy, 1)
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-11
    assert np.abs(dw - wt.grad.numpy()).max() < 1e-11
    assert np.abs(db - bt.grad.numpy()).max() < 1e-11
    # 1x1 and 5x5
    for k in (1, 5):
        w2 = rng.standard_normal((3, 5, k, k))
        assert np.abs(np_prims.conv2d(x, w2, None, k // 2)
                      - F.conv2d(torch.from_numpy(x), torch.from_numpy(w2), padding=k // 2).numpy()).max() < 1e-11
    # pixel shuffle / unshuffle
    for r in (2, 3):
        z = rng.standard_normal((2, 4 * r * r, 3, 5))
        ps = np_prims.pixel_shuffle(z, r)
        assert np.array_equal(ps, F.pixel_shuffle(torch.from_numpy(z), r).numpy())
        assert np.array_equal(np_prims.pixel_unshuffle(ps, r), z)
    # channel attention
    c = 32
    xx = rng.standard_normal((2, c, 5, 4))
    w1 = rng.standard_normal((2, c, 1, 1)); b1 = rng.standard_normal(2)
    w2 = rng.standard_normal((c, 2, 1, 1)); b2 = rng.standard_normal(c)
    sd = {"ca.conv_du.0.weight": torch.from_numpy(w1), "ca.conv_du.0.bias": torch.from_numpy(b1),
          "ca.conv_du.2.weight": torch.from_numpy(w2), "ca.conv_du.2.bias": torch.from_numpy(b2)}
    ref = sr_oracle.ca_layer(torch.from_numpy(xx), sd, "ca").numpy()
    assert np.abs(np_prims.ca_layer(xx, w1, b1, w2, b2) - ref).max() < 1e-12


def test_psnr_restatement():
    a = torch.rand(2, 3, 8, 8, dtype=torch.float64)
    b = (a + 0.1).clamp(0, 1)
    mse = ((a - b) ** 2).flatten(1).mean(1)
    expect = (-10 * torch.log10(mse + 1e-8)).mean()
    assert abs(sr_oracle.psnr(a, b).item() - expect.item()) < 1e-12
    assert sr_oracle.psnr(a, a).item() == pytest.approx(80.0, abs=1e-9)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("name", ["edsr_small_x2", "rcan_small_x4", "rdn_a_x2", "srcnn_x2"])
def test_oracle_matches_live_reference(name):
    g = Golden(name)
    ref = import_reference_models()
    model = getattr(ref, g.cls)(**g.kwargs)
    # default (torch) initialisation this time: independent of synth weights
    torch.manual_seed(3)
    for p in model.parameters():
        if p.requires_grad:
            torch.nn.init.uniform_(p, -0.08, 0.08)
    model = model.double()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, hr = g.inputs()
    xt = torch.from_numpy(x).double()
    sr_ref = model.forward(xt)
    loss_ref = torch.nn.L1Loss()(sr_ref, torch.from_numpy(hr).double())
    loss_ref.backward()
    sr, loss, grads = sr_oracle.forward_backward(g.cls, x, hr, sd, **g.oracle_cfg())
    assert rel_max(sr.numpy(), sr_ref.detach().numpy()) < 1e-12
    assert abs(loss.item() - loss_ref.item()) < 1e-12
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert rel_l2(grads[k].numpy(), p.grad.numpy()) < 1e-10, k


@pytest.mark.parametrize("cls,kw,okw", [
    ("EDSR", dict(n_feats=64, n_resblocks=16, scale_factor=4), dict(n_feats=64, n_resblocks=16, scale=4)),
    ("EDSR", dict(n_feats=256, n_resblocks=3, scale_factor=3), dict(n_feats=256, n_resblocks=3, scale=3)),
    ("EDSR", dict(n_feats=32, n_resblocks=2, scale_factor=8), dict(n_feats=32, n_resblocks=2, scale=8)),
    ("RCAN", dict(n_resblocks=3, n_resgroups=2, scale_factor=4), dict(n_resblocks=3, n_resgroups=2, scale=4)),
    ("RDN", dict(rdn_config="B", scale_factor=4), dict(rdn_config="B", scale=4)),
    ("RDN", dict(rdn_config="A", scale_factor=2), dict(rdn_config="A", scale=2)),
    ("SRCNN", dict(scale_factor=2), dict(scale=2)),
])
def test_state_shapes_match_the_plugin_modules(cls, kw, okw):
    """oracle.sr_oracle.state_shapes (what bench.py's reference arm builds its weights from, without touching this repo's
    models) lists the same keys, in the same order, with the same shapes as the drop-in modules — which
    tests/test_boundary.py in turn pins to the live reference classes."""
    import models
    from oracle import sr_oracle
    m = getattr(models, cls)(**kw)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == list(sr_oracle.state_shapes(cls, **okw).items())
    sd = sr_oracle.init_state(cls, seed=0, **okw)
    m.load_state_dict(sd)       # loads: same dtypes / shapes, frozen MeanShift entries included
