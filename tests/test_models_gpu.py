"""Model-level parity (GPU): this repo's EDSR / RCAN / RDN / SRCNN, driven through the SRModel
plugin API on cuda:0, against (a) the committed golden vectors produced by the unmodified
reference classes and (b) the CPU oracle run live on the same inputs.

Bars (BASELINE.json north_star): fp32 mode <= 1e-4 relative, bf16 mode <= 2e-2 relative, for the
output and for every parameter gradient after one L1 step.  "relative" = ||a-b||_2 / ||b||_2; for
the output it is also evaluated before add_mean (the +0.44 offset flatters relative error)."""
import numpy as np
import pytest
import torch

from golden_util import Golden, RGB_MEAN, golden_names, rel_l2, rel_max

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}
# per-parameter gradient bar; tiny late-layer gradients in bf16 carry more rounding noise
GRAD_TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _build(g: Golden, mode: str):
    import models
    model = getattr(models, g.cls)(**g.kwargs)
    sd = {k: torch.from_numpy(v) for k, v in g.state_dict().items()}
    model.load_state_dict(sd)
    model.compute_dtype = mode
    return model.to("cuda:0")


def _minus_mean(a, has_mean):
    if not has_mean:
        return a
    return a - np.array(RGB_MEAN, dtype=np.float64).reshape(1, 3, 1, 1)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", golden_names())
def test_forward_backward_matches_golden(name, mode):
    g = Golden(name)
    model = _build(g, mode)
    x, hr = g.inputs()
    xd = torch.from_numpy(x).to("cuda:0").requires_grad_(True)
    out = model.training_step({"lr": xd, "hr": torch.from_numpy(hr).to("cuda:0")}, 0)
    loss = out["loss"]
    sr = model._last_sr if hasattr(model, "_last_sr") else None
    loss.backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        sr = model.forward(torch.from_numpy(x).to("cuda:0")).float().cpu().numpy()
    tol = TOL[mode]
    has_mean = g.cls in ("EDSR", "RCAN")
    e_out = rel_l2(sr, g.sr)
    e_pre = rel_l2(_minus_mean(sr.astype(np.float64), has_mean), _minus_mean(g.sr.astype(np.float64), has_mean))
    assert e_out < tol and e_pre < tol, (name, mode, e_out, e_pre)
    assert abs(loss.item() - g.loss) < tol * max(abs(g.loss), 1e-3), (loss.item(), g.loss)
    # gradients: norms and probe projections for every parameter, full tensors where stored
    worst = 0.0
    grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad}
    assert list(grads) == g.grad_names
    gt = GRAD_TOL[mode]
    bad = []
    for i, k in enumerate(g.grad_names):
        gk = grads[k].double().cpu().numpy()
        n_ref = g.grad_norm[i]
        proj = float((gk * g.probe(k, gk.shape)).sum())
        e_n = abs(np.sqrt((gk * gk).sum()) - n_ref) / max(n_ref, 1e-30)
        e_p = abs(proj - g.grad_proj[i]) / max(n_ref, 1e-30)      # projection error relative to the norm
        worst = max(worst, e_n, e_p / 8)
        if e_n > gt or e_p > 8 * gt:
            bad.append((k, e_n, e_p))
    assert not bad, (name, mode, bad[:8], len(bad))
    for k, v in g.full_grads().items():
        got = xd.grad if k == "input" else grads[k]
        e = rel_l2(got.double().cpu().numpy(), v)
        assert e < gt, (name, mode, k, e)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("cls,kwargs,shape", [
    ("EDSR", dict(n_feats=64, n_resblocks=3, res_scale=1.0, scale_factor=4), (3, 3, 20, 28)),
    ("RCAN", dict(n_feats=64, n_resblocks=2, n_resgroups=2, reduction=16, scale_factor=2), (2, 3, 24, 16)),
    ("RDN", dict(rdn_config="B", scale_factor=3), (1, 3, 16, 24)),
])
def test_matches_live_oracle_random_weights(cls, kwargs, shape, mode):
    """Independent of the golden files: default (torch) initialisation, random inputs, ragged
    (non multiple-of-tile) image sizes, oracle evaluated on the CPU of this box."""
    import models
    from oracle import sr_oracle
    torch.manual_seed(11)
    model = getattr(models, cls)(**kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.compute_dtype = mode
    model = model.to("cuda:0")
    s = kwargs["scale_factor"]
    x = torch.rand(*shape)
    hr = torch.rand(shape[0], 3, shape[2] * s, shape[3] * s)
    cfg = {"scale": s}
    if cls == "EDSR":
        cfg.update(n_resblocks=kwargs["n_resblocks"], res_scale=kwargs["res_scale"])
    elif cls == "RCAN":
        cfg.update(n_resblocks=kwargs["n_resblocks"], n_resgroups=kwargs["n_resgroups"])
    else:
        cfg.update(rdn_config=kwargs["rdn_config"])
    sr_ref, loss_ref, grads_ref = sr_oracle.forward_backward(cls, x, hr, sd, **cfg)
    xd = x.to("cuda:0").requires_grad_(True)
    out = model.training_step({"lr": xd, "hr": hr.to("cuda:0")}, 0)
    out["loss"].backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        sr = model.forward(x.to("cuda:0"))
    tol = TOL[mode]
    assert rel_l2(sr.cpu().numpy(), sr_ref.numpy()) < tol
    assert abs(out["loss"].item() - loss_ref.item()) < tol * loss_ref.item()
    assert rel_l2(xd.grad.cpu().numpy(), grads_ref["input"].numpy()) < GRAD_TOL[mode]
    bad = []
    for k, p in model.named_parameters():
        if p.requires_grad:
            e = rel_l2(p.grad.cpu().numpy(), grads_ref[k].numpy())
            if e > GRAD_TOL[mode]:
                bad.append((k, e))
    assert not bad, (bad[:8], len(bad))


def test_psnr_parity_bf16():
    """x4 PSNR of the bf16 path within 0.01 dB of the fp64 oracle on identical weights/inputs."""
    import models
    from oracle import sr_oracle
    g = Golden("edsr_base_x4")
    model = _build(g, "bf16")
    x, hr = g.inputs()
    with torch.no_grad():
        sr = model.validation_step({"lr": torch.from_numpy(x).cuda(), "hr": torch.from_numpy(hr).cuda(), "path": "synthetic"}, 0)
    ours = [v for k, v in sr.items() if k.endswith("PSNR")][0].item()
    ref = sr_oracle.psnr(torch.from_numpy(g.sr).double(), torch.from_numpy(hr).double()).item()
    assert abs(ours - ref) < 0.01, (ours, ref)


def test_no_cpu_fallback():
    import models
    m = models.EDSR(n_feats=64, n_resblocks=1, scale_factor=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.forward(torch.rand(1, 3, 8, 8))
