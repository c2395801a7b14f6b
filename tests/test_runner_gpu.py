"""srb200.runner.Runner: Lightning-free fit / validate / predict / checkpoint-resume around the B200 path
(reference loops: srmodel.py:160-171 training_step, :214-232 validation_step, :375-380 predict_step)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(n_feats=64, n_resblocks=2, n_resgroups=2, reduction=16, scale_factor=4)


def _model(sd=None):
    import models
    torch.manual_seed(5)
    m = models.RCAN(**KW)
    if sd is not None:
        m.load_state_dict(sd)
    m.compute_dtype = "bf16"
    return m.cuda()


def _batches(steps, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"lr": torch.rand(2, 3, 16, 24, generator=g), "hr": torch.rand(2, 3, 64, 96, generator=g)} for _ in range(steps)]


def test_fit_reduces_loss_and_validate_matches_psnr_formula():
    from srb200.runner import Runner
    m = _model()
    r = Runner(m, (2, 3, 16, 24), lr=1e-3)
    try:
        batch = _batches(1)
        losses = r.fit(batch, epochs=12)          # the same batch 12 times: the loss must go down
        assert len(losses) == 12 and r.global_step == 12
        assert losses[-1] < 0.9 * losses[0], losses
        val = _batches(2, seed=9)
        metrics = r.validate(val)
        preds = r.predict(val)
        assert all(p.shape == b["hr"].shape and p.min() >= 0 and p.max() <= 1 for p, b in zip(preds, val))
        key = [k for k in metrics if k.endswith("PSNR")][0]
        want = 0.0
        for p, b in zip(preds, val):
            mse = ((p.double() - b["hr"].double().clamp(0, 1)) ** 2).flatten(1).mean(1)
            want += float((-10 * torch.log10(mse + 1e-8)).mean())
        assert abs(metrics[key] - want / 2) < 1e-3, (metrics, want / 2)
    finally:
        r.close()


def test_eval_after_fit_uses_the_updated_weights():
    """The captured step packs weights before its forward; validate()/predict() must re-pack so that the
    tensor-core copies match the fp32 parameters after the last Adam update."""
    from srb200.runner import Runner
    m = _model()
    r = Runner(m, (2, 3, 16, 24), lr=1e-2)      # a large step so that one update changes the output visibly
    try:
        r.fit(_batches(3))
        val = _batches(1, seed=4)
        got = r.predict(val)[0]
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    finally:
        r.close()
    fresh = _model(sd)
    with torch.no_grad():
        want = fresh.predict_step({"lr": val[0]["lr"].cuda()}, 0).cpu()
    # (the CALayer pool is summed with atomics: the gate may differ in its last fp32 bit between two runs; one
    # stale Adam update at lr=1e-2 would move the output by far more than this bound)
    assert ((got.double() - want.double()).norm() / want.double().norm()) < 2e-3


def test_checkpoint_resume_continues_the_run(tmp_path):
    from srb200.runner import Runner
    data = _batches(6, seed=2)
    m = _model()
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    r = Runner(m, (2, 3, 16, 24))
    try:
        r.fit(data[:4])
        path = str(tmp_path / "ckpt.pt")
        r.save_checkpoint(path)
        tail_a = r.fit(data[4:])
        w_a = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    finally:
        r.close()
    m2 = _model(sd0)
    r2 = Runner(m2, (2, 3, 16, 24))
    try:
        r2.load_checkpoint(path)
        assert r2.global_step == 4
        tail_b = r2.fit(data[4:])
        assert r2.global_step == 6
        w_b = {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()}
    finally:
        r2.close()
    # Split weight-gradient reductions and the CALayer pool add with atomics (order-dependent in the last fp32 bit, which
    # can flip a bf16 rounding downstream), so "identical" means far inside what a lost optimizer state would cause:
    # without the Adam moments or step count the next updates differ by O(lr) = 1e-3 per weight, i.e. ~3e-2 relative.
    for a, b in zip(tail_a, tail_b):
        assert abs(a - b) < 1e-4 * abs(a), (tail_a, tail_b)
    for k in w_a:
        assert ((w_a[k].double() - w_b[k].double()).norm() <= 1e-3 * w_a[k].double().norm() + 1e-12), k
