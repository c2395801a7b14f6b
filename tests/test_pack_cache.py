"""Host logic of the packed-weight caches (srb200/ops.py), no GPU: what is trusted while a pack table is installed.

Regression for the stale-weights bug: with a table installed, ANY cache hit whose source pointer matched was trusted, so a
packed copy first created after capture (full-size validation images take the per-layer path with its own packings) was
packed once and then served stale after every further optimizer step."""
import sys
import types

import pytest
import torch


@pytest.fixture()
def fake_ops(monkeypatch):
    from srb200 import ops
    calls = []

    def fake_pack_weight(w, packing, mode, shuffle=0, out=None):
        calls.append((w.data_ptr(), packing, mode))
        buf = out if out is not None else torch.empty(8, dtype=torch.uint8)
        buf.fill_(int(w.flatten()[0].item()) & 0xFF)       # "packed" content = first weight value
        return buf

    monkeypatch.setattr(ops, "pack_weight", fake_pack_weight)
    monkeypatch.setattr(ops, "_managed", None)
    monkeypatch.setattr(ops, "_generation", 0)
    return ops, calls


def _table(ops, packs_list):
    """What PackTable.__init__/install do, without a device: register the buffers cached right now."""
    keep = [packed for pw in packs_list for (_tag, packed, _src) in pw._cache.values()]
    ops._managed = frozenset(t.data_ptr() for t in keep)
    return keep


def test_entries_created_after_the_table_are_not_trusted(fake_ops):
    ops, calls = fake_ops
    w = torch.full((64, 64, 3, 3), 1.0)
    pw = ops.PackedWeights()
    bank = torch.empty(8, dtype=torch.uint8)
    pw.get(w, 1, 0, 0, out=bank)                   # training path: bank slice, registered in the table
    keep = _table(ops, [pw])
    assert len(calls) == 1
    # a training step: Adam rewrites w in place WITHOUT bumping the version; TrainStep.run() invalidates
    w.data.fill_(2.0)
    ops.invalidate_packed()
    assert pw.get(w, 1, 0, 0, out=bank) is bank and len(calls) == 1          # table entry: trusted (the table refreshes it)
    # validation on another shape: non-bank key, first use after capture
    v1 = pw.get(w, 1, 0, 0)
    assert len(calls) == 2 and int(v1[0]) == 2
    assert int(pw.get(w, 1, 0, 0)[0]) == 2 and len(calls) == 2               # same generation: cache hit
    # more training, then validation again: must re-pack (this returned the stale buffer before the fix)
    w.data.fill_(3.0)
    ops.invalidate_packed()
    v2 = pw.get(w, 1, 0, 0)
    assert len(calls) == 3 and int(v2[0]) == 3
    del keep


def test_without_a_table_version_and_generation_rule(fake_ops):
    ops, calls = fake_ops
    w = torch.full((4, 4, 3, 3), 5.0)
    pw = ops.PackedWeights()
    pw.get(w, 0, 0)
    pw.get(w, 0, 0)
    assert len(calls) == 1
    w.add_(1.0)                                    # torch op: version bump
    pw.get(w, 0, 0)
    assert len(calls) == 2
    ops.invalidate_packed()                        # kernel update without a version bump
    pw.get(w, 0, 0)
    assert len(calls) == 3


def test_fused_step_rejects_unsupported_losses_and_optimizers():
    """Runner / TrainStep must not silently train `0.5*l1+0.5*l2` or SGD configs as L1 / Adam (checked before any CUDA work)."""
    import models
    from srb200.trainer import TrainStep, check_supported
    with pytest.raises(NotImplementedError, match="single unit-weight L1"):
        TrainStep(models.EDSR(n_resblocks=1, losses="0.5*l1+0.5*l2"), (2, 3, 8, 8), 4)
    with pytest.raises(NotImplementedError, match="single unit-weight L1"):
        check_supported(models.EDSR(n_resblocks=1, losses="0.5*l1"))
    with pytest.raises(NotImplementedError, match="fuses Adam"):
        TrainStep(models.EDSR(n_resblocks=1, optimizer="SGD"), (2, 3, 8, 8), 4)
    check_supported(models.EDSR(n_resblocks=1, losses="mae"))


def test_registry_has_the_reference_names_and_fails_lazily():
    """Every loss / metric / optimizer name of the reference registry (srmodel.py:30-66) is known; entries whose optional
    package is missing construct fine as metrics and fail only when used."""
    import models
    from models import srmodel
    assert set(srmodel._supported_losses) == {"adaptive", "dists", "edge_loss", "flip", "haarpsi", "l1", "l2", "lpips", "mae", "mse",
                                              "pencil_sketch", "pieapp"}
    assert set(srmodel._supported_metrics) == {"BRISQUE", "FLIP", "LPIPS", "MS-SSIM", "PSNR", "SSIM"}
    assert set(srmodel._supported_optimizers) == {"ADAM", "Ranger", "RangerVA", "RangerQH", "RMSprop", "SGD"}
    m = models.EDSR(n_resblocks=1, metrics=["BRISQUE", "FLIP", "LPIPS", "MS-SSIM", "PSNR", "SSIM"])   # configs/train_default_sr.yml
    assert [n for n, _ in m._metrics] == ["BRISQUE", "FLIP", "LPIPS", "MS-SSIM", "PSNR", "SSIM"]
    with pytest.raises(AttributeError):
        models.EDSR(n_resblocks=1, metrics=["NOPE"])
