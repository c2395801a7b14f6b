"""Row-strip tiled EDSR inference with per-layer halo exchange (srb200/tiled.py) against the
untiled forward of the same model: bit-identical (same kernels, same per-pixel summation order),
for ragged strip heights and both compute modes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("kw,shape,parts", [
    (dict(n_feats=64, n_resblocks=3, res_scale=1.0, scale_factor=4), (1, 3, 70, 40), 3),
    (dict(n_feats=256, n_resblocks=2, res_scale=0.1, scale_factor=4), (1, 3, 37, 24), 4),
    (dict(n_feats=64, n_resblocks=2, res_scale=1.0, scale_factor=2), (1, 3, 33, 16), 8),
])
def test_local_strips_bit_identical(kw, shape, parts, mode):
    import models
    from srb200.tiled import LocalExchange, TiledEDSR
    torch.manual_seed(1)
    m = models.EDSR(**kw)
    m.compute_dtype = mode
    m = m.cuda()
    x = torch.rand(*shape).cuda()
    with torch.no_grad():
        full = m.forward(x)
        tiled = TiledEDSR(m, LocalExchange(parts)).forward_gathered(x)
    assert tiled.shape == full.shape
    assert torch.equal(tiled, full), (tiled - full).abs().max().item()
