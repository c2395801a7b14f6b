"""Layer-chain kernel (srb_conv_chain, include/srb200.h) against the one-launch-per-layer kernels,
which tests/test_kernels_gpu.py pins to torch fp32 references and tests/test_models_gpu.py to the
oracle / golden vectors.  Plain conv ops run the same MMA sequence and epilogue arithmetic in both
paths, so they must agree BIT FOR BIT; the fused CALayer evaluates the 64->Cr->64 gate in a
different summation order, so it is compared within bf16 rounding (tolerances in the tests)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(16, 48, 48), (2, 16, 16), (3, 20, 12), (1, 48, 48), (5, 33, 7), (40, 16, 8)])
def test_chain_convs_bit_exact(shape):
    """relu conv -> residual conv (x scale) -> masked dgrad-style conv with column sums, as ONE
    chain launch, versus three srb_conv launches."""
    from srb200 import lib as L, ops
    n, h, w = shape
    bf = torch.bfloat16
    x = _rand((n, h, w, 64), seed=1).to(bf)
    m = _rand((n, h, w, 64), seed=2).to(bf)
    ws = [(_rand((64, 64, 3, 3), 0.05, seed=10 + i)).contiguous() for i in range(3)]
    bs = [_rand((64,), 0.1, seed=20 + i) for i in range(3)]
    packs = [ops.PackedWeights() for _ in range(3)]
    # per-layer path
    y1 = torch.empty_like(x)
    ops.conv(x, 0, 64, packs[0], ws[0], bs[0], y1, 0, 64, 3, relu=True)
    y2 = torch.empty_like(x)
    ops.conv(y1, 0, 64, packs[1], ws[1], bs[1], y2, 0, 64, 3, scale=0.5, res=(x, 0))
    y3 = torch.empty_like(x)
    cs = torch.zeros(64, device=DEV)
    ops.conv(y2, 0, 64, packs[2], ws[2], None, y3, 0, 64, 3, mask=(m, 0), colsum=cs, colsum_groups=1)
    # chain path
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.full((3, n, h, w, 64), float("nan"), dtype=bf, device=DEV)
    E = torch.stack([x, m]).contiguous()
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    cs2 = torch.zeros(64, device=DEV)
    ch.conv(ref(1, 0), ref(0, 0), 0, bs[0], relu=True)
    ch.conv(ref(0, 0), ref(0, 1), 1, bs[1], scale=0.5, res=ref(1, 0))
    ch.conv(ref(0, 1), ref(0, 2), 2, None, mask=ref(1, 1), colsum=cs2, colsum_groups=1)
    ch.run(bank)
    torch.cuda.synchronize()
    if w >= 8:
        assert torch.equal(A[0], y1)
        assert torch.equal(A[1], y2)
        assert torch.equal(A[2], y3)
        assert _rel(cs2, cs) < 1e-5      # same addends, atomics in a different order
    else:
        # W < 8: srb_conv takes the CUDA-core kernel (different fp32 summation order), so the bf16
        # results agree to rounding, not bit for bit
        assert _rel(A[0], y1) < 3e-3 and _rel(A[1], y2) < 3e-3 and _rel(A[2], y3) < 5e-3
        assert _rel(cs2, cs) < 5e-3


def test_chain_long_dependency_chain_bit_exact():
    """24 dependent relu convs on the bench shape (every CTA walks 24 ops x 2 tiles): ordering bugs
    between CTAs (a window read before its producer's store is visible) show up as mismatches."""
    from srb200 import lib as L, ops
    n, h, w, depth = 16, 48, 48, 24
    bf = torch.bfloat16
    x = _rand((n, h, w, 64), seed=3).to(bf)
    ws = [(_rand((64, 64, 3, 3), 0.06, seed=100 + i)).contiguous() for i in range(depth)]
    bs = [_rand((64,), 0.1, seed=200 + i) for i in range(depth)]
    packs = [ops.PackedWeights() for _ in range(depth)]
    cur = x
    outs = []
    for i in range(depth):
        y = torch.empty_like(x)
        ops.conv(cur, 0, 64, packs[i], ws[i], bs[i], y, 0, 64, 3, relu=True)
        outs.append(y)
        cur = y
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.zeros((depth, n, h, w, 64), dtype=bf, device=DEV)
    for rep in range(3):
        A.zero_()
        ch = ops.Chain(n, h, w, x.device)
        ch.space(0, A)
        ch.space(1, x.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        for i in range(depth):
            ch.conv(ref(1, 0) if i == 0 else ref(0, i - 1), ref(0, i), i, bs[i], relu=True)
        ch.run(bank)
        torch.cuda.synchronize()
        for i in range(depth):
            assert torch.equal(A[i], outs[i]), (rep, i)


@pytest.mark.parametrize("shape", [(16, 48, 48), (2, 16, 24), (3, 20, 12)])
def test_chain_ca_forward_and_backward(shape):
    """RCAB tail (conv2 + CALayer + skip) and its backward as chain ops vs srb_conv + srb_ca_fwd /
    srb_ca_bwd.  Tolerance: outputs are bf16 (2^-8 relative rounding); the gate differs by fp32
    summation order only, so elementwise agreement is within one bf16 ulp -> rel L2 < 2e-3."""
    from srb200 import lib as L, ops
    n, h, w = shape
    bf = torch.bfloat16
    cr = 4
    y1 = _rand((n, h, w, 64), seed=5).to(bf)
    x = _rand((n, h, w, 64), seed=6).to(bf)
    g = _rand((n, h, w, 64), 0.01, seed=7).to(bf)
    w2 = _rand((64, 64, 3, 3), 0.05, seed=8).contiguous()
    b2 = _rand((64,), 0.1, seed=9)
    cw1, cb1 = _rand((cr, 64), 0.2, seed=11), _rand((cr,), 0.1, seed=12)
    cw2, cb2 = _rand((64, cr), 0.2, seed=13), _rand((64,), 0.1, seed=14)
    pk = ops.PackedWeights()
    # reference path
    t = torch.empty_like(x)
    pool = torch.zeros(n, 64, device=DEV)
    ops.conv(y1, 0, 64, pk, w2, b2, t, 0, 64, 3, colsum=pool, colsum_groups=n)
    out = torch.empty_like(x)
    s = torch.empty(n, 64, device=DEV)
    yg = torch.empty(n, 64, device=DEV)
    ops.ca_fwd(t, x, pool, False, cw1, cb1, cw2, cb2, out, s, yg)
    dt = torch.empty_like(x)
    gr = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2 = torch.zeros(64, device=DEV)
    ops.ca_bwd(g, t, s, yg, cw1, cb1, cw2, cb2, dt, gr[0], gr[1], gr[2], gr[3], db2, torch.zeros(n, 64, device=DEV),
               accumulate=True, scratch_is_zero=True)
    # chain path
    bank = ops.FilterBank().get([(w2, pk)], L.PACK_FWD)
    A = torch.zeros((3, n, h, w, 64), dtype=bf, device=DEV)
    E = torch.stack([y1, x, g]).contiguous()
    pool2 = torch.zeros(n, 64, device=DEV)
    s2 = torch.empty(n, 64, device=DEV)
    yg2 = torch.empty(n, 64, device=DEV)
    gr2 = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2c = torch.zeros(64, device=DEV)
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    ch.conv_ca(ref(1, 0), ref(0, 0), ref(0, 1), ref(1, 1), 0, b2, pool2, cw1, cb1, cw2, cb2, s2, yg2)
    ch.ca_bwd(ref(0, 0), ref(1, 2), ref(0, 2), cw1, cb1, cw2, cb2, s2, yg2, gr2[0], gr2[1], gr2[2], gr2[3],
              torch.zeros(n, 64, device=DEV), colsum_dt=db2c)
    ch.run(bank)
    torch.cuda.synchronize()
    assert torch.equal(A[0], t)
    assert _rel(pool2, pool) < 1e-5 and _rel(s2, s) < 1e-5 and _rel(yg2, yg) < 1e-5
    assert _rel(A[1], out) < 2e-3
    assert _rel(A[2], dt) < 2e-3
    assert _rel(db2c, db2) < 2e-3
    for a, b in zip(gr2, gr):
        assert _rel(a, b) < 1e-3


@pytest.mark.parametrize("shape", [(16, 48, 48), (3, 20, 12)])
def test_chain_conv_with_fused_ca_backward(shape):
    """dgrad-style conv (+ residual) whose output g is dL/dout of an RCAB, with that RCAB's CALayer
    backward fused into the same op (SRB_CHAIN_CA_BWD_FUSED), versus srb_conv + srb_ca_bwd.  g must be
    bit-identical; dt / gradients within bf16 rounding of the reference kernels (gate summation order)."""
    from srb200 import lib as L, ops
    n, h, w = shape
    bf = torch.bfloat16
    cr = 4
    x = _rand((n, h, w, 64), 0.02, seed=31).to(bf)
    res = _rand((n, h, w, 64), 0.02, seed=32).to(bf)
    t = _rand((n, h, w, 64), seed=33).to(bf)
    wt = _rand((64, 64, 3, 3), 0.05, seed=34).contiguous()
    cw1, cb1 = _rand((cr, 64), 0.2, seed=35), _rand((cr,), 0.1, seed=36)
    cw2, cb2 = _rand((64, cr), 0.2, seed=37), _rand((64,), 0.1, seed=38)
    s = _rand((n, 64), 0.3, seed=39)
    yg = torch.sigmoid(_rand((n, 64), seed=40))
    pk = ops.PackedWeights()
    g = torch.empty_like(x)
    ops.conv(x, 0, 64, pk, wt, None, g, 0, 64, 3, res=(res, 0))
    dt = torch.empty_like(x)
    gr = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2 = torch.zeros(64, device=DEV)
    ops.ca_bwd(g, t, s, yg, cw1, cb1, cw2, cb2, dt, gr[0], gr[1], gr[2], gr[3], db2, torch.zeros(n, 64, device=DEV),
               accumulate=True, scratch_is_zero=True)
    bank = ops.FilterBank().get([(wt, pk)], L.PACK_FWD)
    A = torch.zeros((2, n, h, w, 64), dtype=bf, device=DEV)
    E = torch.stack([x, res, t]).contiguous()
    gr2 = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2c = torch.zeros(64, device=DEV)
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    ch.conv(ref(1, 0), ref(0, 0), 0, None, res=ref(1, 1),
            ca_bwd=dict(t=ref(1, 2), dt=ref(0, 1), w1=cw1, b1=cb1, w2=cw2, b2=cb2, s=s, y=yg, dw1=gr2[0], db1=gr2[1],
                        dw2=gr2[2], db2=gr2[3], scratch=torch.zeros(n, 64, device=DEV), colsum_dt=db2c))
    ch.run(bank)
    torch.cuda.synchronize()
    assert torch.equal(A[0], g)
    assert _rel(A[1], dt) < 2e-3
    assert _rel(db2c, db2) < 2e-3
    for a, b in zip(gr2, gr):
        assert _rel(a, b) < 1e-3


def test_chain_rejects_bad_programs():
    from srb200 import lib as L, ops
    n, h, w = 1, 16, 8
    A = torch.zeros((1, n, h, w, 64), dtype=torch.bfloat16, device=DEV)
    ch = ops.Chain(n, h, w, A.device)
    ch.space(0, A)
    ch.conv(ops.Chain.ref(0, 0), ops.Chain.ref(0, 3), 0)          # slot 3 does not exist
    with pytest.raises(RuntimeError, match="outside the declared spaces"):
        ch.run(torch.zeros(ops.CHAIN_LAYER_BYTES, dtype=torch.uint8, device=DEV))
    ch = ops.Chain(n, h, w, A.device)
    ch.space(0, A)
    ch.conv(ops.Chain.ref(0, 0), ops.Chain.ref(0, 0), 5)          # filter 5 of a 1-layer bank
    with pytest.raises(RuntimeError, match="filter index"):
        ch.run(torch.zeros(ops.CHAIN_LAYER_BYTES, dtype=torch.uint8, device=DEV))


@pytest.mark.parametrize("tile_flags", ["1", "0"], ids=["tile-flags", "sample-counters"])
@pytest.mark.parametrize("cfg", [dict(n_resblocks=2, n_resgroups=2), dict(n_resblocks=20, n_resgroups=1),
                                 dict(n_resblocks=23, n_resgroups=1)])
def test_rcan_chain_path_matches_layer_path(cfg, tile_flags, monkeypatch):
    """Whole model, forward + L1 + backward: the chain path (default) against the per-layer path
    (SRB200_NO_CHAIN=1) on identical weights and inputs.  Both are bf16 pipelines whose conv results
    are bit-identical; CALayer gate differences (fp32 summation order, <= 1 bf16 ulp on an
    activation) then propagate through up to 46 layers -> output rel L2 < 5e-3, global gradient rel
    L2 < 2e-2 (the north_star bf16 bar, which each path also meets against the oracle)."""
    import models
    torch.manual_seed(0)
    kw = dict(n_feats=64, reduction=16, scale_factor=4, **cfg)
    m0 = models.RCAN(**kw)
    sd = {k: v.clone() for k, v in m0.state_dict().items()}
    x = torch.rand(4, 3, 24, 24)
    hr = torch.rand(4, 3, 96, 96)
    res = {}
    monkeypatch.setenv("SRB200_CHAIN_TILEFLAGS", tile_flags)     # both layer-ordering protocols of the chain kernel
    for mode in ("chain", "layers"):
        monkeypatch.setenv("SRB200_NO_CHAIN", "0" if mode == "chain" else "1")
        m = models.RCAN(**kw)
        m.load_state_dict(sd)
        m.compute_dtype = "bf16"
        m = m.to(DEV)
        out = m.training_step({"lr": x.to(DEV), "hr": hr.to(DEV)}, 0)
        out["loss"].backward()
        torch.cuda.synchronize()
        with torch.no_grad():
            sr = m.forward(x.to(DEV)).float().cpu()
        res[mode] = (sr, out["loss"].item(), {k: p.grad.double().cpu() for k, p in m.named_parameters() if p.requires_grad})
    assert _rel(res["chain"][0], res["layers"][0]) < 5e-3
    assert abs(res["chain"][1] - res["layers"][1]) < 1e-3 * abs(res["layers"][1])
    num = sum(float(((res["chain"][2][k] - v) ** 2).sum()) for k, v in res["layers"][2].items())
    den = sum(float((v ** 2).sum()) for v in res["layers"][2].values())
    assert (num / den) ** 0.5 < 2e-2


@pytest.mark.parametrize("cfg", [dict(n_resblocks=16, res_scale=1.0), dict(n_resblocks=3, res_scale=0.1),
                                 dict(n_resblocks=34, res_scale=0.5)])
def test_edsr_chain_path_matches_layer_path(cfg, monkeypatch):
    """EDSR 64-channel body as one chain launch per direction vs the per-layer path.  Every op is a
    plain conv with the same MMA order and epilogue arithmetic, so outputs agree bit for bit and
    gradients to fp32 atomics order (weight gradients come from the same wgrad kernel on identical
    operands; bias gradients are column sums accumulated in a different order)."""
    import models
    torch.manual_seed(0)
    kw = dict(n_feats=64, scale_factor=4, **cfg)
    m0 = models.EDSR(**kw)
    sd = {k: v.clone() for k, v in m0.state_dict().items()}
    x = torch.rand(4, 3, 24, 24)
    hr = torch.rand(4, 3, 96, 96)
    res = {}
    for mode in ("chain", "layers"):
        monkeypatch.setenv("SRB200_NO_CHAIN", "0" if mode == "chain" else "1")
        m = models.EDSR(**kw)
        m.load_state_dict(sd)
        m.compute_dtype = "bf16"
        m = m.to(DEV)
        out = m.training_step({"lr": x.to(DEV), "hr": hr.to(DEV)}, 0)
        out["loss"].backward()
        torch.cuda.synchronize()
        with torch.no_grad():
            sr = m.forward(x.to(DEV)).float().cpu()
        res[mode] = (sr, out["loss"].item(), {k: p.grad.double().cpu() for k, p in m.named_parameters() if p.requires_grad})
    assert torch.equal(res["chain"][0], res["layers"][0])
    assert abs(res["chain"][1] - res["layers"][1]) < 1e-6 * abs(res["layers"][1])   # L1 sum: fp32 atomics order
    for k, v in res["layers"][2].items():
        assert _rel(res["chain"][2][k], v) < 1e-4, k
