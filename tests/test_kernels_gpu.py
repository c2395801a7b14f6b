"""Kernel-level parity (GPU): every libsrb200 kernel, called through the C ABI (ctypes), against
the plain-torch fp64 definition of the same op on the same seeded inputs.

Tolerances: fp32 kernels 1e-5 relative (accumulation-order noise only); bf16 kernels compare
against the fp64 result computed FROM THE SAME bf16-rounded operands, so only the fp32
accumulation and the final bf16 rounding (2^-8 relative) separate them."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _nhwc(t):  # NCHW (cpu, any dtype) -> NHWC contiguous
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _rel(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _conv_case(dtype, n, h, w, cin, cout, k, *, relu=False, scale=1.0, residual=False, mask=False, shuffle=0,
               colsum=0, backend=0, seed=0, x_cs=None, x_co=0, y_cs=None, y_co=0):
    from srb200 import lib as L, ops
    g = torch.Generator().manual_seed(seed)
    xs = cin if x_cs is None else x_cs
    xfull = torch.randn(n, h, w, xs, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    r = shuffle if shuffle > 1 else 1
    cp = cout // (r * r)
    ys = cp if y_cs is None else y_cs
    res_t = torch.randn(n, h * r, w * r, cp, generator=g) if residual else None
    mask_t = torch.randn(n, h * r, w * r, cp, generator=g) if mask else None
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    # operands as the kernel sees them
    xq = xfull.to(tdt)
    wq = wt.to(tdt).float() if (dtype == "bf16" and backend != L.BACKEND_SIMT) else wt
    resq = res_t.to(tdt) if residual else None
    maskq = mask_t.to(tdt) if mask else None
    # ---- fp64 definition
    xin = _nchw(xq[..., x_co:x_co + cin].double())
    ref = F.conv2d(xin, wq.double(), bias.double(), padding=k // 2)
    if relu:
        ref = ref.relu()
    ref = ref * scale
    if r > 1:
        ref = F.pixel_shuffle(ref, r)
    if mask:
        ref = torch.where(_nchw(maskq.double()) > 0, ref, torch.zeros_like(ref))
    if residual:
        ref = ref + _nchw(resq.double())
    ref = _nhwc(ref)
    # ---- kernel
    dev = _dev()
    xd = xq.to(dev)
    y = torch.full((n, h * r, w * r, ys), 7.0, dtype=tdt, device=dev)
    packs = ops.PackedWeights()
    wd = wt.to(dev)
    bd = packs.get_bias(bias.to(dev), shuffle)
    cs = None
    groups = 0
    if colsum:
        groups = n if colsum == 2 else 1
        cs = torch.zeros(groups, cout, dtype=torch.float32, device=dev)
    ops.conv(xd, x_co, cin, packs, wd, bd, y, y_co, cout, k, relu=relu, scale=scale, shuffle=shuffle,
             res=(resq.to(dev), 0) if residual else None, mask=(maskq.to(dev), 0) if mask else None,
             colsum=cs, colsum_groups=groups, backend=backend)
    torch.cuda.synchronize()
    out = y[..., y_co:y_co + cp].float().cpu()
    if ys != cp:  # untouched channels must keep the fill value
        untouched = torch.cat([y[..., :y_co], y[..., y_co + cp:]], dim=-1).float().cpu()
        assert torch.all(untouched == 7.0), "kernel wrote outside its channel slice"
    return out, ref, cs, xd, y


SIMT_CASES = [
    dict(n=2, h=9, w=11, cin=3, cout=64, k=3),
    dict(n=1, h=16, w=16, cin=64, cout=64, k=3, relu=True),
    dict(n=2, h=8, w=8, cin=64, cout=3, k=3),
    dict(n=1, h=12, w=10, cin=64, cout=64, k=3, scale=0.1, residual=True),
    dict(n=1, h=8, w=8, cin=32, cout=32, k=3, mask=True, residual=True),
    dict(n=1, h=8, w=8, cin=16, cout=36, k=3, shuffle=3),
    dict(n=2, h=8, w=8, cin=16, cout=32, k=3, shuffle=2, residual=True),
    dict(n=1, h=10, w=10, cin=3, cout=64, k=9, relu=True),
    dict(n=1, h=10, w=10, cin=32, cout=3, k=5),
    dict(n=1, h=7, w=9, cin=96, cout=64, k=1, residual=True),
    dict(n=2, h=8, w=8, cin=64, cout=64, k=3, colsum=2),
    dict(n=2, h=8, w=8, cin=64, cout=64, k=3, colsum=1, mask=True),
    dict(n=1, h=8, w=8, cin=64, cout=32, k=3, x_cs=160, x_co=32, y_cs=160, y_co=96, relu=True),
]


@pytest.mark.parametrize("case", SIMT_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_conv_simt(case, dtype):
    from srb200 import lib as L
    out, ref, cs, _, _ = _conv_case(dtype, backend=L.BACKEND_SIMT, **case)
    l2, mx = _rel(out, ref)
    tol = 2e-5 if dtype == "fp32" else 6e-3
    assert l2 < tol and mx < 4 * tol, (l2, mx)
    if cs is not None:
        want = ref.sum(dim=(1, 2)) if cs.shape[0] > 1 else ref.sum(dim=(0, 1, 2))[None]
        l2c, _ = _rel(cs, want)
        assert l2c < (1e-4 if dtype == "fp32" else 2e-2), l2c


UMMA_CASES = [
    dict(n=1, h=16, w=8, cin=64, cout=64, k=3),                      # exactly one tile
    dict(n=2, h=48, w=48, cin=64, cout=64, k=3, relu=True),          # the RCAN/EDSR hot shape
    dict(n=1, h=20, w=13, cin=64, cout=64, k=3),                     # ragged tiles (H, W not multiples)
    dict(n=2, h=16, w=16, cin=64, cout=64, k=3, scale=0.1, residual=True),
    dict(n=1, h=16, w=16, cin=64, cout=64, k=3, mask=True, residual=True),
    dict(n=2, h=16, w=16, cin=64, cout=64, k=3, colsum=2),
    dict(n=2, h=16, w=24, cin=64, cout=64, k=3, colsum=1, mask=True),
    dict(n=1, h=16, w=16, cin=64, cout=256, k=3, shuffle=2),         # UpscaleBlock conv + PixelShuffle
    dict(n=1, h=16, w=16, cin=256, cout=256, k=3, relu=True),        # EDSR-large (BN=128, 4 K chunks)
    dict(n=1, h=8, w=16, cin=128, cout=576, k=3, shuffle=3),         # x3: 9 * 64
    dict(n=1, h=16, w=16, cin=576, cout=64, k=1, residual=True),     # RDN LFF
    dict(n=1, h=16, w=16, cin=1024, cout=64, k=1),                   # RDN GFF.0
    dict(n=1, h=16, w=16, cin=128, cout=64, k=3, x_cs=576, x_co=0, y_cs=576, y_co=128, relu=True),  # RDN dense
    dict(n=1, h=16, w=16, cin=64, cout=192, k=3, x_cs=576, x_co=192, y_cs=576, y_co=0, residual=False),
    dict(n=1, h=16, w=16, cin=64, cout=32, k=3),                     # BN=32
    dict(n=1, h=5, w=40, cin=64, cout=64, k=3),                      # short & wide -> 16x8 tiles
    dict(n=2, h=16, w=16, cin=3, cout=64, k=3, x_cs=8),              # head conv: partial K chunk, padded stride
    dict(n=1, h=32, w=24, cin=64, cout=3, k=3, y_cs=8),              # tail conv: narrow output (N tile 16)
    dict(n=1, h=32, w=24, cin=64, cout=3, k=3, y_cs=8, relu=True),
    dict(n=1, h=16, w=16, cin=96, cout=32, k=3, x_cs=256, y_cs=256, y_co=96, relu=True),   # RDN-A dense layer
    dict(n=3, h=48, w=48, cin=64, cout=64, k=3, residual=True),      # persistent kernel, >1 tile per CTA? (54 tiles)
    dict(n=16, h=48, w=48, cin=64, cout=64, k=3, relu=True),         # 288 tiles on 148 SMs: 2 tiles per CTA
    dict(n=16, h=48, w=48, cin=64, cout=64, k=3, colsum=2, mask=True),
    dict(n=7, h=50, w=45, cin=64, cout=64, k=3, scale=0.5, residual=True),   # ragged, 3+ tiles per CTA
    # wide-layer kernel (conv_wide.cu): two pixel tiles per filter stage, persistent, double-buffered TMEM
    dict(n=1, h=40, w=72, cin=256, cout=256, k=3, scale=0.1, residual=True),       # EDSR-large ResBlock conv2, >148 items
    dict(n=2, h=33, w=29, cin=256, cout=256, k=3, relu=True),                        # ragged pairs (W not a multiple of 16)
    dict(n=1, h=16, w=32, cin=256, cout=1024, k=3, shuffle=2),                       # EDSR-large up-sampling conv
    dict(n=1, h=17, w=16, cin=64, cout=128, k=3, mask=True),                         # one K chunk, masked (dgrad form)
    dict(n=3, h=48, w=48, cin=128, cout=128, k=3, x_cs=576, x_co=64, y_cs=576, y_co=192, relu=True),  # channel slices
]


@pytest.mark.parametrize("case", UMMA_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_umma(case):
    from srb200 import lib as L
    out, ref, cs, _, _ = _conv_case("bf16", backend=L.BACKEND_UMMA, **case)
    l2, mx = _rel(out, ref)
    assert l2 < 6e-3 and mx < 2e-2, (l2, mx)
    if cs is not None:
        want = ref.sum(dim=(1, 2)) if cs.shape[0] > 1 else ref.sum(dim=(0, 1, 2))[None]
        l2c, _ = _rel(cs, want)
        assert l2c < 2e-2, l2c


@pytest.mark.parametrize("shuffle", [0, 2])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_conv_dgrad_matches_autograd(dtype, shuffle):
    """DGRAD packing: the forward kernel with rotated/swapped weights gives autograd's input grad."""
    from srb200 import lib as L, ops
    g = torch.Generator().manual_seed(1)
    n, h, w, cin, cout, k = 2, 16, 16, 64, 64 if not shuffle else 256, 3
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * 9) ** 0.5
    r = shuffle if shuffle else 1
    gy = torch.randn(n, h * r, w * r, cout // (r * r), generator=g).to(tdt)
    x = torch.zeros(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    wq = wt.to(tdt).double() if dtype == "bf16" else wt.double()
    y = F.conv2d(x, wq, padding=1)
    if shuffle:
        y = F.pixel_shuffle(y, r)
    y.backward(_nchw(gy.double()))
    ref = _nhwc(x.grad)
    dev = _dev()
    packs = ops.PackedWeights()
    dx = torch.empty(n, h, w, cin, dtype=tdt, device=dev)
    gyd = gy.to(dev)
    if shuffle:
        gu = ops.pixel_unshuffle(gyd, r)
        ops.conv_dgrad_shuffled(gu, packs, wt.to(dev), dx, r)
    else:
        ops.conv(gyd, 0, cout, packs, wt.to(dev), None, dx, 0, cin, k, mode=L.PACK_DGRAD)
    torch.cuda.synchronize()
    l2, mx = _rel(dx.float(), ref)
    assert l2 < (2e-5 if dtype == "fp32" else 6e-3), (l2, mx)


WGRAD_CASES = [
    dict(n=2, h=16, w=16, cin=64, cout=64, k=3),
    dict(n=1, h=9, w=11, cin=3, cout=64, k=3),
    dict(n=2, h=12, w=12, cin=64, cout=3, k=3),
    dict(n=1, h=10, w=10, cin=3, cout=64, k=9),
    dict(n=1, h=8, w=8, cin=96, cout=64, k=1),
    dict(n=1, h=8, w=8, cin=64, cout=256, k=3, shuffle=2),
    dict(n=1, h=8, w=8, cin=64, cout=64, k=3, alpha=0.1, accumulate=True),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_conv_wgrad(case, dtype):
    from srb200 import ops
    c = dict(case)
    n, h, w, cin, cout, k = c["n"], c["h"], c["w"], c["cin"], c["cout"], c["k"]
    shuffle, alpha, accumulate = c.get("shuffle", 0), c.get("alpha", 1.0), c.get("accumulate", False)
    g = torch.Generator().manual_seed(2)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    x = torch.randn(n, h, w, cin, generator=g).to(tdt)
    r = shuffle if shuffle else 1
    gy = torch.randn(n, h * r, w * r, cout // (r * r), generator=g).to(tdt)
    wt = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(_nchw(x.double()), wt, bt, padding=k // 2)
    if shuffle:
        y = F.pixel_shuffle(y, r)
    y.backward(_nchw(gy.double()))
    dev = _dev()
    base = torch.randn(cout, cin, k, k, generator=g)
    dw = base.to(dev).clone() if accumulate else torch.full((cout, cin, k, k), 3.0, device=dev)
    db = torch.zeros(cout, device=dev) if accumulate else torch.full((cout,), 3.0, device=dev)
    gyd = gy.to(dev)
    if shuffle:
        gyd = ops.pixel_unshuffle(gyd, r)
    ops.conv_wgrad(x.to(dev), 0, cin, gyd, 0, cout, k, dw, db, accumulate=accumulate, shuffle=shuffle, alpha=alpha)
    torch.cuda.synchronize()
    want_w = wt.grad * alpha + (base.double() if accumulate else 0)
    want_b = bt.grad * alpha
    l2w, _ = _rel(dw, want_w)
    l2b, _ = _rel(db, want_b)
    assert l2w < 2e-5 and l2b < 2e-5, (l2w, l2b)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("compute_pool", [True, False])
def test_channel_attention_fwd_bwd(dtype, compute_pool):
    from srb200 import ops
    g = torch.Generator().manual_seed(3)
    n, h, w, c, cr = 3, 12, 10, 64, 4
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    t = torch.randn(n, h, w, c, generator=g).to(tdt)
    skip = torch.randn(n, h, w, c, generator=g).to(tdt)
    w1 = torch.randn(cr, c, generator=g) * 0.3
    b1 = torch.randn(cr, generator=g) * 0.1
    w2 = torch.randn(c, cr, generator=g) * 0.3
    b2 = torch.randn(c, generator=g) * 0.1
    gout = torch.randn(n, h, w, c, generator=g).to(tdt)
    # fp64 definition (rcan.py:23-29, :54)
    td = t.double().requires_grad_(True)
    prm = [p.double().requires_grad_(True) for p in (w1, b1, w2, b2)]
    s = td.mean(dim=(1, 2))
    z = (s @ prm[0].T + prm[1]).relu()
    yg = torch.sigmoid(z @ prm[2].T + prm[3])
    out_ref = td * yg[:, None, None, :] + skip.double()
    out_ref.backward(gout.double())
    dev = _dev()
    td_, sk_ = t.to(dev), skip.to(dev)
    pool = torch.zeros(n, c, device=dev)
    if not compute_pool:
        pool = t.float().sum(dim=(1, 2)).to(dev)
    out = torch.empty_like(td_)
    s_out = torch.empty(n, c, device=dev)
    y_out = torch.empty(n, c, device=dev)
    pw = [p.to(dev).contiguous() for p in (w1, b1, w2, b2)]
    ops.ca_fwd(td_, sk_, pool, compute_pool, *pw, out, s_out, y_out)
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == "fp32" else 6e-3
    assert _rel(out.float(), out_ref.detach())[0] < tol
    assert _rel(s_out, s.detach())[0] < 1e-5
    assert _rel(y_out, yg.detach())[0] < 1e-5
    dt = torch.empty_like(td_)
    dws = [torch.full_like(p, 5.0) for p in pw]
    colsum_dt = torch.full((c,), 5.0, device=dev)
    scratch = torch.empty(n, c, device=dev)
    ops.ca_bwd(gout.to(dev), td_, s_out, y_out, *pw, dt, dws[0], dws[1], dws[2], dws[3], colsum_dt, scratch)
    torch.cuda.synchronize()
    assert _rel(dt.float(), td.grad)[0] < tol
    gtol = 1e-4 if dtype == "fp32" else 1e-2
    for got, want in zip(dws, prm):
        assert _rel(got, want.grad)[0] < gtol
    assert _rel(colsum_dt, dt.float().sum(dim=(0, 1, 2)))[0] < 1e-4


def test_layout_and_elementwise():
    from srb200 import ops
    g = torch.Generator().manual_seed(4)
    dev = _dev()
    x = torch.rand(2, 3, 7, 9, generator=g)
    add = torch.tensor([-0.4488, -0.4371, -0.4040])
    for dt in (torch.float32, torch.bfloat16):
        y = ops.nchw_to_nhwc(x.to(dev), add.to(dev), dt)
        want = (x + add[None, :, None, None]).permute(0, 2, 3, 1)
        assert torch.equal(y.cpu(), want.to(dt))
        back = ops.nhwc_to_nchw(y, 0, 3, (-add).to(dev))
        assert _rel(back, y.float().cpu().permute(0, 3, 1, 2) - add[None, :, None, None])[1] < 1e-6
    for dt in (torch.float32, torch.bfloat16):
        a = torch.randn(2, 5, 6, 24, generator=g).to(dt).to(dev)
        b = torch.randn(2, 5, 6, 40, generator=g).to(dt).to(dev)
        out = torch.zeros(2, 5, 6, 16, dtype=dt, device=dev)
        ops.add_channels(a, 8, b, 24, out, 4, 8)
        want = (a[..., 8:16].float() + b[..., 24:32].float()).to(dt)
        assert torch.equal(out[..., 4:12], want) and torch.all(out[..., :4] == 0) and torch.all(out[..., 12:] == 0)
        ops.copy_channels(b, 3, out, 1, 5)   # unaligned -> scalar path
        assert torch.equal(out[..., 1:6], b[..., 3:8])
        m = torch.randn(2, 5, 6, 24, generator=g).to(dt).to(dev)
        o2 = torch.empty_like(a)
        ops.relu_bwd(a, 0, m, 0, o2, 0, 24)
        assert torch.equal(o2, torch.where(m > 0, a, torch.zeros_like(a)))
        # channel slices, in place (the RDN dense-block backward masks one 64-channel slice per layer);
        # zeros and negative zeros in the activation both mask
        gd = torch.randn(2, 5, 6, 40, generator=g).to(dt).to(dev)
        act = torch.randn(2, 5, 6, 40, generator=g).to(dt).to(dev)
        act[0, 0, 0, 8:12] = 0.0
        act[0, 0, 1, 8:12] = -0.0
        keep = gd.clone()
        ops.relu_bwd(gd, 8, act, 8, gd, 8, 16)
        assert torch.equal(gd[..., 8:24], torch.where(act[..., 8:24] > 0, keep[..., 8:24], torch.zeros_like(keep[..., 8:24])))
        assert torch.equal(gd[..., :8], keep[..., :8]) and torch.equal(gd[..., 24:], keep[..., 24:])
        gsh = torch.randn(2, 8, 12, 5, generator=g).to(dt).to(dev)
        un = ops.pixel_unshuffle(gsh, 2)
        ref = F.pixel_unshuffle(gsh.permute(0, 3, 1, 2).float(), 2)         # channels (c', i, j)
        ref = ref.reshape(2, 5, 4, 4, 6).permute(0, 3, 4, 2, 1).reshape(2, 4, 6, 20)  # -> (ij, c')
        assert torch.equal(un.float(), ref)
        cs = torch.full((24,), 9.0, device=dev)
        ops.colsum(a, 0, 24, cs)
        assert _rel(cs, a.float().sum(dim=(0, 1, 2)))[0] < 1e-5


def test_l1_loss_and_adam():
    from srb200 import ops
    g = torch.Generator().manual_seed(5)
    dev = _dev()
    sr = torch.rand(2, 3, 17, 19, generator=g)
    hr = torch.rand(2, 3, 17, 19, generator=g)
    hr[0, 0, 0, :5] = sr[0, 0, 0, :5]     # exact ties -> zero gradient
    srd = sr.double().requires_grad_(True)
    ref = F.l1_loss(srd, hr.double())
    ref.backward()
    loss, grad = ops.l1_loss(sr.to(dev), hr.to(dev))
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) < 1e-6
    assert _rel(grad, srd.grad)[1] < 1e-6
    # Adam against torch.optim.Adam for 3 steps: 16-byte lanes (n = 1000), lanes + scalar tail (1003), and a parameter
    # slice that is not 16-byte aligned (all-scalar path)
    for n, off in ((1000, 0), (1003, 0), (1001, 1)):
        p0 = torch.randn(n, generator=g)
        p_ref = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([p_ref], lr=1e-3)
        p = torch.zeros(n + 4, device=dev)[off:off + n]
        p.copy_(p0)
        m = torch.zeros_like(p)
        v = torch.zeros_like(p)
        step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        for step in range(1, 4):
            gr = torch.randn(n, generator=g)
            p_ref.grad = gr.clone()
            opt.step()
            ops.inc_counter(step_dev)
            ops.adam_step(p, (gr * 4).to(dev), m, v, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
                          step=0, step_dev=step_dev, grad_scale=0.25)
        torch.cuda.synchronize()
        assert _rel(p, p_ref.detach())[1] < 1e-6, (n, off)


def test_errors_are_reported_not_thrown():
    """C ABI error contract: non-zero status + srb_last_error text -> RuntimeError in the wrapper."""
    from srb200 import lib as L, ops
    dev = _dev()
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=dev)
    y = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(64, 64, 4, 4, device=dev)
    with pytest.raises(RuntimeError, match="kernel size"):
        ops.conv(x, 0, 64, ops.PackedWeights(), w, None, y, 0, 64, 4)
    xf = torch.zeros(1, 8, 8, 64, device=dev)
    w3 = torch.zeros(64, 64, 3, 3, device=dev)
    with pytest.raises(RuntimeError, match="not eligible"):        # the tcgen05 path is bf16-only
        ops.conv(xf, 0, 64, ops.PackedWeights(), w3, None, torch.zeros_like(xf), 0, 64, 3, backend=L.BACKEND_UMMA)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.nchw_to_nhwc(torch.zeros(1, 3, 4, 4), None, torch.float32)


WGRAD_UMMA_CASES = [
    dict(n=1, h=16, w=8, cin=64, cout=64),                         # one tile
    dict(n=2, h=48, w=48, cin=64, cout=64),                        # the hot RCAN/EDSR shape
    dict(n=1, h=20, w=13, cin=64, cout=64),                        # ragged tiles
    dict(n=1, h=16, w=16, cin=128, cout=192),                      # several 64x64 blocks per layer
    dict(n=1, h=16, w=16, cin=64, cout=256, shuffle=2),            # un-shuffled gradient channel order
    dict(n=2, h=16, w=16, cin=64, cout=64, alpha=0.1, accumulate=True),
    dict(n=1, h=16, w=16, cin=64, cout=64, x_cs=192, x_co=64, g_cs=128, g_co=64),   # channel slices
    dict(n=2, h=16, w=16, cin=3, cout=64, x_cs=8),                 # head conv (partial ci block)
    dict(n=1, h=32, w=24, cin=64, cout=3, g_cs=8),                 # tail conv (partial co block)
    dict(n=1, h=16, w=16, cin=96, cout=32, x_cs=256, g_cs=256, g_co=96),   # RDN-A dense layer
    dict(n=2, h=48, w=48, cin=576, cout=64, k=1),                  # RDN LFF (1x1: centre tap only)
    dict(n=1, h=20, w=13, cin=1024, cout=64, k=1),                 # RDN GFF.0, ragged tiles
    dict(n=1, h=16, w=16, cin=64, cout=64, k=1, x_cs=192, x_co=64, alpha=0.5, accumulate=True),
]


@pytest.mark.parametrize("case", WGRAD_UMMA_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_wgrad_umma(case):
    """tcgen05 weight gradient (MN-major operands, taps stacked along M) vs autograd in fp64 on
    the same bf16-rounded operands; bias gradient through the fused column-sum launch."""
    from srb200 import lib as L, ops
    c = dict(case)
    n, h, w, cin, cout = c["n"], c["h"], c["w"], c["cin"], c["cout"]
    shuffle, alpha, accumulate = c.get("shuffle", 0), c.get("alpha", 1.0), c.get("accumulate", False)
    x_cs, x_co = c.get("x_cs", cin), c.get("x_co", 0)
    g_cs, g_co = c.get("g_cs", cout), c.get("g_co", 0)
    k = c.get("k", 3)
    g = torch.Generator().manual_seed(6)
    xfull = torch.randn(n, h, w, x_cs, generator=g).to(torch.bfloat16)
    r = shuffle if shuffle else 1
    gy = torch.randn(n, h * r, w * r, cout // (r * r), generator=g).to(torch.bfloat16)
    wt = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(_nchw(xfull[..., x_co:x_co + cin].double()), wt, bt, padding=k // 2)
    if shuffle:
        y = F.pixel_shuffle(y, r)
    y.backward(_nchw(gy.double()))
    dev = _dev()
    base = torch.randn(cout, cin, k, k, generator=g)
    dw = base.to(dev).clone() if accumulate else torch.full((cout, cin, k, k), 3.0, device=dev)
    db = torch.zeros(cout, device=dev) if accumulate else torch.full((cout,), 3.0, device=dev)
    gyd = gy.to(dev)
    if shuffle:
        gyd = ops.pixel_unshuffle(gyd, r)
    if g_cs != cout:
        gfull = torch.randn(n, h, w, g_cs, generator=g).to(torch.bfloat16).to(dev)
        gfull[..., g_co:g_co + cout] = gyd
        gyd = gfull
    ops.conv_wgrad(xfull.to(dev), x_co, cin, gyd, g_co, cout, k, dw, db, accumulate=accumulate, shuffle=shuffle,
                   alpha=alpha, backend=L.BACKEND_UMMA)
    torch.cuda.synchronize()
    want_w = wt.grad * alpha + (base.double() if accumulate else 0)
    l2w, mxw = _rel(dw, want_w)
    l2b, _ = _rel(db, bt.grad * alpha)
    assert l2w < 2e-5 and l2b < 2e-5, (l2w, mxw, l2b)


def test_conv_wgrad_batched_deferred():
    """Several layers through ops.deferred_wgrads(): one srb_conv_wgrad_batched call, results equal
    to the one-by-one launches (tcgen05-eligible and CUDA-core layers mixed)."""
    from srb200 import ops
    g = torch.Generator().manual_seed(7)
    dev = _dev()
    layers = []
    for i, (cin, cout, hw) in enumerate([(64, 64, 48), (64, 64, 48), (3, 64, 24), (64, 128, 16), (64, 3, 32), (64, 64, 48)]):
        x = torch.randn(2, hw, hw, cin, generator=g).to(torch.bfloat16).to(dev)
        gy = torch.randn(2, hw, hw, cout, generator=g).to(torch.bfloat16).to(dev)
        layers.append((x, gy, cin, cout))
    ref = []
    for x, gy, cin, cout in layers:
        dw = torch.empty(cout, cin, 3, 3, device=dev)
        db = torch.empty(cout, device=dev)
        ops.conv_wgrad(x, 0, cin, gy, 0, cout, 3, dw, db)
        ref.append((dw, db))
    got = []
    with ops.deferred_wgrads():
        for x, gy, cin, cout in layers:
            dw = torch.full((cout, cin, 3, 3), 9.0, device=dev)
            db = torch.full((cout,), 9.0, device=dev)
            ops.conv_wgrad(x, 0, cin, gy, 0, cout, 3, dw, db)
            got.append((dw, db))
    torch.cuda.synchronize()
    for (dw, db), (rw, rb) in zip(got, ref):
        assert _rel(dw, rw)[0] < 1e-5 and _rel(db, rb)[0] < 1e-5


def test_deferred_wgrads_mixed_1x1_and_3x3_batch():
    """A deferred batch that mixes 3x3 and 1x1 layers (RDN: dense layers + LFF / GFF, rdn.py:24-40,59-72) is split by kernel
    size — single-window N = 128 launches for the 3x3 layers, three-window launches for the 1x1 ones — and must give the
    gradients of the per-layer calls and of torch's conv2d weight gradient."""
    from srb200 import ops
    g = torch.Generator().manual_seed(12)
    dev = _dev()
    layers = []
    for cin, cout, k, hw in [(64, 64, 3, 32), (192, 64, 1, 32), (128, 64, 3, 32), (64, 64, 1, 24), (64, 64, 3, 24)]:
        x = torch.randn(2, hw, hw, cin, generator=g).to(torch.bfloat16).to(dev)
        gy = torch.randn(2, hw, hw, cout, generator=g).to(torch.bfloat16).to(dev)
        layers.append((x, gy, cin, cout, k))
    got = []
    with ops.deferred_wgrads():
        for x, gy, cin, cout, k in layers:
            dw = torch.full((cout, cin, k, k), 9.0, device=dev)
            ops.conv_wgrad(x, 0, cin, gy, 0, cout, k, dw, None)
            got.append(dw)
    torch.cuda.synchronize()
    for (x, gy, cin, cout, k), dw in zip(layers, got):
        want = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (cout, cin, k, k), gy.permute(0, 3, 1, 2).double(),
                                           padding=k // 2)
        assert _rel(dw, want)[0] < 1e-5, (cin, cout, k)


def test_deferred_bias_grads_share_one_launch():
    """Bias gradients of a deferred batch go through colsum_batched_kernel (one launch): channel
    slices, the un-shuffled (i,j,c') channel order of PixelShuffle convs, accumulate and alpha,
    checked against torch sums of the same bf16 tensors."""
    from srb200 import lib as L, ops
    g = torch.Generator().manual_seed(11)
    dev = _dev()
    cases = [  # (cout, g_cs, g_co, hw, shuffle, accumulate, alpha)
        (64, 64, 0, 48, 0, False, 1.0),
        (64, 192, 64, 24, 0, True, 0.5),
        (256, 256, 0, 16, 2, False, 1.0),
        (128, 128, 0, 20, 0, True, 1.0),
        (32, 256, 96, 16, 0, False, 1.0),
        (3, 8, 0, 32, 0, False, 1.0),          # tail conv: RGB gradient stored 8 wide
    ]
    keep, want, got = [], [], []
    c0 = L.launch_count()
    with ops.deferred_wgrads():
        for cout, g_cs, g_co, hw, shuffle, acc, alpha in cases:
            x = torch.randn(2, hw, hw, 64, generator=g).to(torch.bfloat16).to(dev)
            gy = torch.randn(2, hw, hw, g_cs, generator=g).to(torch.bfloat16).to(dev)
            dw = torch.zeros(cout, 64, 3, 3, device=dev)
            base = torch.randn(cout, generator=g).to(dev)
            db = base.clone() if acc else torch.full((cout,), 7.0, device=dev)
            ops.conv_wgrad(x, 0, 64, gy, g_co, cout, 3, dw, db, accumulate=acc, shuffle=shuffle, alpha=alpha)
            sums = gy[..., g_co:g_co + cout].double().sum(dim=(0, 1, 2))
            if shuffle:      # stored order (i*r+j)*C' + c'  ->  parameter order c'*r*r + (i*r+j)
                rr = shuffle * shuffle
                sums = sums.view(rr, cout // rr).t().reshape(-1)
            want.append(sums * alpha + (base.double() if acc else 0))
            got.append(db)
            keep.append((x, gy, dw))
    torch.cuda.synchronize()
    n_launch = L.launch_count() - c0
    for w_, g_ in zip(want, got):
        assert _rel(g_, w_)[0] < 1e-5
    assert n_launch <= 3, n_launch     # one column-sum launch + the batched weight-gradient launch(es)


@pytest.mark.parametrize("shape", [(64, 64, 3), (256, 64, 3), (64, 576, 1), (64, 40, 3), (3, 64, 3), (64, 3, 3), (128, 192, 3), (256, 256, 3)])
@pytest.mark.parametrize("grid_elems", [None, 2048])
def test_pack_table_matches_pack_weight(shape, grid_elems):
    """srb_pack_table (one launch re-packing many weights, used once per optimizer step) must write
    exactly the bytes srb_pack_weight writes, for every packing / mode / shuffle, including the
    zero-padded partial 64-channel chunk."""
    import ctypes as C
    import numpy as np
    from srb200 import lib as L, ops
    cout, cin, k = shape
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7)
    w = torch.randn(cout, cin, k, k, generator=g).to(dev)
    rows, keep = [], []
    for packing in (L.PACK_SIMT, L.PACK_UMMA):
        for mode in (L.PACK_FWD, L.PACK_DGRAD):
            for shuffle in ((0, 2) if cout % 4 == 0 else (0,)):
                want = ops.pack_weight(w, packing, mode, shuffle)
                got = torch.ones_like(want)        # every byte is rewritten, the zero padding of a partial chunk included
                rows.append((w.data_ptr(), got.data_ptr(), cout, cin, k, packing, mode, shuffle))
                keep.append((want, got, packing, mode, shuffle))
    dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("ksize", "<i4"),
                   ("packing", "<i4"), ("mode", "<i4"), ("shuffle", "<i4")])
    table = torch.from_numpy(np.array(rows, dtype=dt).view(np.uint8).copy()).to(dev)
    # grid_elems = 2048: one CTA per item, every path loops over its whole item
    L.check(L.load().srb_pack_table(C.c_void_p(L.ctx(0)), C.c_void_p(table.data_ptr()), len(rows), grid_elems or w.numel(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "srb_pack_table")
    torch.cuda.synchronize()
    for want, got, packing, mode, shuffle in keep:
        assert torch.equal(want, got), (packing, mode, shuffle)
