"""Differentiable operators of the SR hot path: `torch.autograd.Function`s whose forward AND
backward are libsrb200 kernel launches (the reference has no custom Functions — its backward is
autograd over ATen ops; here every adjoint is hand-written).

Gradient delivery for parameters.  By default a Function returns parameter gradients to autograd
(drop-in with any torch optimizer; with `zero_grad(set_to_none=True)` AccumulateGrad adopts the
tensor without a copy).  When a parameter carries a `_srb_grad` attribute (a view into a flat
fp32 gradient buffer installed by `srb200.trainer.FlatParams`), the kernels write straight into
that view and `None` is returned — one contiguous buffer for the NCCL all-reduce and the fused
Adam, no per-parameter launches.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import lib as L
from . import ops
from .ops import PackedWeights


def _grad_target(p: torch.Tensor):
    """(buffer to write the gradient into, accumulate?, value to return to autograd)."""
    flat = getattr(p, "_srb_grad", None)
    if flat is not None:
        acc = bool(getattr(p, "_srb_grad_live", False))
        p._srb_grad_live = True
        return flat, acc, None
    buf = torch.empty_like(p, memory_format=torch.contiguous_format)
    return buf, False, buf


def _padded(c: int, dtype) -> int:
    """Channel stride used for a c-channel NHWC tensor: bf16 tensors keep pixels 16-byte aligned
    (multiple of 8 channels) so that TMA can address them — the 3-channel RGB tensors at the model
    boundary are stored 8 wide; the extra channels are never read (TMA clips at the logical extent)."""
    if dtype == torch.bfloat16 and c % 8:
        return (c + 7) // 8 * 8
    return c


class ToNHWC(Function):
    """NCHW fp32 -> NHWC compute dtype, with MeanShift (common.py:58-71) as a per-channel add."""

    @staticmethod
    def forward(ctx, x, chan_add, dtype):
        n, c, h, w = x.shape
        ctx.c = c
        out = torch.empty((n, h, w, _padded(c, dtype)), dtype=dtype, device=x.device)
        return ops.nchw_to_nhwc(x.contiguous().float(), chan_add, dtype, out=out)

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(g.contiguous(), 0, ctx.c, None), None, None


class ToNCHW(Function):
    """NHWC compute dtype -> NCHW fp32, with add_mean (edsr.py:52, rcan.py:127) as a per-channel add."""

    @staticmethod
    def forward(ctx, y, chan_add, channels):
        ctx.dtype, ctx.cs = y.dtype, y.shape[3]
        return ops.nhwc_to_nchw(y.contiguous(), 0, channels, chan_add)

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = g.shape
        out = torch.empty((n, h, w, ctx.cs), dtype=ctx.dtype, device=g.device)
        return ops.nchw_to_nhwc(g.contiguous().float(), None, ctx.dtype, out=out), None, None


class ConvFn(Function):
    """y = [PixelShuffle_r]( relu?(conv_k(x) + b) * scale ) + residual?

    Covers DefaultConv2d / nn.Conv2d call sites (common.py:7-30; edsr.py:21-33; rcan.py:68,101;
    rdn.py:57-59,71-72,87-93) together with the elementwise ops that follow them.  The logical
    channel counts come from the weight; tensors may be wider (padded channel stride)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, packs: PackedWeights, relu: bool, scale: float, shuffle: int):
        assert not (relu and (residual is not None or scale != 1.0)), "unsupported epilogue combination"
        x = x.contiguous()
        n, h, w, _ = x.shape
        cout, cin, k, _ = weight.shape
        r = shuffle if shuffle > 1 else 1
        cp = cout // (r * r)
        y = torch.empty((n, h * r, w * r, _padded(cp, x.dtype)), dtype=x.dtype, device=x.device)
        b = packs.get_bias(bias, shuffle) if bias is not None else None
        res = (residual.contiguous(), 0) if residual is not None else None
        ops.conv(x, 0, cin, packs, weight, b, y, 0, cout, k, relu=relu, scale=scale, shuffle=shuffle, res=res)
        ctx.save_for_backward(x, weight, bias, y if relu else None)
        ctx.packs, ctx.relu, ctx.scale, ctx.shuffle = packs, relu, scale, shuffle
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, bias, y = ctx.saved_tensors
        packs, relu, scale, shuffle = ctx.packs, ctx.relu, ctx.scale, ctx.shuffle
        g = g.contiguous()
        cout, cin, k, _ = weight.shape
        gm = g
        if relu:
            gm = torch.empty_like(g)
            ops.relu_bwd(g, 0, y, 0, gm, 0, cout)
        if shuffle > 1:
            gm = ops.pixel_unshuffle(gm, shuffle, cout // (shuffle * shuffle))   # [N,H,W,Cout] in (ij, c') channel order
        dw = db = dx = None
        if ctx.needs_input_grad[1]:
            wbuf, wacc, dw = _grad_target(weight)
            bbuf = None
            if bias is not None and ctx.needs_input_grad[2]:
                bbuf, bacc, db = _grad_target(bias)
                if bacc != wacc:
                    # the weight is a derived tensor (weight-norm: its gradient goes back to autograd in a fresh buffer) while the
                    # bias lives in the flat gradient buffer: the two targets accumulate differently, so the bias gets its own item
                    ops.bias_grad(gm, 0, cout, bbuf, accumulate=bacc, alpha=scale, shuffle=shuffle)
                    bbuf = None
            # a gradient handed back to autograd (dw is not None) must exist when this function returns
            ops.conv_wgrad(x, 0, cin, gm, 0, cout, k, wbuf, bbuf, accumulate=wacc, shuffle=shuffle, alpha=scale, defer=dw is None)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if shuffle > 1:
                assert scale == 1.0
                ops.conv_dgrad_shuffled(gm, packs, weight, dx, shuffle)
            else:
                ops.conv(gm, 0, cout, packs, weight, None, dx, 0, cin, k, mode=L.PACK_DGRAD, scale=scale)
        dres = g if ctx.has_res else None
        return dx, dw, db, dres, None, None, None, None


class ResBlockFn(Function):
    """EDSR ResBlock (common.py:74-109): out = conv2(relu(conv1(x))) * res_scale + x."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, p1: PackedWeights, p2: PackedWeights, scale: float):
        x = x.contiguous()
        c = x.shape[3]
        y1 = torch.empty_like(x)
        ops.conv(x, 0, c, p1, w1, b1.detach(), y1, 0, c, 3, relu=True)
        out = torch.empty_like(x)
        ops.conv(y1, 0, c, p2, w2, b2.detach(), out, 0, c, 3, scale=scale, res=(x, 0))
        ctx.save_for_backward(x, y1, w1, b1, w2, b2)
        ctx.p1, ctx.p2, ctx.scale = p1, p2, scale
        return out

    @staticmethod
    def backward(ctx, g):
        x, y1, w1, b1, w2, b2 = ctx.saved_tensors
        g = g.contiguous()
        c = x.shape[3]
        w2buf, acc2, dw2 = _grad_target(w2)
        b2buf, _, db2 = _grad_target(b2)
        ops.conv_wgrad(y1, 0, c, g, 0, c, 3, w2buf, b2buf, accumulate=acc2, alpha=ctx.scale)
        d1 = torch.empty_like(x)
        ops.conv(g, 0, c, ctx.p2, w2, None, d1, 0, c, 3, mode=L.PACK_DGRAD, scale=ctx.scale, mask=(y1, 0))
        w1buf, acc1, dw1 = _grad_target(w1)
        b1buf, _, db1 = _grad_target(b1)
        ops.conv_wgrad(x, 0, c, d1, 0, c, 3, w1buf, b1buf, accumulate=acc1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.conv(d1, 0, c, ctx.p1, w1, None, dx, 0, c, 3, mode=L.PACK_DGRAD, res=(g, 0))
        return dx, dw1, db1, dw2, db2, None, None, None


class RCABFn(Function):
    """RCAN residual channel-attention block (rcan.py:33-55):
    out = CA(conv2(relu(conv1(x)))) + x, CA = rcan.py:10-29.  conv2's epilogue emits the pooled
    sums, one streaming kernel applies gate + skip."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, cw1, cb1, cw2, cb2, p1: PackedWeights, p2: PackedWeights):
        x = x.contiguous()
        n, h, w, c = x.shape
        y1 = torch.empty_like(x)
        ops.conv(x, 0, c, p1, w1, b1.detach(), y1, 0, c, 3, relu=True)
        t = torch.empty_like(x)
        pool = ops.zeros_f32((n, c), x.device)
        ops.conv(y1, 0, c, p2, w2, b2.detach(), t, 0, c, 3, colsum=pool, colsum_groups=n)
        out = torch.empty_like(x)
        s = torch.empty((n, c), dtype=torch.float32, device=x.device)
        yg = torch.empty((n, c), dtype=torch.float32, device=x.device)
        ops.ca_fwd(t, x, pool, False, cw1.detach(), cb1.detach(), cw2.detach(), cb2.detach(), out, s, yg)
        ctx.save_for_backward(x, y1, t, s, yg, w1, b1, w2, b2, cw1, cb1, cw2, cb2)
        ctx.p1, ctx.p2 = p1, p2
        return out

    @staticmethod
    def backward(ctx, g):
        x, y1, t, s, yg, w1, b1, w2, b2, cw1, cb1, cw2, cb2 = ctx.saved_tensors
        g = g.contiguous()
        n, h, w, c = x.shape
        dt = torch.empty_like(t)
        cw1buf, cacc, dcw1 = _grad_target(cw1)
        cb1buf, _, dcb1 = _grad_target(cb1)
        cw2buf, _, dcw2 = _grad_target(cw2)
        cb2buf, _, dcb2 = _grad_target(cb2)
        b2buf, bacc2, db2 = _grad_target(b2)
        scratch = ops.zeros_f32((n, c), x.device)
        if cacc != bacc2:
            raise RuntimeError("inconsistent gradient-accumulation state between CA and conv parameters")
        ops.ca_bwd(g, t, s, yg, cw1, cb1, cw2, cb2, dt, cw1buf, cb1buf, cw2buf, cb2buf, b2buf, scratch, accumulate=cacc,
                   scratch_is_zero=True)
        w2buf, acc2, dw2 = _grad_target(w2)
        ops.conv_wgrad(y1, 0, c, dt, 0, c, 3, w2buf, None, accumulate=acc2)
        d1 = torch.empty_like(x)
        b1buf, bacc1, db1 = _grad_target(b1)
        if not bacc1:
            b1buf.zero_()
        ops.conv(dt, 0, c, ctx.p2, w2, None, d1, 0, c, 3, mode=L.PACK_DGRAD, mask=(y1, 0), colsum=b1buf, colsum_groups=1)
        w1buf, acc1, dw1 = _grad_target(w1)
        ops.conv_wgrad(x, 0, c, d1, 0, c, 3, w1buf, None, accumulate=acc1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.conv(d1, 0, c, ctx.p1, w1, None, dx, 0, c, 3, mode=L.PACK_DGRAD, res=(g, 0))
        return dx, dw1, db1, dw2, db2, dcw1, dcb1, dcw2, dcb2, None, None


def chain_enabled() -> bool:
    """Layer-chain kernels (srb_conv_chain) are the default for the 64-channel bf16 trunks;
    SRB200_NO_CHAIN=1 selects the one-launch-per-layer path (A/B measurements, debugging)."""
    import os
    return os.environ.get("SRB200_NO_CHAIN", "0") in ("", "0")


def _zeroed_grad_target(p):
    """Like _grad_target for gradients the kernels ACCUMULATE with atomics: the buffer is zero
    unless it is a live slice of the flat gradient buffer."""
    buf, acc, ret = _grad_target(p)
    if not acc:
        buf.zero_()
    return buf, ret


class RCANGroupFn(Function):
    """RCAN ResidualGroup (rcan.py:59-74): n x RCAB (rcan.py:33-55) + conv + skip, forward and
    backward each as ONE persistent chain launch (ops.Chain -> srb_conv_chain).

    Forward space 0 ("A", saved for backward) holds, per RCAB b: slot 3b = relu(conv1), 3b+1 = t =
    conv2 output (pre-attention), 3b+2 = block output; slot 3n = group output.  Backward space 0
    ("B"): slot 3b = dt, 3b+1 = d(relu(conv1)) masked, 3b+2 = dL/d(block input); slot 3n = dL/d(last
    block output).  params = per RCAB (w1, b1, w2, b2, ca_w1, ca_b1, ca_w2, ca_b2), then (w_tail, b_tail)."""

    @staticmethod
    def forward(ctx, x, owner, *params):
        x = x.contiguous()
        n, h, w, c = x.shape
        assert c == 64 and x.dtype == torch.bfloat16
        nb = (len(params) - 2) // 8
        dev = x.device
        A = torch.empty((3 * nb + 1, n, h, w, 64), dtype=x.dtype, device=dev)
        s_all = torch.empty((nb, n, 64), dtype=torch.float32, device=dev)
        y_all = torch.empty((nb, n, 64), dtype=torch.float32, device=dev)
        pool = ops.zeros_f32((nb, n, 64), dev)
        bank = owner.filter_bank(L.PACK_FWD)
        ch = ops.Chain(n, h, w, dev)
        ch.space(0, A)
        ch.space(2, x.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        xin = ref(2, 0)
        cur = xin
        for b in range(nb):
            w1, b1, w2, b2, cw1, cb1, cw2, cb2 = (t.detach() for t in params[8 * b:8 * b + 8])
            ch.conv(cur, ref(0, 3 * b), 2 * b, b1, relu=True)
            ch.conv_ca(ref(0, 3 * b), ref(0, 3 * b + 1), ref(0, 3 * b + 2), cur, 2 * b + 1, b2, pool[b],
                       cw1.reshape(cw1.shape[0], 64), cb1, cw2.reshape(64, cw2.shape[1]), cb2, s_all[b], y_all[b])
            cur = ref(0, 3 * b + 2)
        ch.conv(cur, ref(0, 3 * nb), 2 * nb, params[-1].detach(), res=xin)
        ch.run(bank, hint=ops.chain_forward_hint())
        ctx.save_for_backward(x, A, s_all, y_all, *params)
        ctx.owner, ctx.nb = owner, nb
        return A[3 * nb]

    @staticmethod
    def backward(ctx, g):
        x, A, s_all, y_all, *params = ctx.saved_tensors
        nb, owner = ctx.nb, ctx.owner
        g = g.contiguous()
        _, n, h, w, _ = A.shape
        dev = g.device
        B = torch.empty((3 * nb + 1, n, h, w, 64), dtype=A.dtype, device=dev)
        scratch = ops.zeros_f32((nb, n, 64), dev)
        bank = owner.filter_bank(L.PACK_DGRAD)
        ch = ops.Chain(n, h, w, dev)
        ch.space(0, B)
        ch.space(1, A)
        ch.space(2, g.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        grads = [None] * len(params)
        wq = []   # (x tensor, gy tensor, weight index, bias index or None)
        def ca_args(b):
            """Fused CALayer backward of RCAB b: consumes the dL/dout the current op produces."""
            w1, b1, w2, b2, cw1, cb1, cw2, cb2 = params[8 * b:8 * b + 8]
            db2, grads[8 * b + 3] = _zeroed_grad_target(b2)
            dcw1, grads[8 * b + 4] = _zeroed_grad_target(cw1)
            dcb1, grads[8 * b + 5] = _zeroed_grad_target(cb1)
            dcw2, grads[8 * b + 6] = _zeroed_grad_target(cw2)
            dcb2, grads[8 * b + 7] = _zeroed_grad_target(cb2)
            cr = cw1.shape[0]
            return dict(t=ref(1, 3 * b + 1), dt=ref(0, 3 * b), w1=cw1.detach().reshape(cr, 64), b1=cb1.detach(),
                        w2=cw2.detach().reshape(64, cr), b2=cb2.detach(), s=s_all[b], y=y_all[b], dw1=dcw1.view(cr, 64),
                        db1=dcb1, dw2=dcw2.view(64, cr), db2=dcb2, scratch=scratch[b], colsum_dt=db2)

        # group tail conv: out = conv(last) + x; its input gradient is dL/dout of the last RCAB
        # slots 3b+2 (b > 0) and 3nb hold dL/dout of an RCAB: read only as the residual two ops later, never by the host
        ch.conv(ref(2, 0), ref(0, 3 * nb), 2 * nb, ca_bwd=ca_args(nb - 1), y_scratch=True)
        wq.append((A[3 * nb - 1], g, len(params) - 2, len(params) - 1))
        gref = ref(0, 3 * nb)
        for b in range(nb - 1, -1, -1):
            b1 = params[8 * b + 1]
            db1, grads[8 * b + 1] = _zeroed_grad_target(b1)
            ch.conv(ref(0, 3 * b), ref(0, 3 * b + 1), 2 * b + 1, mask=ref(1, 3 * b), colsum=db1, colsum_groups=1)
            ch.conv(ref(0, 3 * b + 1), ref(0, 3 * b + 2), 2 * b, res=gref, ca_bwd=ca_args(b - 1) if b > 0 else None, y_scratch=b > 0)
            wq.append((A[3 * b], B[3 * b], 8 * b + 2, None))
            wq.append((A[3 * b - 1] if b > 0 else x, B[3 * b + 1], 8 * b, None))
            gref = ref(0, 3 * b + 2)
        if ops.chain_backward_hint() == 0:
            ops.wgrad_overlap_adopt()     # (first group of the pass only) what is already queued runs beside this chain
        # the cluster kernel (96 SMs) only pays when weight gradients run beside it; alone, the L2-flag kernel is faster
        ch.run(bank, hint=ops.chain_backward_hint())
        ops.wgrad_overlap_kick()      # the previous group's weight gradients start BEHIND this launch, on the SMs it leaves free
        late_bias = []
        with ops.wgrad_overlap_section(bool(ch.used_cluster)) as sec:     # under the next group's backward chain if possible
            for xt, gy, wi, bi in wq:
                wbuf, acc, grads[wi] = _grad_target(params[wi])
                bbuf = None
                if bi is not None:
                    bbuf, _, grads[bi] = _grad_target(params[bi])
                    if sec is not None:       # a column-sum launch per group would sit on the side stream's critical path:
                        late_bias.append((gy, bbuf, acc))     # leave the bias to the step's one batched column-sum launch
                        bbuf = None
                ops.conv_wgrad(xt, 0, 64, gy, 0, 64, 3, wbuf, bbuf, accumulate=acc)
        for gy, bbuf, acc in late_bias:
            ops.bias_grad(gy, 0, 64, bbuf, accumulate=acc)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(g)
            ops.add_channels(B[2], 0, g, 0, dx, 0, 64)
        return (dx, None, *grads)


class ResTrunkFn(Function):
    """EDSR body (edsr.py:27-30,46-47): n x ResBlock (common.py:74-109: conv-ReLU-conv, * res_scale,
    += x) + conv + global skip, forward and backward each as ONE chain launch (64-channel bf16).

    Forward space 0 ("A", saved): slot 2b = relu(conv1_b), 2b+1 = block output; slot 2n = trunk output.
    Backward space 0 ("B"): slot 2b = d(relu(conv1_b)) (masked, scaled), 2b+1 = dL/d(block input);
    slot 2n = dL/d(last block output).  params = per block (w1, b1, w2, b2), then (w_tail, b_tail)."""

    @staticmethod
    def forward(ctx, x, owner, scale: float, *params):
        x = x.contiguous()
        n, h, w, c = x.shape
        assert c == 64 and x.dtype == torch.bfloat16
        nb = (len(params) - 2) // 4
        dev = x.device
        A = torch.empty((2 * nb + 1, n, h, w, 64), dtype=x.dtype, device=dev)
        bank = owner.filter_bank(L.PACK_FWD)
        ch = ops.Chain(n, h, w, dev)
        ch.space(0, A)
        ch.space(2, x.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        xin = ref(2, 0)
        cur = xin
        for b in range(nb):
            w1, b1, w2, b2 = (t.detach() for t in params[4 * b:4 * b + 4])
            ch.conv(cur, ref(0, 2 * b), 2 * b, b1, relu=True)
            ch.conv(ref(0, 2 * b), ref(0, 2 * b + 1), 2 * b + 1, b2, scale=scale, res=cur)
            cur = ref(0, 2 * b + 1)
        ch.conv(cur, ref(0, 2 * nb), 2 * nb, params[-1].detach(), res=xin)
        ch.run(bank, hint=ops.chain_forward_hint())
        ctx.save_for_backward(x, A, *params)
        ctx.owner, ctx.nb, ctx.scale = owner, nb, scale
        return A[2 * nb]

    @staticmethod
    def backward(ctx, g):
        x, A, *params = ctx.saved_tensors
        nb, owner, scale = ctx.nb, ctx.owner, ctx.scale
        g = g.contiguous()
        _, n, h, w, _ = A.shape
        dev = g.device
        B = torch.empty((2 * nb + 1, n, h, w, 64), dtype=A.dtype, device=dev)
        bank = owner.filter_bank(L.PACK_DGRAD)
        ch = ops.Chain(n, h, w, dev)
        ch.space(0, B)
        ch.space(1, A)
        ch.space(2, g.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        grads = [None] * len(params)
        wq = []   # (x, gy, weight index, bias index or None, alpha)

        def b2_target(b):
            """bias gradient of block b's second conv = res_scale * column sums of dL/d(block output)."""
            buf, grads[4 * b + 3] = _zeroed_grad_target(params[4 * b + 3])
            return buf

        ch.conv(ref(2, 0), ref(0, 2 * nb), 2 * nb, colsum=b2_target(nb - 1), colsum_scale=scale)
        wq.append((A[2 * nb - 1], g, len(params) - 2, len(params) - 1, 1.0))
        gslot = 2 * nb
        for b in range(nb - 1, -1, -1):
            db1, grads[4 * b + 1] = _zeroed_grad_target(params[4 * b + 1])
            ch.conv(ref(0, gslot), ref(0, 2 * b), 2 * b + 1, scale=scale, mask=ref(1, 2 * b), colsum=db1)
            if b > 0:
                ch.conv(ref(0, 2 * b), ref(0, 2 * b + 1), 2 * b, res=ref(0, gslot), colsum=b2_target(b - 1), colsum_scale=scale)
            else:
                ch.conv(ref(0, 2 * b), ref(0, 2 * b + 1), 2 * b, res=ref(0, gslot))
            wq.append((A[2 * b], B[gslot], 4 * b + 2, None, scale))
            wq.append((A[2 * b - 1] if b > 0 else x, B[2 * b], 4 * b, None, 1.0))
            gslot = 2 * b + 1
        # a single chain: nothing to run beside it, the L2-flag kernel on all SMs is faster (SRB200_CHAIN_BWD=cluster overrides)
        ch.run(bank, hint=0 if __import__("os").environ.get("SRB200_CHAIN_BWD") == "cluster" else 1)
        for xt, gy, wi, bi, alpha in wq:
            wbuf, acc, grads[wi] = _grad_target(params[wi])
            bbuf = None
            if bi is not None:
                bbuf, _, grads[bi] = _grad_target(params[bi])
            ops.conv_wgrad(xt, 0, 64, gy, 0, 64, 3, wbuf, bbuf, accumulate=acc, alpha=alpha)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(g)
            ops.add_channels(B[1], 0, g, 0, dx, 0, 64)
        return (dx, None, None, *grads)


class RDBFn(Function):
    """RDN residual dense block (rdn.py:24-40).  The C dense layers read a channel prefix of ONE
    [N,H,W,G0+C*G] buffer and write their G new channels in place (no torch.cat, rdn.py:21);
    LFF 1x1 + skip (rdn.py:40) closes the block."""

    @staticmethod
    def forward(ctx, x, packs, g0: int, g: int, *params):
        # params = (w_0, b_0, ..., w_{C-1}, b_{C-1}, w_lff, b_lff)
        x = x.contiguous()
        n, h, w, _ = x.shape
        nl = (len(params) - 2) // 2
        ctot = g0 + nl * g
        dbuf = torch.empty((n, h, w, ctot), dtype=x.dtype, device=x.device)
        ops.copy_channels(x, 0, dbuf, 0, g0)
        for c in range(nl):
            ops.conv(dbuf, 0, g0 + c * g, packs[c], params[2 * c], params[2 * c + 1].detach(), dbuf, g0 + c * g, g, 3,
                     relu=True)
        out = torch.empty_like(x)
        ops.conv(dbuf, 0, ctot, packs[nl], params[-2], params[-1].detach(), out, 0, g0, 1, res=(x, 0))
        ctx.save_for_backward(dbuf, *params)
        ctx.packs, ctx.g0, ctx.g, ctx.nl = packs, g0, g, nl
        return out

    @staticmethod
    def backward(ctx, gout):
        dbuf, *params = ctx.saved_tensors
        packs, g0, g, nl = ctx.packs, ctx.g0, ctx.g, ctx.nl
        gout = gout.contiguous()
        ctot = g0 + nl * g
        grads = [None] * len(params)
        # LFF (1x1): weight grad, then its input gradient fills the whole dense gradient buffer
        wbuf, acc, grads[-2] = _grad_target(params[-2])
        bbuf, _, grads[-1] = _grad_target(params[-1])
        ops.conv_wgrad(dbuf, 0, ctot, gout, 0, g0, 1, wbuf, bbuf, accumulate=acc)
        gd = torch.empty_like(dbuf)
        ops.conv(gout, 0, g0, packs[nl], params[-2], None, gd, 0, ctot, 1, mode=L.PACK_DGRAD)
        for c in range(nl - 1, -1, -1):
            off = g0 + c * g
            ops.relu_bwd(gd, off, dbuf, off, gd, off, g)     # in place on the slice
            wbuf, acc, grads[2 * c] = _grad_target(params[2 * c])
            bbuf, _, grads[2 * c + 1] = _grad_target(params[2 * c + 1])
            ops.conv_wgrad(dbuf, 0, off, gd, off, g, 3, wbuf, bbuf, accumulate=acc)
            # gd[..., :off] += dgrad(gd[..., off:off+g])   (read-modify-write of the same element)
            ops.conv(gd, off, g, packs[c], params[2 * c], None, gd, 0, off, 3, mode=L.PACK_DGRAD, res=(gd, 0))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(gout)
            ops.add_channels(gd, 0, gout, 0, dx, 0, g0)
        return (dx, None, None, None, *grads)


class ConcatConv1x1Fn(Function):
    """GFF.0 (rdn.py:70-71,108): 1x1 conv over the channel concatenation of all RDB outputs.
    The concat buffer is filled by channel-slice copies; its gradient is handed back as slices."""

    @staticmethod
    def forward(ctx, weight, bias, packs: PackedWeights, *xs):
        n, h, w, c = xs[0].shape
        ctot = c * len(xs)
        cat = torch.empty((n, h, w, ctot), dtype=xs[0].dtype, device=xs[0].device)
        for i, xi in enumerate(xs):
            ops.copy_channels(xi.contiguous(), 0, cat, i * c, c)
        cout = weight.shape[0]
        y = torch.empty((n, h, w, cout), dtype=cat.dtype, device=cat.device)
        ops.conv(cat, 0, ctot, packs, weight, bias.detach(), y, 0, cout, 1)
        ctx.save_for_backward(cat, weight, bias)
        ctx.packs, ctx.c, ctx.k = packs, c, len(xs)
        return y

    @staticmethod
    def backward(ctx, g):
        cat, weight, bias = ctx.saved_tensors
        g = g.contiguous()
        cout, ctot = weight.shape[0], weight.shape[1]
        wbuf, acc, dw = _grad_target(weight)
        bbuf, _, db = _grad_target(bias)
        ops.conv_wgrad(cat, 0, ctot, g, 0, cout, 1, wbuf, bbuf, accumulate=acc)
        gcat = torch.empty_like(cat)
        ops.conv(g, 0, cout, ctx.packs, weight, None, gcat, 0, ctot, 1, mode=L.PACK_DGRAD)
        outs = []
        for i in range(ctx.k):
            gi = torch.empty(cat.shape[:3] + (ctx.c,), dtype=cat.dtype, device=cat.device)
            ops.copy_channels(gcat, i * ctx.c, gi, 0, ctx.c)
            outs.append(gi)
        return (dw, db, None, *outs)


class BnActFn(Function):
    """y = PReLU_a( BatchNorm(x) ) + residual on NHWC tensors — each of the three optional — as one element-wise launch
    (+ two statistics launches in training mode): the conv -> norm -> act chains of the reference's BasicBlock / ResBlock
    (common.py:33-55,74-109) as SRResNet builds them (srresnet.py:13-30).  Training mode normalises with the batch
    statistics and updates running_mean / running_var in place, evaluation mode uses the running statistics (forward only).
    gamma / beta / a are fp32 parameters; a is nn.PReLU()'s single slope."""

    @staticmethod
    def forward(ctx, x, gamma, beta, prelu_a, residual, running_mean, running_var, training: bool, eps: float, momentum: float):
        x = x.contiguous()
        c = gamma.shape[0] if gamma is not None else x.shape[3]
        mean = rstd = None
        if gamma is not None:
            if training:
                mean, rstd = ops.bn_stats(x, c, eps, momentum, running_mean, running_var)
            else:
                mean, rstd = running_mean.detach().float(), torch.rsqrt(running_var.detach().float() + eps)
        res = residual.contiguous() if residual is not None else None
        y = torch.empty_like(x)
        ops.bn_act_fwd(x, c, mean, rstd, gamma.detach() if gamma is not None else None, beta.detach() if beta is not None else None,
                       prelu_a.detach() if prelu_a is not None else None, res, y)
        ctx.save_for_backward(x, mean, rstd, gamma, beta, prelu_a)
        ctx.c, ctx.training, ctx.has_res = c, training, residual is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, mean, rstd, gamma, beta, prelu_a = ctx.saved_tensors
        assert gamma is None or ctx.training, "BatchNorm backward is implemented for training-mode statistics only"
        g = g.contiguous()
        dx = torch.empty_like(x)
        dg = db = da = None
        gbuf = bbuf = abuf = None
        acc = None
        if gamma is not None:
            gbuf, acc, dg = _grad_target(gamma)
            bbuf, acc_b, db = _grad_target(beta)
            assert acc == acc_b
        if prelu_a is not None:
            abuf, acc_a, da = _grad_target(prelu_a)
            if acc is None:
                acc = acc_a
            elif acc != acc_a:
                raise NotImplementedError("BatchNorm and PReLU gradients with different accumulation states")
        ops.bn_act_bwd(g, x, ctx.c, mean, rstd, gamma.detach() if gamma is not None else None,
                       beta.detach() if beta is not None else None, prelu_a.detach() if prelu_a is not None else None, dx, gbuf, bbuf,
                       abuf, bool(acc))
        return dx, dg, db, da, (g if ctx.has_res else None), None, None, None, None, None


class AddFn(Function):
    """Skip connection between two NHWC tensors (edsr.py:47, rdn.py:109) as one kernel."""

    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous()
        b = b.contiguous()
        out = torch.empty_like(a)
        ops.add_channels(a, 0, b, 0, out, 0, a.shape[3])
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class L1LossFn(Function):
    """nn.L1Loss() (srmodel.py:37,549) with its seed gradient sign(sr-hr)/n computed in the same pass."""

    @staticmethod
    def forward(ctx, sr, hr):
        loss, grad = ops.l1_loss(sr.contiguous(), hr.contiguous(), want_grad=True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def l1_loss(sr, hr):
    return L1LossFn.apply(sr, hr)
