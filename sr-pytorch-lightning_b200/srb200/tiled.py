"""Large-image inference, spatially tiled across GPUs (BASELINE.json config 5: EDSR x4,
960x540 -> 3840x2160 on 8 GPUs).  The reference has no tiling at all (`SRModel.predict_step`,
/root/reference/models/srmodel.py:375-380, pushes the whole frame through `forward` on one device);
this is the multi-GPU form of that same `forward` (edsr.py:40-54).

Partition: row strips of the LR frame (W kept whole so strip rows are contiguous NHWC slabs).
Every strip buffer carries `t` halo rows above and below its owned rows.  A conv runs over the whole
buffer; the outermost output rows are incomplete and are overwritten by the **per-layer halo
exchange**: each strip sends its first/last `t` owned rows to its neighbours.  True image borders
keep zeros in the halo (the conv's zero padding), so the result is bit-identical to the untiled
forward (same kernels, same per-pixel summation order) — tests/test_tiled_gpu.py checks equality.

Exchange back-ends: `LocalExchange` (all strips on this GPU; used for testing and for single-GPU
chopping of frames that do not fit) and `DistExchange` (one strip per rank, NCCL send/recv of the
contiguous row slabs over NVLink, batched per layer).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .functional import _padded


def partition_rows(h: int, parts: int):
    """[(r0, r1)] contiguous row ranges, sizes differing by at most one, larger strips first."""
    base, extra = divmod(h, parts)
    out, r = [], 0
    for i in range(parts):
        n = base + (1 if i < extra else 0)
        out.append((r, r + n))
        r += n
    return out


class LocalExchange:
    """All strips live in this process: the exchange is row copies between their buffers."""

    def __init__(self, parts: int):
        self.parts = parts

    def strips(self):
        return list(range(self.parts))

    def exchange(self, bufs, t: int):
        """bufs[i]: [1, hs_i + 2t, W, C]; fills each buffer's halo rows from its neighbours."""
        for i in range(self.parts):
            b = bufs[i]
            if i > 0:
                up = bufs[i - 1]
                b[:, :t].copy_(up[:, up.shape[1] - 2 * t: up.shape[1] - t])
            else:
                b[:, :t].zero_()
            if i < self.parts - 1:
                dn = bufs[i + 1]
                b[:, b.shape[1] - t:].copy_(dn[:, t:2 * t])
            else:
                b[:, b.shape[1] - t:].zero_()


class DistExchange:
    """One strip per rank of `group`; neighbours trade halo rows with batched P2P ops."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.parts = dist.get_world_size(group)

    def strips(self):
        return [self.rank]

    def exchange(self, bufs, t: int):
        b = bufs[0]
        ops_ = []
        hs2 = b.shape[1]
        up, dn = self.rank - 1, self.rank + 1
        if up >= 0:
            ops_.append(dist.P2POp(dist.isend, b[:, t:2 * t], dist.get_global_rank(self.group, up) if self.group else up, self.group))
            ops_.append(dist.P2POp(dist.irecv, b[:, :t], dist.get_global_rank(self.group, up) if self.group else up, self.group))
        else:
            b[:, :t].zero_()
        if dn < self.parts:
            ops_.append(dist.P2POp(dist.isend, b[:, hs2 - 2 * t: hs2 - t], dist.get_global_rank(self.group, dn) if self.group else dn, self.group))
            ops_.append(dist.P2POp(dist.irecv, b[:, hs2 - t:], dist.get_global_rank(self.group, dn) if self.group else dn, self.group))
        else:
            b[:, hs2 - t:].zero_()
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()


class TiledEDSR:
    """Strip-parallel `EDSR.forward` (no autograd).  `model` is a models.EDSR on this device."""

    def __init__(self, model, exchange):
        self.m = model
        self.ex = exchange

    def _conv(self, conv, xs, *, relu=False, scale=1.0, residual=None, shuffle=0):
        outs = []
        for i, x in enumerate(xs):
            n, h, w, _ = x.shape
            cout, cin = conv.weight.shape[0], conv.weight.shape[1]
            r = shuffle if shuffle > 1 else 1
            y = torch.empty((n, h * r, w * r, _padded(cout // (r * r), x.dtype)), dtype=x.dtype, device=x.device)
            b = conv.packs.get_bias(conv.bias, shuffle)
            ops.conv(x, 0, cin, conv.packs, conv.weight, b, y, 0, cout, conv.kernel_size, relu=relu, scale=scale,
                     shuffle=shuffle, res=(residual[i], 0) if residual is not None else None)
            outs.append(y)
        return outs

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """x: full LR frame [1,3,H,W] fp32 on this device (every rank holds it: it is 6 MB).
        Returns {strip index: SR rows [1,3,(r1-r0)*s,W*s] fp32} for the strips this process owns."""
        m = self.m
        assert x.shape[0] == 1, "tiled inference takes one frame at a time"
        H = x.shape[2]
        parts = partition_rows(H, self.ex.parts)
        mine = self.ex.strips()
        t = 1
        rgb = m._channels == 3
        add_in = m.sub_mean.channel_add() if rgb else None
        # input strips with halo straight from the frame (zero outside the image, AFTER mean shift)
        xs = []
        for i in mine:
            r0, r1 = parts[i]
            lo, hi = max(r0 - t, 0), min(r1 + t, H)
            cs = _padded(x.shape[1], m.act_dtype)
            buf = torch.zeros((1, (r1 - r0) + 2 * t, x.shape[3], cs), dtype=m.act_dtype, device=x.device)
            rows = buf[:, (lo - (r0 - t)):(lo - (r0 - t)) + (hi - lo)]     # contiguous row slab of buf
            ops.nchw_to_nhwc(x[:, :, lo:hi].contiguous(), add_in, m.act_dtype, out=rows)
            xs.append(buf)
        conv = self._conv
        ex = lambda bufs, tt: self.ex.exchange(bufs, tt)  # noqa: E731
        h0 = conv(m.head[0], xs)
        ex(h0, t)
        res = h0
        blocks = list(m.body)
        for blk in blocks[:-1]:
            c1, c2 = blk.body[0], blk.body[2]
            y1 = conv(c1, res, relu=True)
            ex(y1, t)
            res = conv(c2, y1, scale=blk.res_scale, residual=res)
            ex(res, t)
        res = conv(blocks[-1], res, residual=h0)
        ex(res, t)
        y = res
        up = list(m.tail[0])
        for cv, ps in zip(up[0::2], up[1::2]):
            y = conv(cv, y, shuffle=ps.upscale_factor)
            t *= ps.upscale_factor
            ex(y, t)
        y = conv(m.tail[1], y)
        out = {}
        add_out = m.add_mean.channel_add() if rgb else None
        for i, yy in zip(mine, y):
            owned = yy[:, t:yy.shape[1] - t].contiguous()
            out[i] = ops.nhwc_to_nchw(owned, 0, m._channels, add_out)
        return out

    @torch.no_grad()
    def forward_gathered(self, x: torch.Tensor):
        """Full SR frame on every rank (all_gather of the strips) / on this process (local)."""
        out = self.forward(x)
        if isinstance(self.ex, LocalExchange):
            return torch.cat([out[i] for i in sorted(out)], dim=2)
        s = self.m._scale_factor
        parts = partition_rows(x.shape[2], self.ex.parts)
        mine = out[self.ex.rank]
        bufs = [torch.empty((1, mine.shape[1], (r1 - r0) * s, mine.shape[3]), dtype=mine.dtype, device=mine.device)
                for r0, r1 in parts]
        dist.all_gather(bufs, mine, group=self.ex.group) if len({b.shape for b in bufs}) == 1 else \
            [dist.broadcast(bufs[i] if i != self.ex.rank else mine,
                            src=dist.get_global_rank(self.ex.group, i) if self.ex.group is not None else i, group=self.ex.group)
             for i in range(self.ex.parts)]
        bufs[self.ex.rank] = mine
        return torch.cat(bufs, dim=2)
