"""RCAN on the B200 path (mirror of /root/reference/models/rcan.py:10-129)."""
from __future__ import annotations

from typing import Any

import torch.nn as nn

from srb200 import functional as F200

from .common import DefaultConv2d, MeanShift, UpscaleBlock
from .srmodel import SRModel


class CALayer(nn.Module):
    """Channel attention parameters (reference rcan.py:10-29): conv_du.0 [C/r,C,1,1], conv_du.2
    [C,C/r,1,1].  Computed by the fused pooling/gate/scale kernels, never as convolutions."""

    def __init__(self, channel: int, reduction: int = 16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)   # placeholder (no parameters)
        self.conv_du = nn.Sequential(
            DefaultConv2d(channel, channel // reduction, 1),
            nn.ReLU(inplace=True),
            DefaultConv2d(channel // reduction, channel, 1),
            nn.Sigmoid())


class RCAB(nn.Module):
    """conv-ReLU-conv-CA + skip; `res_scale` is stored but unused, as in the reference (rcan.py:53)."""

    def __init__(self, n_feat: int, kernel_size: int, reduction: int, res_scale=1):
        super().__init__()
        self.body = nn.Sequential(
            DefaultConv2d(n_feat, n_feat, kernel_size),
            nn.ReLU(True),
            DefaultConv2d(n_feat, n_feat, kernel_size),
            CALayer(n_feat, reduction))
        self.res_scale = res_scale

    def forward(self, x):
        c1, c2, ca = self.body[0], self.body[2], self.body[3]
        d0, d2 = ca.conv_du[0], ca.conv_du[2]
        return F200.RCABFn.apply(x, c1.weight, c1.bias, c2.weight, c2.bias, d0.weight, d0.bias, d2.weight, d2.bias,
                                 c1.packs, c2.packs)


class ResidualGroup(nn.Module):
    """n_resblocks x RCAB + conv + skip (reference rcan.py:59-74)."""

    def __init__(self, n_feat: int, kernel_size: int, reduction: int, res_scale, n_resblocks: int):
        super().__init__()
        body = [RCAB(n_feat, kernel_size, reduction, res_scale=1) for _ in range(n_resblocks)]
        body.append(DefaultConv2d(n_feat, n_feat, kernel_size))
        self.body = nn.Sequential(*body)

    def filter_bank(self, mode):
        """Contiguous packed 3x3 filters of the group in chain order: (conv1, conv2) per RCAB, tail."""
        from srb200 import ops
        if not hasattr(self, "_bank"):
            self._bank = ops.FilterBank()
        mods = list(self.body)
        convs = []
        for blk in mods[:-1]:
            convs.append((blk.body[0].weight, blk.body[0].packs))
            convs.append((blk.body[2].weight, blk.body[2].packs))
        convs.append((mods[-1].weight, mods[-1].packs))
        return self._bank.get(convs, mode)

    def _chain_ok(self, x) -> bool:
        import torch
        mods = list(self.body)
        if not (F200.chain_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[3] == 64):
            return False
        if len(mods) < 2:
            return False
        # the fused CALayer needs every tile of a sample on its own (CTA, chain): see srb_conv_chain
        from srb200 import ops
        n, h, w, _ = x.shape
        if ((h + 15) // 16) * ((w + 7) // 8) > 2 * ops.chain_grid(x.device, n, h, w):
            return False
        for blk in mods[:-1]:
            c1, c2, ca = blk.body[0], blk.body[2], blk.body[3]
            if tuple(c1.weight.shape) != (64, 64, 3, 3) or tuple(c2.weight.shape) != (64, 64, 3, 3):
                return False
            if c1.bias is None or c2.bias is None or ca.conv_du[0].weight.shape[0] > 16:
                return False
        return tuple(mods[-1].weight.shape) == (64, 64, 3, 3) and mods[-1].bias is not None

    def forward(self, x):
        mods = list(self.body)
        if self._chain_ok(x):
            params = []
            for blk in mods[:-1]:
                c1, c2, ca = blk.body[0], blk.body[2], blk.body[3]
                d0, d2 = ca.conv_du[0], ca.conv_du[2]
                params += [c1.weight, c1.bias, c2.weight, c2.bias, d0.weight, d0.bias, d2.weight, d2.bias]
            params += [mods[-1].weight, mods[-1].bias]
            return F200.RCANGroupFn.apply(x, self, *params)
        res = x
        for blk in mods[:-1]:
            res = blk(res)
        return mods[-1](res, residual=x)


class RCAN(SRModel):
    """sub_mean -> head -> n_resgroups x ResidualGroup + conv, global skip -> tail -> add_mean.
    Constructor defaults are the reference's (n_resblocks=16, rcan.py:82); the benchmark config
    uses 20 (run_comparisons.sh:40)."""

    def __init__(self, n_feats: int = 64, n_resblocks: int = 16, n_resgroups: int = 10, reduction: int = 16,
                 res_scale: float = 1, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        k = 3
        if self._channels == 3:
            self.sub_mean = MeanShift()
        self.head = nn.Sequential(DefaultConv2d(self._channels, n_feats, k))
        body = [ResidualGroup(n_feats, k, reduction, res_scale=res_scale, n_resblocks=n_resblocks)
                for _ in range(n_resgroups)]
        body.append(DefaultConv2d(n_feats, n_feats, k))
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(UpscaleBlock(self._scale_factor, n_feats), DefaultConv2d(n_feats, self._channels, k))
        if self._channels == 3:
            self.add_mean = MeanShift(sign=1)      # registered last, as in the reference (rcan.py:112-113)

    def forward(self, x):
        rgb = self._channels == 3
        x = F200.ToNHWC.apply(x, self.sub_mean.channel_add() if rgb else None, self.act_dtype)
        x = self.head[0](x)
        res = x
        mods = list(self.body)
        for grp in mods[:-1]:
            res = grp(res)
        res = mods[-1](res, residual=x)
        y = self.tail[0](res)
        y = self.tail[1](y)
        return F200.ToNCHW.apply(y, self.add_mean.channel_add() if rgb else None, self._channels)
