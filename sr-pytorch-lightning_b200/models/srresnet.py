"""SRResNet on the B200 path (mirror of /root/reference/models/srresnet.py:9-36).

The reference assembles it from common.py's BasicBlock / ResBlock / UpscaleBlock with `norm=nn.BatchNorm2d(n_feats)` and
`act=nn.PReLU()` INSTANCES: inside one ResBlock both convs are followed by the same BatchNorm module (common.py:94-98 appends
the instance it was given twice), and both stages of the UpscaleBlock by the same PReLU (common.py:136-137).  The module tree
below repeats that sharing, so the state_dict has the reference's keys (body.i.body.1.* and body.i.body.4.* are one module)
and the shared statistics / parameters behave as there (two running-statistics updates per block and step; their gradients
are the sums over both uses).

Compute: 3x3 64->64 and 64->256 convs on tcgen05 (`srb_conv`, PixelShuffle folded into the store), the 9x9 head (3->64) and
tail (64->3) on the CUDA-core kernels, conv -> BatchNorm -> PReLU (-> + x) as `BnActFn` = batch statistics (two reduction
passes) + one fused element-wise launch (csrc/bn.cu)."""
from __future__ import annotations

import math
from typing import Any

import torch.nn as nn

from srb200 import functional as F200

from .common import DefaultConv2d
from .srmodel import SRModel


def _bn_act(x, bn: nn.BatchNorm2d | None, act: nn.PReLU | None, residual=None):
    if bn is None:
        return F200.BnActFn.apply(x, None, None, act.weight, residual, None, None, False, 0.0, 0.0)
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked.add_(1)           # nn.BatchNorm2d does this first; the momentum is a constant here (0.1)
    return F200.BnActFn.apply(x, bn.weight, bn.bias, act.weight if act is not None else None, residual, bn.running_mean,
                              bn.running_var, bn.training, bn.eps, bn.momentum)


class _BasicBlock(nn.Sequential):
    """conv [-> norm] [-> act] (common.py:33-55)."""

    def __init__(self, in_channels, out_channels, kernel_size, norm=None, act=None):
        m = [DefaultConv2d(in_channels, out_channels, kernel_size)]
        if norm is not None:
            m.append(norm)
        if act is not None:
            m.append(act)
        super().__init__(*m)

    def forward(self, x):
        y = self[0](x)
        norm = next((m for m in self if isinstance(m, nn.BatchNorm2d)), None)
        act = next((m for m in self if isinstance(m, nn.PReLU)), None)
        if norm is None and act is None:
            return y
        return _bn_act(y, norm, act)


class _ResBlock(nn.Module):
    """conv, norm, act, conv, norm (the SAME norm module), * res_scale (= 1), += x (common.py:74-109 with n_conv_layers=2)."""

    def __init__(self, n_feats, kernel_size, norm, act):
        super().__init__()
        self.body = nn.Sequential(DefaultConv2d(n_feats, n_feats, kernel_size), norm, act,
                                  DefaultConv2d(n_feats, n_feats, kernel_size), norm)
        self.res_scale = 1.

    def forward(self, x):
        b = self.body
        y = _bn_act(b[0](x), b[1], b[2])
        return _bn_act(b[3](y), b[4], None, residual=x)


class _UpscaleBlock(nn.Sequential):
    """log2(s) x [conv F -> F*r*r, PixelShuffle(r), act] with ONE act module (common.py:112-139)."""

    def __init__(self, scale_factor, n_feats, act):
        assert scale_factor in {2, 3, 4, 8}
        layers = []
        for _ in range(int(math.log2(scale_factor))):
            r = 2 if scale_factor % 2 == 0 else 3
            layers += [DefaultConv2d(n_feats, n_feats * r * r, 3), nn.PixelShuffle(r), act]
        super().__init__(*layers)

    def forward(self, x):
        mods = list(self)
        for conv, ps, act in zip(mods[0::3], mods[1::3], mods[2::3]):
            x = conv(x, shuffle=ps.upscale_factor)
            x = _bn_act(x, None, act)        # a scalar-slope PReLU commutes with the shuffle: applied on the shuffled tensor
        return x


class SRResNet(SRModel):
    """head 9x9 + PReLU -> n x ResBlock(BN, PReLU) + conv-BN, global skip -> UpscaleBlock(PReLU) -> 9x9 conv (srresnet.py:9-36).
    No mean shift (the reference has none here)."""

    def __init__(self, n_resblocks: int = 16, n_feats: int = 64, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        self.head = _BasicBlock(self._channels, n_feats, 9, act=nn.PReLU())
        body = [_ResBlock(n_feats, 3, norm=nn.BatchNorm2d(n_feats), act=nn.PReLU()) for _ in range(n_resblocks)]
        body.append(_BasicBlock(n_feats, n_feats, 3, norm=nn.BatchNorm2d(n_feats), act=None))
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(_UpscaleBlock(self._scale_factor, n_feats, act=nn.PReLU()),
                                  DefaultConv2d(n_feats, self._channels, 9))

    def forward(self, x):
        x = F200.ToNHWC.apply(x, None, self.act_dtype)
        h = self.head(x)
        y = h
        for blk in self.body:
            y = blk(y)
        y = F200.AddFn.apply(y, h)
        y = self.tail[0](y)
        y = self.tail[1](y)
        return F200.ToNCHW.apply(y, None, self._channels)
