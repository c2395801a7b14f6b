"""Building blocks of the B200 path (mirror of /root/reference/models/common.py:7-139).

The classes keep the reference's module tree — hence its `state_dict` keys, shapes and dtypes
(fp32 OIHW master weights) — but they do not compute with torch.nn: `forward` launches libsrb200
kernels on NHWC tensors through `srb200.functional`.  Packed / bf16 copies of the weights are
caches and never enter the state_dict.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from srb200 import functional as F200
from srb200.ops import PackedWeights


class DefaultConv2d(nn.Module):
    """Parameters of a stride-1 convolution that keeps H and W ('same' = k//2 padding,
    reference common.py:7-30).  Holds `weight` [Cout,Cin,k,k] and `bias` [Cout] like nn.Conv2d
    (default init: kaiming-uniform a=sqrt(5), as nn.Conv2d.reset_parameters)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, padding='same', bias: bool = True):
        super().__init__()
        if isinstance(padding, str):
            assert padding.lower() in ('valid', 'same')
            if padding.lower() == 'valid' and kernel_size != 1:
                raise ValueError("the B200 path implements 'same' convolutions only")
        elif padding != kernel_size // 2:
            raise ValueError("the B200 path implements padding = kernel_size // 2 only")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.packs = PackedWeights()
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size * self.kernel_size
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, relu: bool = False, scale: float = 1.0, residual=None, shuffle: int = 0):
        """x: NHWC activation."""
        return F200.ConvFn.apply(x, self.weight, self.bias, residual, self.packs, relu, float(scale), int(shuffle))

    def extra_repr(self):
        return f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}'


class MeanShift(nn.Module):
    """Frozen 1x1 conv with W = I/std, b = sign*range*mean/std (reference common.py:58-71).
    With std = 1 (the only configuration the reference builds) it is a per-channel add, which
    the layout-conversion kernels at the model boundary apply for free."""

    def __init__(self, rgb_range: int = 1, rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1.0, 1.0, 1.0), sign: int = -1):
        super().__init__()
        std = torch.tensor(rgb_std, dtype=torch.float32)
        self.weight = nn.Parameter(torch.eye(3).view(3, 3, 1, 1) / std.view(3, 1, 1, 1), requires_grad=False)
        self.bias = nn.Parameter(sign * rgb_range * torch.tensor(rgb_mean, dtype=torch.float32) / std,
                                 requires_grad=False)
        self._checked = None

    def channel_add(self) -> torch.Tensor:
        tag = (self.weight.data_ptr(), self.weight._version)
        if self._checked != tag:
            eye = torch.eye(3, device=self.weight.device).view(3, 3, 1, 1)
            if not torch.equal(self.weight.detach(), eye):
                raise ValueError('MeanShift with rgb_std != 1 is not supported by the fused boundary kernels')
            self._checked = tag
        return self.bias.detach()


class ResBlock(nn.Module):
    """conv-ReLU-conv, `* res_scale`, `+= x` (reference common.py:74-109) as two fused launches."""

    def __init__(self, n_feats: int = 64, kernel_size: int = 3, res_scale: float = 1.):
        super().__init__()
        self.body = nn.Sequential(
            DefaultConv2d(n_feats, n_feats, kernel_size),
            nn.ReLU(True),          # placeholder: keeps the reference's indices (body.0 / body.2)
            DefaultConv2d(n_feats, n_feats, kernel_size))
        self.res_scale = res_scale

    def forward(self, x):
        c1, c2 = self.body[0], self.body[2]
        if c1.kernel_size == 3:
            return F200.ResBlockFn.apply(x, c1.weight, c1.bias, c2.weight, c2.bias, c1.packs, c2.packs,
                                         float(self.res_scale))
        y = c1(x, relu=True)
        return c2(y, scale=self.res_scale, residual=x)


class UpscaleBlock(nn.Sequential):
    """log2(s) x [conv F -> F*r*r, PixelShuffle(r)] (reference common.py:112-139); the shuffle is
    folded into the conv's store addressing."""

    def __init__(self, scale_factor: int = 4, n_feats: int = 64, kernel_size: int = 3):
        assert scale_factor in {2, 3, 4, 8}
        layers = []
        for _ in range(int(math.log2(scale_factor))):
            r = 2 if scale_factor % 2 == 0 else 3
            layers += [DefaultConv2d(n_feats, n_feats * r * r, kernel_size), nn.PixelShuffle(r)]
        super().__init__(*layers)

    def forward(self, x):
        mods = list(self)
        for conv, ps in zip(mods[0::2], mods[1::2]):
            x = conv(x, shuffle=ps.upscale_factor)
        return x
