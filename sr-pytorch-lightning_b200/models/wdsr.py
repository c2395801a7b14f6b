"""WDSR on the B200 path (mirror of /root/reference/models/wdsr.py:9-118).

Every convolution of WDSR is wrapped in `nn.utils.weight_norm` (wdsr.py:65), so a conv's parameters are `bias`, `weight_g`
[Cout,1,1,1] and `weight_v` [Cout,Cin,k,k] — registered in that order, which is the state_dict layout kept here — and the
filter it applies is w = g * v / ||v|| (norm over everything but the output channel).  The re-parameterisation is parameter-
side arithmetic (Cout*Cin*k*k elements per layer, once per step): it is written with torch ops on the fp32 masters, so that
autograd carries dL/dw — which the conv kernels' weight-gradient launch produces — on to g and v; the convolutions themselves,
their ReLU, `* res_scale`, `+= x`, the PixelShuffle of tail / skip and the mean shift run in the libsrb200 kernels:
1x1 and 3x3 convs with >= 8 input channels on tcgen05 (`conv_umma_kernel`, partial channel blocks zero-filled by TMA),
the 3-channel head and the 5x5 skip conv on the CUDA-core kernels."""
from __future__ import annotations

import math
from typing import Any

import torch
import torch.nn as nn

from srb200 import functional as F200
from srb200.ops import PackedWeights

from .srmodel import SRModel


class WNConv2d(nn.Module):
    """`nn.utils.weight_norm(nn.Conv2d(cin, cout, k, padding=k//2))` (wdsr.py:17-20,39-44,73,87,93): parameters bias,
    weight_g, weight_v in the reference's registration order and with its initial values (kaiming-uniform v, g = ||v||)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        v = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_channels * kernel_size * kernel_size)
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1, 1))
        self.weight_v = nn.Parameter(v)
        self.packs = PackedWeights()

    def weight(self) -> torch.Tensor:
        v = self.weight_v
        return v * (self.weight_g / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1))

    def forward(self, x, relu: bool = False, scale: float = 1.0, residual=None, shuffle: int = 0):
        """x: NHWC activation."""
        return F200.ConvFn.apply(x, self.weight(), self.bias, residual, self.packs, relu, float(scale), int(shuffle))

    def extra_repr(self):
        return f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, weight_norm'


class _Block_A(nn.Module):
    """conv3x3 F->4F, ReLU, conv3x3 4F->F, * res_scale, += x (wdsr.py:9-27)."""

    def __init__(self, n_feats, kernel_size, res_scale=1):
        super().__init__()
        self.res_scale = res_scale
        self.body = nn.Sequential(WNConv2d(n_feats, 4 * n_feats, kernel_size), nn.ReLU(True),
                                  WNConv2d(4 * n_feats, n_feats, kernel_size))

    def forward(self, x):
        y = self.body[0](x, relu=True)
        return self.body[2](y, scale=self.res_scale, residual=x)


class _Block_B(nn.Module):
    """conv1x1 F->6F, ReLU, conv1x1 6F->int(0.8F), conv3x3 int(0.8F)->F, * res_scale, += x (wdsr.py:30-52)."""

    def __init__(self, n_feats, kernel_size, res_scale=1):
        super().__init__()
        self.res_scale = res_scale
        expand, linear = 6, 0.8
        self.body = nn.Sequential(WNConv2d(n_feats, n_feats * expand, 1), nn.ReLU(True),
                                  WNConv2d(n_feats * expand, int(n_feats * linear), 1),
                                  WNConv2d(int(n_feats * linear), n_feats, kernel_size))

    def forward(self, x):
        y = self.body[0](x, relu=True)
        y = self.body[2](y)
        return self.body[3](y, scale=self.res_scale, residual=x)


class WDSR(SRModel):
    """x - rgb_mean; s = PixelShuffle(skip 5x5 conv); head; n blocks; PixelShuffle(tail conv) + s; + rgb_mean (wdsr.py:55-118).

    state_dict keys as the reference: head.0.{bias,weight_g,weight_v}, body.{i}.body.{0,2[,3]}.*, tail.0.*, skip.0.*
    (rgb_mean is a plain attribute there, not a buffer, and is none here either)."""

    def __init__(self, type: str = 'B', n_feats: int = 128, n_resblocks: int = 16, res_scale: int = 1, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        k = 3
        block = _Block_A if type == 'A' else _Block_B
        out_feats = self._scale_factor * self._scale_factor * self._channels
        self.head = nn.Sequential(WNConv2d(self._channels, n_feats, 3))
        self.body = nn.Sequential(*[block(n_feats, k, res_scale=res_scale) for _ in range(n_resblocks)])
        self.tail = nn.Sequential(WNConv2d(n_feats, out_feats, 3), nn.PixelShuffle(self._scale_factor))
        self.skip = nn.Sequential(WNConv2d(3, out_feats, 5), nn.PixelShuffle(self._scale_factor))
        # wdsr.py:69-70: a plain attribute there; here non-persistent buffers (they follow .to(device), stay out of the
        # state_dict, and no tensor is created inside a captured step)
        mean = torch.tensor((0.4488, 0.4371, 0.4040), dtype=torch.float32)
        self.register_buffer("_mean_sub", -mean, persistent=False)
        self.register_buffer("_mean_add", mean.clone(), persistent=False)

    def forward(self, x):
        rgb = self._channels == 3
        x = F200.ToNHWC.apply(x, self._mean_sub if rgb else None, self.act_dtype)
        s = self.skip[0](x, shuffle=self._scale_factor)
        y = self.head[0](x)
        for blk in self.body:
            y = blk(y)
        y = self.tail[0](y, shuffle=self._scale_factor)
        y = F200.AddFn.apply(y, s)
        return F200.ToNCHW.apply(y, self._mean_add if rgb else None, self._channels)
