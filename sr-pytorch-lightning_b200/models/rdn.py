"""RDN on the B200 path (mirror of /root/reference/models/rdn.py:9-111)."""
from __future__ import annotations

from typing import Any

import torch.nn as nn

from srb200 import functional as F200

from .common import DefaultConv2d
from .srmodel import SRModel


class _RDB_Conv(nn.Module):
    """Dense layer parameters: conv.0 = 3x3 Cin->G (+ReLU, + concat in the reference, rdn.py:9-21)."""

    def __init__(self, in_channels: int, grow_rate: int, k: int = 3):
        super().__init__()
        self.conv = nn.Sequential(DefaultConv2d(in_channels, grow_rate, k), nn.ReLU())


class _RDB(nn.Module):
    """Residual dense block (reference rdn.py:24-40): C dense layers writing into one channel-strided
    buffer, 1x1 local feature fusion, skip."""

    def __init__(self, g0: int, g: int, n_layers: int, k: int = 3):
        super().__init__()
        self.convs = nn.Sequential(*[_RDB_Conv(g0 + c * g, g, k) for c in range(n_layers)])
        self.LFF = DefaultConv2d(g0 + n_layers * g, g0, 1)
        self.g0, self.g = g0, g

    def forward(self, x):
        layers = [m.conv[0] for m in self.convs]
        params = []
        for m in layers:
            params += [m.weight, m.bias]
        params += [self.LFF.weight, self.LFF.bias]
        packs = [m.packs for m in layers] + [self.LFF.packs]
        return F200.RDBFn.apply(x, packs, self.g0, self.g, *params)


class RDN(SRModel):
    """SFENet1/2 -> D RDBs -> GFF (1x1 over the concat of all RDB outputs, 3x3) + f_-1 -> UPNet."""

    def __init__(self, rdn_config: str = 'B', G0: int = 64, kernel_size: int = 3, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        self.D, C, G = {'A': (20, 6, 32), 'B': (16, 8, 64)}[rdn_config]
        k = kernel_size
        self.SFENet1 = DefaultConv2d(self._channels, G0, k)
        self.SFENet2 = DefaultConv2d(G0, G0, k)
        self._RDBs = nn.ModuleList([_RDB(G0, G, C, k) for _ in range(self.D)])
        self.GFF = nn.Sequential(DefaultConv2d(self.D * G0, G0, 1), DefaultConv2d(G0, G0, k))
        s = self._scale_factor
        if s in (2, 3):
            self.UPNet = nn.Sequential(DefaultConv2d(G0, G * s * s, k), nn.PixelShuffle(s), DefaultConv2d(G, 3, k))
        elif s == 4:
            self.UPNet = nn.Sequential(DefaultConv2d(G0, G * 4, k), nn.PixelShuffle(2),
                                       DefaultConv2d(G, G * 4, k), nn.PixelShuffle(2),
                                       DefaultConv2d(G, self._channels, k))
        else:
            raise ValueError("scale must be 2 or 3 or 4.")

    def forward(self, x):
        x = F200.ToNHWC.apply(x, None, self.act_dtype)
        f1 = self.SFENet1(x)
        x = self.SFENet2(f1)
        outs = []
        for rdb in self._RDBs:
            x = rdb(x)
            outs.append(x)
        g0 = self.GFF[0]
        x = F200.ConcatConv1x1Fn.apply(g0.weight, g0.bias, g0.packs, *outs)
        x = self.GFF[1](x, residual=f1)            # `x += f__1` (rdn.py:109)
        mods = list(self.UPNet)
        i = 0
        while i < len(mods):
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.PixelShuffle):
                x = mods[i](x, shuffle=mods[i + 1].upscale_factor)
                i += 2
            else:
                x = mods[i](x)
                i += 1
        return F200.ToNCHW.apply(x, None, self.UPNet[-1].out_channels)
