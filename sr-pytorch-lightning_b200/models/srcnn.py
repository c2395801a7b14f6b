"""SRCNN (mirror of /root/reference/models/srcnn.py:9-27): bicubic upsample, then 9x9 / 1x1 / 5x5
convolutions.  BASELINE config 1 — the reference's CPU-runnable case; here the three convs run on
the CUDA-core kernels of libsrb200 (K = 27*27.. tiny channel counts, not tensor-core shaped)."""
from __future__ import annotations

from typing import Any

import torch.nn as nn
import torch.nn.functional as F

from srb200 import functional as F200

from .common import DefaultConv2d
from .srmodel import SRModel


class SRCNN(SRModel):
    def __init__(self, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        self._net = nn.Sequential(
            DefaultConv2d(self._channels, 64, 9), nn.ReLU(True),
            DefaultConv2d(64, 32, 1), nn.ReLU(True),
            DefaultConv2d(32, self._channels, 5))

    def forward(self, x):
        x = F.interpolate(x, scale_factor=self._scale_factor, mode='bicubic')
        x = F200.ToNHWC.apply(x, None, self.act_dtype)
        x = self._net[0](x, relu=True)
        x = self._net[2](x, relu=True)
        x = self._net[4](x)
        return F200.ToNCHW.apply(x, None, self._channels)
