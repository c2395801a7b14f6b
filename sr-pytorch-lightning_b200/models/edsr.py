"""EDSR on the B200 path (mirror of /root/reference/models/edsr.py:9-54)."""
from __future__ import annotations

from typing import Any

import torch.nn as nn

from srb200 import functional as F200

from .common import DefaultConv2d, MeanShift, ResBlock, UpscaleBlock
from .srmodel import SRModel


class EDSR(SRModel):
    """sub_mean -> head 3->F -> n x ResBlock + conv, global skip -> UpscaleBlock -> conv F->3 -> add_mean.

    state_dict keys as the reference: sub_mean.*, add_mean.*, head.0.*, body.{i}.body.{0,2}.*,
    body.{n}.*, tail.0.{0,2}.*, tail.1.*."""

    def __init__(self, n_feats: int = 64, n_resblocks: int = 16, res_scale: float = 1, **kwargs: dict[str, Any]):
        super().__init__(**kwargs)
        k = 3
        if self._channels == 3:
            self.sub_mean = MeanShift()
            self.add_mean = MeanShift(sign=1)
        self.head = nn.Sequential(DefaultConv2d(self._channels, n_feats, k))
        body = [ResBlock(n_feats=n_feats, kernel_size=k, res_scale=res_scale) for _ in range(n_resblocks)]
        body.append(DefaultConv2d(n_feats, n_feats, k))
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(UpscaleBlock(self._scale_factor, n_feats), DefaultConv2d(n_feats, self._channels, k))

    def filter_bank(self, mode):
        """Contiguous packed 3x3 filters of the body in chain order: (conv1, conv2) per ResBlock, tail."""
        from srb200 import ops
        if not hasattr(self, "_bank"):
            self._bank = ops.FilterBank()
        blocks = list(self.body)
        convs = []
        for blk in blocks[:-1]:
            convs.append((blk.body[0].weight, blk.body[0].packs))
            convs.append((blk.body[2].weight, blk.body[2].packs))
        convs.append((blocks[-1].weight, blocks[-1].packs))
        return self._bank.get(convs, mode)

    def _chain_ok(self, x) -> bool:
        import torch
        blocks = list(self.body)
        if not (F200.chain_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[3] == 64 and len(blocks) >= 2):
            return False
        for blk in blocks[:-1]:
            if float(blk.res_scale) != float(blocks[0].res_scale):
                return False
            for conv in (blk.body[0], blk.body[2]):
                if tuple(conv.weight.shape) != (64, 64, 3, 3) or conv.bias is None:
                    return False
        return tuple(blocks[-1].weight.shape) == (64, 64, 3, 3) and blocks[-1].bias is not None

    def forward(self, x):
        rgb = self._channels == 3
        x = F200.ToNHWC.apply(x, self.sub_mean.channel_add() if rgb else None, self.act_dtype)
        x = self.head[0](x)
        res = x
        blocks = list(self.body)
        if self._chain_ok(x):
            # 64-channel trunk: the whole body (blocks + conv + global skip) is one persistent launch
            params = []
            for blk in blocks[:-1]:
                params += [blk.body[0].weight, blk.body[0].bias, blk.body[2].weight, blk.body[2].bias]
            params += [blocks[-1].weight, blocks[-1].bias]
            res = F200.ResTrunkFn.apply(x, self, float(blocks[0].res_scale), *params)
        else:
            for blk in blocks[:-1]:
                res = blk(res)
            res = blocks[-1](res, residual=x)       # body tail conv + global skip (edsr.py:46-47)
        y = self.tail[0](res)
        y = self.tail[1](y)
        return F200.ToNCHW.apply(y, self.add_mean.channel_add() if rgb else None, self._channels)
