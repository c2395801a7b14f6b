"""GPU data path for training batches: what the reference's `DataLoader` workers do per sample with PIL
(/root/reference/srdata.py:57-92 `_get_item` train branch, :136-169 `_get_patch`, :514-516 the loader) — aligned random LR / HR
crop, rotation by a multiple of 90 degrees, flips, `to_tensor` — done by ONE kernel per batch (`srb_patch_batch`, csrc/data.cu)
from uint8 images that stay resident in HBM.  At 2 140 patches/s per GPU (17 k/s on an 8-GPU node) the PIL loader with
`cpu_count() // 2` workers is what a real training run would wait for.

The random choices are drawn on the host from a `random.Random`, in the reference's order and with its quirks (see
`PatchSampler.draw`), so a seeded run sees exactly the patches the reference's dataset would produce for the same image
indices; a step uploads 56 bytes per sample.  The LR images are made once, when an image is added, with the very call the
reference uses (`TF.resize(..., BICUBIC)` on the PIL image, srdata.py:228-229) if only the HR image is given.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import lib as L
from . import ops


class PatchSampler:
    """Resident image set + batch builder.

    add(hr, lr=None): HWC uint8 RGB arrays (numpy / torch) or PIL images.
    fill(lr_out, hr_out, indices=None): writes one batch into the given NCHW fp32 CUDA tensors — e.g. `TrainStep.x` and
    `TrainStep.hr`, so that no host batch exists at all — and returns the choices it made."""

    def __init__(self, scale: int, lr_patch: int, device="cuda:0", augment: bool = True, seed: int | None = None):
        self.scale, self.lr_patch, self.augment = int(scale), int(lr_patch), bool(augment)
        self.device = torch.device(device)
        self.rng = random.Random(seed)
        self.images = []          # (lr uint8 [h,w,3] cuda, hr uint8 [H,W,3] cuda)
        self._items_host = None
        self._items_dev = None

    @staticmethod
    def _to_u8(img) -> torch.Tensor:
        if isinstance(img, torch.Tensor):
            t = img
        elif isinstance(img, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(img))
        else:                                   # PIL image
            t = torch.from_numpy(np.asarray(img.convert("RGB")).copy())
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "images are HWC uint8 RGB"
        return t.contiguous()

    def add(self, hr, lr=None):
        if lr is None:
            # the reference's own LR: antialiased bicubic of the WHOLE image by PIL (srdata.py:226-229), done once here
            from PIL import Image
            import torchvision.transforms.functional as TF
            from torchvision.transforms import InterpolationMode
            pil = hr if isinstance(hr, Image.Image) else Image.fromarray(self._to_u8(hr).numpy())
            w, h = pil.size
            lr = TF.resize(pil, [h // self.scale, w // self.scale], interpolation=InterpolationMode.BICUBIC)
        lr_t, hr_t = self._to_u8(lr), self._to_u8(hr)
        assert lr_t.shape[0] >= self.lr_patch and lr_t.shape[1] >= self.lr_patch, "image smaller than the patch"
        self.images.append((lr_t.to(self.device), hr_t.to(self.device)))
        return len(self.images) - 1

    def draw(self, index: int):
        """One sample's random choices exactly as srdata.py:165-166,77-92 makes them (the crop's top comes from the WIDTH
        range and its left from the HEIGHT range, because `_get_patch` reads PIL's (w, h) as (h, w); boxes that leave a
        non-square image are padded with black, as PIL does)."""
        lr = self.images[index][0]
        lr_h, lr_w = lr.shape[1], lr.shape[0]        # sic: PIL .size is (width, height)
        top = self.rng.randrange(0, lr_h - self.lr_patch + 1)
        left = self.rng.randrange(0, lr_w - self.lr_patch + 1)
        angle, hflip, vflip = 0, False, False
        if self.augment:
            angle = self.rng.choice((0, 90, 180, 270))
            hflip = self.rng.choice((True, False))
            vflip = self.rng.choice((True, False))
        return top, left, angle, hflip, vflip

    def fill(self, lr_out: torch.Tensor | None, hr_out: torch.Tensor | None, indices=None, choices=None):
        ref = lr_out if lr_out is not None else hr_out
        n = ref.shape[0]
        if indices is None:
            indices = [self.rng.randrange(len(self.images)) for _ in range(n)]
        assert len(indices) == n
        if choices is None:
            choices = [self.draw(i) for i in indices]
        p, s = self.lr_patch, self.scale
        for t, size in ((lr_out, p), (hr_out, p * s)):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (n, 3, size, size), \
                    f"expected a contiguous fp32 CUDA tensor [{n},3,{size},{size}]"
        if self._items_host is None or self._items_host.numel() != n * C.sizeof(L.PatchItem):
            self._items_host = torch.empty(n * C.sizeof(L.PatchItem), dtype=torch.uint8).pin_memory()
            self._items_dev = torch.empty(n * C.sizeof(L.PatchItem), dtype=torch.uint8, device=self.device)
        arr = (L.PatchItem * n).from_address(self._items_host.data_ptr())
        for k, (i, (top, left, angle, hflip, vflip)) in enumerate(zip(indices, choices)):
            lr, hr = self.images[i]
            it = arr[k]
            it.lr_img, it.hr_img = lr.data_ptr(), hr.data_ptr()
            it.lr_h, it.lr_w, it.hr_h, it.hr_w = lr.shape[0], lr.shape[1], hr.shape[0], hr.shape[1]
            it.lr_top, it.lr_left, it.angle, it.hflip, it.vflip, it.reserved = top, left, angle, int(hflip), int(vflip), 0
        self._items_dev.copy_(self._items_host, non_blocking=True)
        L.check(L.load().srb_patch_batch(C.c_void_p(L.ctx(self.device.index)), C.c_void_p(self._items_dev.data_ptr()), n, p, s,
                                         C.c_void_p(lr_out.data_ptr()) if lr_out is not None else None,
                                         C.c_void_p(hr_out.data_ptr()) if hr_out is not None else None, ops._stream()),
                "srb_patch_batch")
        return list(zip(indices, choices))
