"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-sample training data path
(/root/reference/srdata.py:57-92 `_get_item` train branch, :136-169 `_get_patch`, :120-128 `TF.to_tensor`), in numpy index
arithmetic.  Only tests/ (and bench.py's checker legs) may import this; the product path is srb200/data.py + csrc/data.cu.

Pinned by tests/test_data.py against the third-party calls the reference itself makes (torchvision.transforms.functional
crop / rotate / hflip / vflip / to_tensor on PIL images; torchvision 0.26 / Pillow 12.2 in this image) on seeded random
images, every angle / flip combination, non-square images and boxes that leave the image.
"""
from __future__ import annotations

import random

import numpy as np


def draw(rng: random.Random, lr_size_wh, lr_patch: int, augment: bool = True):
    """The reference's random choices for one sample, in its order and with its quirk: `_get_patch` unpacks
    `lr_image.size` — PIL's (width, height) — as (h, w) (srdata.py:152-153), so the crop's TOP is drawn from the width
    range and its LEFT from the height range (srdata.py:165-166), and `TF.crop(img, top, left, ...)` may leave a
    non-square image (PIL pads with black).  Returns (top, left, angle, hflip, vflip)."""
    lr_h, lr_w = lr_size_wh                       # sic
    top = rng.randrange(0, lr_h - lr_patch + 1)
    left = rng.randrange(0, lr_w - lr_patch + 1)
    angle, hflip, vflip = 0, False, False
    if augment:
        angle = rng.choice((0, 90, 180, 270))     # srdata.py:78
        hflip = rng.choice((True, False))         # :83
        vflip = rng.choice((True, False))         # :88
    return top, left, angle, hflip, vflip


def crop(img: np.ndarray, top: int, left: int, size: int) -> np.ndarray:
    """TF.crop on a PIL image: the part of the box outside the image is black."""
    h, w, c = img.shape
    out = np.zeros((size, size, c), dtype=img.dtype)
    y0, y1 = max(top, 0), min(top + size, h)
    x0, x1 = max(left, 0), min(left + size, w)
    if y1 > y0 and x1 > x0:
        out[y0 - top:y1 - top, x0 - left:x1 - left] = img[y0:y1, x0:x1]
    return out


def augment(p: np.ndarray, angle: int, hflip: bool, vflip: bool) -> np.ndarray:
    """TF.rotate (PIL: counter-clockwise, an exact transpose for square images and multiples of 90), then TF.hflip,
    then TF.vflip."""
    if angle:
        p = np.rot90(p, k=angle // 90, axes=(0, 1))       # np.rot90 is counter-clockwise for axes (0, 1)
    if hflip:
        p = p[:, ::-1]
    if vflip:
        p = p[::-1]
    return np.ascontiguousarray(p)


def to_tensor(p: np.ndarray) -> np.ndarray:
    """TF.to_tensor: HWC uint8 -> CHW float32 / 255 (a true division, srdata.py:125-128)."""
    return (p.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def get_item(lr_img: np.ndarray, hr_img: np.ndarray, choice, lr_patch: int, scale: int):
    """(lr [3,p,p], hr [3,p*s,p*s]) float32 for one sample given draw()'s choice."""
    top, left, angle, hflip, vflip = choice
    lr = augment(crop(lr_img, top, left, lr_patch), angle, hflip, vflip)
    hr = augment(crop(hr_img, top * scale, left * scale, lr_patch * scale), angle, hflip, vflip)
    return to_tensor(lr), to_tensor(hr)
