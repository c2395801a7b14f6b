"""Plain-numpy restatement of the ATen primitives the reference's hot path calls.

TEST INFRASTRUCTURE ONLY.  These are the published definitions (torch.nn docs) written as
explicit index arithmetic so the oracle does not rest on torch alone:

* conv2d        — cross-correlation, zero padding, stride 1 (reference call sites:
                  models/common.py:7-30, rcan.py:17-19, rdn.py:15,37,57-94, srcnn.py:17-21)
* pixel_shuffle — out[n, c, h*r+i, w*r+j] = in[n, c*r*r + i*r + j, h, w] (common.py:133)
* adaptive_avg_pool2d(.,1) — mean over H,W (rcan.py:14)
* conv2d_backward — dgrad (correlation with the 180-degree-rotated, channel-swapped filter),
                  wgrad (correlation of input with dY) and bias grad, i.e. what autograd
                  derives for models/*.py (SURVEY §2.2 last-but-one row)
"""
from __future__ import annotations

import numpy as np


def conv2d(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, pad: int) -> np.ndarray:
    n, cin, h, wd = x.shape
    cout, cin2, kh, kw = w.shape
    assert cin == cin2
    xp = np.zeros((n, cin, h + 2 * pad, wd + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    ho, wo = h + 2 * pad - kh + 1, wd + 2 * pad - kw + 1
    y = np.zeros((n, cout, ho, wo), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + ho, j:j + wo]                       # n,cin,ho,wo
            y += np.einsum("nchw,oc->nohw", patch, w[:, :, i, j].astype(np.float64))
    if b is not None:
        y += b.reshape(1, -1, 1, 1)
    return y


def conv2d_backward(x, w, gy, pad):
    """Returns (dx, dw, db) for y = conv2d(x, w, b, pad), stride 1."""
    n, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    xp = np.zeros((n, cin, h + 2 * pad, wd + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    ho, wo = gy.shape[2], gy.shape[3]
    dw = np.zeros_like(w, dtype=np.float64)
    dxp = np.zeros_like(xp)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + ho, j:j + wo]
            dw[:, :, i, j] = np.einsum("nohw,nchw->oc", gy, patch)
            dxp[:, :, i:i + ho, j:j + wo] += np.einsum("nohw,oc->nchw", gy, w[:, :, i, j].astype(np.float64))
    dx = dxp[:, :, pad:pad + h, pad:pad + wd]
    db = gy.sum(axis=(0, 2, 3))
    return dx, dw, db


def pixel_shuffle(x: np.ndarray, r: int) -> np.ndarray:
    n, c, h, w = x.shape
    co = c // (r * r)
    y = x.reshape(n, co, r, r, h, w).transpose(0, 1, 4, 2, 5, 3)
    return y.reshape(n, co, h * r, w * r)


def pixel_unshuffle(y: np.ndarray, r: int) -> np.ndarray:
    n, co, hr, wr = y.shape
    h, w = hr // r, wr // r
    x = y.reshape(n, co, h, r, w, r).transpose(0, 1, 3, 5, 2, 4)
    return x.reshape(n, co * r * r, h, w)


def global_avg_pool(x: np.ndarray) -> np.ndarray:
    return x.mean(axis=(2, 3), keepdims=True)


def ca_layer(x, w1, b1, w2, b2):
    """CALayer (rcan.py:23-29) on NCHW fp64."""
    s = global_avg_pool(x)[:, :, 0, 0]                    # n,c
    z = np.maximum(s @ w1[:, :, 0, 0].T + b1, 0.0)        # n,c/r
    u = z @ w2[:, :, 0, 0].T + b2
    y = 1.0 / (1.0 + np.exp(-u))
    return x * y[:, :, None, None]
