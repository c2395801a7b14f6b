"""Deterministic, library-independent synthetic weights and inputs.

TEST INFRASTRUCTURE ONLY.  A counter-based generator (splitmix64 finaliser) written in plain
numpy integer arithmetic, so the same (name, seed) gives bit-identical fp32 values on any box
and any numpy/torch version.  Used by oracle/make_golden.py to fill the *reference* modules
and by the tests to fill this repo's modules with the very same numbers — the committed
golden outputs then pin both.
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x.copy()
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform01(n: int, key: str, seed: int = 0) -> np.ndarray:
    """n float64 values in [0,1), 24-bit resolution (exactly representable in fp32)."""
    h = hashlib.sha256(f"{key}|{seed}".encode()).digest()
    base = np.uint64(int.from_bytes(h[:8], "little"))
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + base
        bits = _splitmix64(ctr)
    return (bits >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def synth_tensor(shape, key: str, seed: int = 0, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform01(n, key, seed)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def synth_state_dict(shapes: dict[str, tuple], seed: int = 0, gain: float = 1.0,
                     frozen: dict[str, np.ndarray] | None = None) -> dict[str, np.ndarray]:
    """Kaiming-uniform-like weights: conv weights U(-b, b) with b = gain*sqrt(3/fan_in)
    (variance-preserving, so deep stacks neither blow up nor vanish), biases U(-0.05, 0.05).
    `frozen` entries (MeanShift) are passed through unchanged."""
    out = {}
    for name in shapes:
        shp = tuple(shapes[name])
        if frozen is not None and name in frozen:
            out[name] = np.asarray(frozen[name], dtype=np.float32).reshape(shp)
            continue
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shp, dtype=np.int64)
        elif name.endswith("running_var"):
            out[name] = synth_tensor(shp, name, seed, 0.5, 1.5)
        elif name.endswith("running_mean"):
            out[name] = synth_tensor(shp, name, seed, -0.1, 0.1)
        elif name.endswith(".weight") and len(shp) == 1:        # PReLU slope (one element) / BatchNorm gamma
            out[name] = synth_tensor(shp, name, seed, 0.1, 0.4) if shp == (1,) else synth_tensor(shp, name, seed, 0.5, 1.5)
        elif name.endswith("weight_g"):        # weight-norm gains (wdsr.py:65): the filter's norm itself, keep it O(1) and positive
            out[name] = synth_tensor(shp, name, seed, 0.5 * gain, 1.5 * gain)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            b = gain * math.sqrt(3.0 / fan_in)
            out[name] = synth_tensor(shp, name, seed, -b, b)
        else:
            out[name] = synth_tensor(shp, name, seed, -0.05, 0.05)
    return out


def synth_image_batch(n: int, c: int, h: int, w: int, key: str = "lr", seed: int = 0) -> np.ndarray:
    """Smooth-ish image in [0,1]: low-frequency sinusoids plus hash noise, NCHW fp32."""
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    out = np.empty((n, c, h, w), dtype=np.float64)
    ph = uniform01(n * c * 4, key + "/phase", seed).reshape(n, c, 4)
    noise = uniform01(n * c * h * w, key + "/noise", seed).reshape(n, c, h, w)
    for i in range(n):
        for j in range(c):
            p = ph[i, j]
            base = 0.5 + 0.25 * np.sin(2 * np.pi * (p[0] + xx * (0.03 + 0.1 * p[1]))) \
                       * np.cos(2 * np.pi * (p[2] + yy * (0.02 + 0.1 * p[3])))
            out[i, j] = np.clip(base + 0.3 * (noise[i, j] - 0.5), 0.0, 1.0)
    return out.astype(np.float32)
