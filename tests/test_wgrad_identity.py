"""CPU restatement of the N = 128 weight-gradient scheme of csrc/wgrad_umma.cu (wgrad_umma_kernel<2>) in numpy.

The kernel stacks the taps (0,kw) | (1,kw) of the input window on M and uses the gradient tile twice on N —
[gy shifted up one row | gy] — so that one accumulator per kw yields three taps: (0,kw) and (1,kw) against the unshifted
block, (2,kw) as tap (1,kw) against the shifted block; the fourth quadrant repeats tap (1,kw) and is discarded.  Shifting
both operands of a product re-tiles the same pixel sum; the row it drops at the bottom edge multiplies the zero padding of
x, the row it adds at the top is the zero fill of gy.  This test walks the same 16x8 tiles with the same zero fills and
checks the identity against the plain weight gradient of a 3x3 'same' convolution
(autograd of nn.Conv2d, /root/reference/models/common.py:7-30), for ragged sizes too."""
import numpy as np
import pytest

TH, TW = 16, 8


def _padded(a, top, left, rows, cols):
    """rows x cols window of a [H, W, C] starting at (top, left), zero outside the image (TMA out-of-bounds fill)."""
    H, W, C = a.shape
    out = np.zeros((rows, cols, C), dtype=a.dtype)
    r0, r1 = max(top, 0), min(top + rows, H)
    c0, c1 = max(left, 0), min(left + cols, W)
    if r0 < r1 and c0 < c1:
        out[r0 - top:r1 - top, c0 - left:c1 - left] = a[r0:r1, c0:c1]
    return out


def wgrad_n128_scheme(x, gy):
    """x [N,H,W,Cin], gy [N,H,W,Cout] -> dW [Cout,Cin,3,3] the way wgrad_umma_kernel<2> accumulates it."""
    N, H, W, Cin = x.shape
    Cout = gy.shape[3]
    acc = np.zeros((3, 2, 2, Cin, Cout))          # [kw][A block: tap row 0/1][B block: shifted/unshifted]
    for n in range(N):
        for h0 in range(0, H, TH):
            for w0 in range(0, W, TW):
                win = _padded(x[n], h0 - 1, w0 - 1, TH + 2, TW + 2)      # one (16+2) x (8+2) window
                g17 = _padded(gy[n], h0 - 1, w0, TH + 1, TW)            # gradient tile + one row on top
                b = [g17[0:TH].reshape(-1, Cout), g17[1:TH + 1].reshape(-1, Cout)]
                for kw in range(3):
                    for kh in range(2):
                        a = win[kh:kh + TH, kw:kw + TW].reshape(-1, Cin)
                        for blk in range(2):
                            acc[kw, kh, blk] += a.T @ b[blk]
    dw = np.zeros((Cout, Cin, 3, 3))
    for kw in range(3):
        dw[:, :, 0, kw] = acc[kw, 0, 1].T      # lanes 0-63,  unshifted columns
        dw[:, :, 1, kw] = acc[kw, 1, 1].T      # lanes 64-127, unshifted columns
        dw[:, :, 2, kw] = acc[kw, 1, 0].T      # lanes 64-127, shifted columns; acc[kw, 0, 0] (tap 1 again) is discarded
    return dw, acc


def wgrad_direct(x, gy):
    N, H, W, Cin = x.shape
    Cout = gy.shape[3]
    xp = np.zeros((N, H + 2, W + 2, Cin))
    xp[:, 1:H + 1, 1:W + 1] = x
    dw = np.zeros((Cout, Cin, 3, 3))
    for kh in range(3):
        for kw in range(3):
            dw[:, :, kh, kw] = np.einsum("nhwo,nhwi->oi", gy, xp[:, kh:kh + H, kw:kw + W])
    return dw


@pytest.mark.parametrize("shape", [(2, 16, 8, 3, 4), (1, 48, 24, 2, 3), (2, 17, 9, 3, 2), (1, 5, 8, 2, 2), (1, 33, 20, 1, 5)])
def test_shifted_block_scheme_equals_the_weight_gradient(shape):
    N, H, W, Cin, Cout = shape
    rs = np.random.RandomState(H * 100 + W)
    x = rs.randn(N, H, W, Cin)
    gy = rs.randn(N, H, W, Cout)
    got, acc = wgrad_n128_scheme(x, gy)
    want = wgrad_direct(x, gy)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10)
    # the discarded quadrant is tap (1, kw) once more — minus, when the tiles end exactly at the image's last row, the row of
    # x * gy products the shift moves out of the last tile (which is why THAT quadrant is the one to drop: for tap (2, kw)
    # the same row multiplies x's zero padding)
    for kw in range(3):
        edge = np.einsum("nwo,nwi->io", gy[:, H - 1], np.pad(x, ((0, 0), (0, 0), (1, 1), (0, 0)))[:, H - 1, kw:kw + W])
        lost = edge if H % TH == 0 else 0.0
        assert np.allclose(acc[kw, 0, 0] + lost, want[:, :, 1, kw].T, rtol=1e-10, atol=1e-10)
