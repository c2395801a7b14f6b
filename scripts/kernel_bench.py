"""GPU-time micro-benchmarks of individual libsrb200 kernels: each kernel is captured `reps` times
into a CUDA graph over rotating buffers (working set > L2) and the graph replay is timed with CUDA
events, so no host launch overhead is in the number.  Prints one line per kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

dev = torch.device("cuda", 0)
N, H, W, C = 16, 48, 48, 64
NBUF = 48


def timed(fn, reps=192):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for i in range(reps):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    bf = torch.bfloat16
    xs = [torch.randn(N, H, W, C, device=dev).to(bf) for _ in range(NBUF)]
    ys = [torch.empty(N, H, W, C, device=dev, dtype=bf) for _ in range(NBUF)]
    zs = [torch.randn(N, H, W, C, device=dev).to(bf) for _ in range(NBUF)]
    w = torch.randn(64, 64, 3, 3, device=dev) * 0.04
    b = torch.zeros(64, device=dev)
    packs = ops.PackedWeights()
    flop = 2.0 * N * H * W * 64 * 64 * 9
    out = {}

    def conv_plain(i):
        ops.conv(xs[i % NBUF], 0, 64, packs, w, b, ys[i % NBUF], 0, 64, 3, relu=True)
    out["conv3x3_64_relu"] = timed(conv_plain)

    def conv_res(i):
        ops.conv(xs[i % NBUF], 0, 64, packs, w, b, ys[i % NBUF], 0, 64, 3, res=(zs[i % NBUF], 0))
    out["conv3x3_64_residual"] = timed(conv_res)

    pool = torch.zeros(N, 64, device=dev)

    def conv_colsum(i):
        ops.conv(xs[i % NBUF], 0, 64, packs, w, b, ys[i % NBUF], 0, 64, 3, colsum=pool, colsum_groups=N)
    out["conv3x3_64_colsum"] = timed(conv_colsum)

    def conv_dgrad_mask(i):
        ops.conv(xs[i % NBUF], 0, 64, packs, w, None, ys[i % NBUF], 0, 64, 3, mode=L.PACK_DGRAD, mask=(zs[i % NBUF], 0),
                 colsum=pool[0], colsum_groups=1)
    out["dgrad3x3_64_mask_colsum"] = timed(conv_dgrad_mask)

    # L2-resident variant: same two buffers over and over
    def conv_hot(i):
        ops.conv(xs[i % 2], 0, 64, packs, w, b, ys[i % 2], 0, 64, 3, relu=True)
    out["conv3x3_64_relu_L2hot"] = timed(conv_hot)

    # channel attention
    cw1 = torch.randn(4, 64, device=dev) * 0.1
    cb1 = torch.zeros(4, device=dev)
    cw2 = torch.randn(64, 4, device=dev) * 0.1
    cb2 = torch.zeros(64, device=dev)
    s_out = torch.empty(N, 64, device=dev)
    y_out = torch.empty(N, 64, device=dev)
    pool2 = torch.rand(N, 64, device=dev) * 100

    def ca_f(i):
        ops.ca_fwd(xs[i % NBUF], zs[i % NBUF], pool2, False, cw1, cb1, cw2, cb2, ys[i % NBUF], s_out, y_out)
    out["ca_fwd_scale"] = timed(ca_f)
    dws = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    cs = torch.zeros(64, device=dev)
    scratch = torch.zeros(N, 64, device=dev)

    def ca_b(i):
        ops.ca_bwd(xs[i % NBUF], zs[i % NBUF], s_out, y_out, cw1, cb1, cw2, cb2, ys[i % NBUF], *dws, cs, scratch,
                   accumulate=True, scratch_is_zero=False)
    out["ca_bwd_reduce+apply(+memset)"] = timed(ca_b)

    # weight gradient: one layer per call, and 74 layers batched
    dw = torch.zeros(64, 64, 3, 3, device=dev)

    def wg_single(i):
        ops.conv_wgrad(xs[i % NBUF], 0, 64, zs[i % NBUF], 0, 64, 3, dw, None, accumulate=True)
    out["wgrad3x3_64_single_layer"] = timed(wg_single, reps=48)
    dws74 = [torch.zeros(64, 64, 3, 3, device=dev) for _ in range(74)]

    def wg_batch(i):
        with ops.deferred_wgrads(max_items=1000):
            for j in range(74):
                ops.conv_wgrad(xs[(i + j) % NBUF], 0, 64, zs[(i + j) % NBUF], 0, 64, 3, dws74[j], None, accumulate=True)
    out["wgrad3x3_64_batched_per_layer"] = timed(wg_batch, reps=6) / 74
    # tail-shaped layers (192x192, 64 <-> 3 channels) and the pixel-unshuffle
    xt = [torch.randn(N, 192, 192, 64, device=dev).to(bf) for _ in range(4)]
    y3 = [torch.empty(N, 192, 192, 8, device=dev, dtype=bf) for _ in range(4)]
    g3 = [torch.randn(N, 192, 192, 8, device=dev).to(bf) for _ in range(4)]
    wt = torch.randn(3, 64, 3, 3, device=dev) * 0.04
    bt = torch.zeros(3, device=dev)
    pk = ops.PackedWeights()
    out["tail_conv_64to3_fwd"] = timed(lambda i: ops.conv(xt[i % 4], 0, 64, pk, wt, bt, y3[i % 4], 0, 3, 3), reps=16)
    out["tail_conv_dgrad_3to64"] = timed(lambda i: ops.conv(g3[i % 4], 0, 3, pk, wt, None, xt[(i + 1) % 4], 0, 64, 3,
                                                          mode=L.PACK_DGRAD), reps=16)
    dwt = torch.zeros(3, 64, 3, 3, device=dev)
    dbt = torch.zeros(3, device=dev)
    out["tail_conv_wgrad"] = timed(lambda i: ops.conv_wgrad(xt[i % 4], 0, 64, g3[i % 4], 0, 3, 3, dwt, dbt, accumulate=True), reps=8)
    gs = [torch.randn(N, 192, 192, 64, device=dev).to(bf) for _ in range(2)]
    out["pixel_unshuffle_192"] = timed(lambda i: ops.pixel_unshuffle(gs[i % 2], 2), reps=8)
    for k, v in out.items():
        print(f"{k:36s} {v:9.2f} us   {flop / (v * 1e-6) / 1e12 if 'ca_' not in k else 0:8.1f} TFLOP/s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
