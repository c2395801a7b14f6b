"""Chain kernel micro-benchmark: us per dependent layer for (a) a 40-deep relu-conv chain, (b) an
RCAN ResidualGroup forward (20 RCAB + tail, 41 ops) and backward (61 ops), each one launch, on
the bench shape [16,48,48,64] bf16.  CUDA events around graph replays of `reps` launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

dev = torch.device("cuda", 0)
N, H, W = 16, 48, 48
bf = torch.bfloat16


def timed(fn, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    depth = 40
    x = torch.randn(N, H, W, 64, device=dev).to(bf)
    ws = [(torch.randn(64, 64, 3, 3, device=dev) * 0.06).contiguous() for _ in range(depth)]
    bs = [torch.randn(64, device=dev) * 0.1 for _ in range(depth)]
    packs = [ops.PackedWeights() for _ in range(depth)]
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.zeros((depth, N, H, W, 64), dtype=bf, device=dev)
    ref = ops.Chain.ref
    counters = torch.zeros(depth * 2 * N + 64, device=dev)

    def relu_chain():
        counters.zero_()
        ops.set_arena(None)
        ch = ops.Chain(N, H, W, dev)
        ch.space(0, A)
        ch.space(1, x.view(1, N, H, W, 64))
        for i in range(depth):
            ch.conv(ref(1, 0) if i == 0 else ref(0, i - 1), ref(0, i), i, bs[i], relu=True)
        ch.run(bank)
    us = timed(relu_chain)
    flop = 2.0 * N * H * W * 64 * 64 * 9
    print(f"chain_relu_conv_x{depth}: {us:9.2f} us per launch, {us / depth:6.3f} us per dependent layer, "
          f"{flop * depth / us / 1e6:7.1f} TFLOP/s")

    # event trace of one launch (ns): [grid][chain][op][event]
    grid = ops.chain_grid(dev, N, H, W)
    trace = torch.zeros(grid * 2 * depth * 8 + grid * 4, dtype=torch.int64, device=dev)
    ch = ops.Chain(N, H, W, dev)
    ch.space(0, A)
    ch.space(1, x.view(1, N, H, W, 64))
    for i in range(depth):
        ch.conv(ref(1, 0) if i == 0 else ref(0, i - 1), ref(0, i), i, bs[i], relu=True)
    ch.run(bank, trace=trace)
    torch.cuda.synchronize()
    tail = trace[grid * 2 * depth * 8:].view(grid, 4).double().cpu()
    mhz = ((tail[:, 3] - tail[:, 1]) / (tail[:, 2] - tail[:, 0]) * 1e3)
    print(f"trace: SM clock during the chain kernel: {mhz.mean().item():.0f} MHz (min {mhz.min().item():.0f}, max {mhz.max().item():.0f})")
    t = trace[:grid * 2 * depth * 8].view(grid, 2, depth, 8).double().cpu()
    names = ["dep", "afull", "issued", "acc", "staged", "stored", "released"]
    t0 = t[:, :, 0, 0].min()
    rel = t[:, :, :, 6].amax(dim=(0, 1)) - t0
    d = (rel[1:] - rel[:-1]) / 1e3
    print(f"trace: steady-state op period {d[4:].mean().item():.3f} us (min {d[4:].min().item():.3f}, max {d[4:].max().item():.3f})")
    sl = slice(8, depth - 2)
    for a, b in [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6)]:
        seg = (t[:, :, sl, b] - t[:, :, sl, a])
        print(f"trace: {names[a]:>8s} -> {names[b]:<8s} mean {seg.mean().item():7.0f} ns  p10 {seg.quantile(0.1).item():7.0f}  p90 {seg.quantile(0.9).item():7.0f}")
    nxt = t[:, :, 9:depth - 1, 0] - t[:, :, 8:depth - 2, 6]
    print(f"trace: released(op) -> dep(op+1) same tile: mean {nxt.mean().item():7.0f} ns  p10 {nxt.quantile(0.1).item():7.0f}  p90 {nxt.quantile(0.9).item():7.0f}")
    for c in (0, 1):
        row = t[0, c, 10:13, :7] - t0
        print(f"trace: CTA0 chain{c} ops 10-12 (us):", " | ".join(" ".join(f"{v / 1e3:.2f}" for v in r.tolist()) for r in row))

    # RCAN ResidualGroup through the model classes
    import models
    torch.manual_seed(0)
    grp = models.rcan.ResidualGroup(64, 3, 16, 1, 20).to(dev)
    xin = torch.randn(N, H, W, 64, device=dev).to(bf).requires_grad_(True)
    gout = (torch.randn(N, H, W, 64, device=dev) * 0.01).to(bf)
    # event trace of the group's forward and backward chains
    ev = ["dep", "afull", "issued", "acc", "staged", "stored", "released", "pool"]

    def analyse(tr, n_ops, kinds, title):
        t = tr.view(grid, 2, n_ops, 8).double().cpu()
        t0 = t[:, :, 0, 0].min()
        done = (t[:, :, :, 6].amax(dim=(0, 1)) - t0) / 1e3
        per = done[1:] - done[:-1]
        print(f"{title}: total {done[-1].item():.1f} us over {n_ops} ops")
        for kind in sorted(set(kinds)):
            idx = [i for i, k in enumerate(kinds) if k == kind and 3 <= i < n_ops - 1]
            if not idx:
                continue
            sub = t[:, :, idx, :]
            period = per[[i - 1 for i in idx]].mean().item()
            def gap(a, b):
                m = (sub[..., a] > 0) & (sub[..., b] > 0)
                return ((sub[..., b] - sub[..., a])[m].mean().item()) if m.any() else float("nan")
            parts = [("dep>afull", gap(0, 1)), ("afull>issued", gap(1, 2)), ("issued>acc", gap(2, 3)), ("acc>staged", gap(3, 4)),
                     ("staged>pool", gap(4, 7)), ("pool>stored", gap(7, 5)), ("staged>stored", gap(4, 5)), ("stored>released", gap(5, 6)),
                     ("dep>released", gap(0, 6))]
            print(f"  {kind:10s} period {period:6.2f} us | " + "  ".join(f"{k} {v:6.0f}" for k, v in parts))

    for mode in ("chain", "layers"):
        os.environ["SRB200_NO_CHAIN"] = "0" if mode == "chain" else "1"

        def fwd():
            with torch.no_grad():
                grp(xin)
        us_f = timed(fwd, reps=5)

        def fwd_bwd():
            y = grp(xin)
            y.backward(gout)
        us_fb = timed(fwd_bwd, reps=5)
        print(f"rcan_group_20rcab[{mode}]: forward {us_f:8.1f} us ({us_f / 41:.2f} us per conv), "
              f"forward+backward(+wgrad) {us_fb:8.1f} us")

    os.environ["SRB200_NO_CHAIN"] = "0"
    nb = 20
    tr = torch.zeros(grid * 2 * 64 * 8 + grid * 4, dtype=torch.int64, device=dev)
    ops.CHAIN_TRACE = tr
    y = grp(xin)
    torch.cuda.synchronize()
    analyse(tr[:grid * 2 * (2 * nb + 1) * 8], 2 * nb + 1, ["conv1", "conv2+ca"] * nb + ["tail"], "group forward")
    tr.zero_()
    y.backward(gout)
    torch.cuda.synchronize()
    analyse(tr[:grid * 2 * (2 * nb + 1) * 8], 2 * nb + 1, ["tail_dgrad"] + ["dgrad2+mask", "dgrad1+cabwd"] * nb, "group backward")
    ops.CHAIN_TRACE = None


if __name__ == "__main__":
    main()
