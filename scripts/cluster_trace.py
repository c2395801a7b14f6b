"""In-kernel timeline of the per-sample cluster chain kernel (conv_cluster.cu, CL_TRACE events), plus launch timings.
  python scripts/cluster_trace.py [long|group]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16
EV = ["mma:rdy0", "mma:iss0", "mma:rdy1", "mma:issAll", "epi:acc0", "epi:ld0", "epi:st0", "epi:fence0", "epi:pub0",
      "epi:acc1", "epi:pub1", "epi:acc2", "epi:pub2", "epi:end", "epi:gather", "-"]


def rnd(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def report(tr, n_ops, ctas, label, kinds=None):
    """tr: int64 [ctas, n_ops, 16] ns"""
    t = tr.astype(np.float64)
    t0 = t[:, 0, 0].min()
    print(f"--- {label}: {ctas} CTAs x {n_ops} ops; kernel span {(t[:, -1, 13].max() - t0) / 1e3:.1f} us")
    per = np.diff(t[:, :, 0], axis=1)        # op period seen by the MMA thread
    mid = slice(n_ops // 4, 3 * n_ops // 4)
    print(f"op period (mma:rdy0 -> next mma:rdy0), middle ops: mean {per[:, mid].mean():.0f} ns  p10 {np.percentile(per[:, mid], 10):.0f}  p90 {np.percentile(per[:, mid], 90):.0f}")
    if kinds is not None:
        for k in sorted(set(kinds[:-1])):
            sel = [i for i in range(n_ops - 1) if kinds[i] == k]
            print(f"  kind {k}: period mean {per[:, sel].mean():.0f} ns over {len(sel)} ops")
    def d(a, b, ops_sel=mid):
        x = t[:, ops_sel, b] - t[:, ops_sel, a]
        return x.mean()
    pairs = [(0, 1, "rdy0 -> q0 MMAs issued"), (0, 2, "rdy0 -> rdy1 (q1 window ready)"), (0, 3, "rdy0 -> all MMAs issued"),
             (0, 4, "rdy0 -> acc0 full (epilogue sees q0)"), (4, 5, "acc0 -> TMEM loaded"), (5, 6, "loaded -> stores issued"),
             (6, 7, "stores -> proxy fence done"), (7, 8, "fence -> published (bar + arrive)"), (4, 8, "acc0 -> q0 published"),
             (9, 10, "acc1 -> q1 published"), (11, 12, "acc2 -> q2 published"), (3, 11, "all issued -> acc2 full"),
             (8, 9, "q0 published -> acc1 full"), (10, 11, "q1 published -> acc2 full"), (0, 13, "rdy0 -> op end")]
    for a, b, name in pairs:
        print(f"  {name:40s} {d(a, b):8.0f} ns")
    # next op's rdy0 relative to this op's publishes
    nxt = t[:, 1:, 0]
    print(f"  {'q1 published -> NEXT op rdy0':40s} {(nxt[:, mid] - t[:, :-1, :][:, mid, 10]).mean():8.0f} ns")
    print(f"  {'q2 published -> NEXT op rdy1':40s} {(t[:, 1:, 2][:, mid] - t[:, :-1, :][:, mid, 12]).mean():8.0f} ns")
    w = t[:, mid, 16:24] - t[:, mid, 9:10]
    print("  q1: acc1 -> tile-barrier arrive per epilogue warp (ns): " + " ".join(f"{x:.0f}" for x in w.mean(axis=(0, 1))))
    print(f"  q1: last warp arrive -> publisher released {(t[:, mid, 24] - t[:, mid, 16:24].max(axis=2)).mean():.0f} ns;  released -> arrives issued (lane 0) {(t[:, mid, 10] - t[:, mid, 24]).mean():.0f} ns, (lane 1) {(t[:, mid, 25] - t[:, mid, 24]).mean():.0f} ns")
    if kinds is not None:      # timeline of every kind of op relative to its own mma:rdy0 (means over CTAs and middle ops)
        names = {3: "issAll", 4: "acc0", 8: "pub0", 9: "acc1", 10: "pub1", 11: "acc2", 12: "pub2", 26: "ctaSums", 27: "vecReady",
                 28: "matvec", 14: "gathered", 29: "sideDone", 13: "end"}
        for k in sorted(set(kinds[:-1])):
            sel = [i for i in range(n_ops // 4, 3 * n_ops // 4) if kinds[i] == k]
            items = []
            for e, nm in names.items():
                v = t[:, sel, e]
                if (v > 0).all():
                    items.append((float((v - t[:, sel, 0]).mean()), nm))
            print(f"  timeline {k}: " + "  ".join(f"{nm}={x:.0f}" for x, nm in sorted(items)))
    c = ctas // 2
    for op in range(n_ops // 2, n_ops // 2 + 2):
        base = t[c, op, 0]
        print(f"  CTA {c} op {op}: " + "  ".join(f"{EV[e]}={t[c, op, e] - base:.0f}" for e in range(15) if t[c, op, e] > 0))


def run_long(n=16, depth=24):
    h = w = 48
    x = rnd((n, h, w, 64), seed=3).to(bf)
    ws = [rnd((64, 64, 3, 3), 0.06, seed=100 + i).contiguous() for i in range(depth)]
    bs = [rnd((64,), 0.1, seed=200 + i) for i in range(depth)]
    packs = [ops.PackedWeights() for _ in range(depth)]
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.zeros((depth, n, h, w, 64), dtype=bf, device=DEV)
    ctas = n * 6
    trace = torch.zeros(ctas * depth * 32 + 4096, dtype=torch.int64, device=DEV)

    def build():
        ch = ops.Chain(n, h, w, x.device)
        ch.space(0, A)
        ch.space(1, x.view(1, n, h, w, 64))
        for i in range(depth):
            ch.conv(ops.Chain.ref(1, 0) if i == 0 else ops.Chain.ref(0, i - 1), ops.Chain.ref(0, i), i, bs[i], relu=True)
        return ch
    ch = build()
    for _ in range(3):
        ch.run(bank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ch.run(bank)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    fl = depth * 2.0 * n * h * w * 64 * 64 * 9
    print(f"long chain (cluster={ch.used_cluster}): {us:.1f} us per launch, {us / depth:.2f} us per layer, {fl / us / 1e6:.0f} TFLOP/s")
    trace.zero_()
    ch.run(bank, trace=trace)
    torch.cuda.synchronize()
    tr = trace[:ctas * depth * 32].cpu().numpy().reshape(ctas, depth, 32)
    report(tr, depth, ctas, "24 dependent relu convs on [16,48,48,64]")


def run_group():
    import models
    from srb200.trainer import FlatParams
    torch.manual_seed(0)
    grp = models.rcan.ResidualGroup(64, 3, 16, 1, 20).to(DEV)
    flat = FlatParams(grp)
    flat.begin_step(zero=True)
    n = 16
    x = torch.randn(n, 48, 48, 64, device=DEV).to(bf).requires_grad_(True)
    g = (torch.randn(n, 48, 48, 64, device=DEV) * 0.01).to(bf)
    ctas = n * 6
    n_ops = 41
    for _ in range(2):
        with ops.deferred_wgrads() as q:
            grp(x).backward(g)
            q.items.clear()
    torch.cuda.synchronize()
    for direction in ("forward", "backward"):
        trace = torch.zeros(ctas * n_ops * 32 + 65536, dtype=torch.int64, device=DEV)
        if direction == "forward":
            ops.CHAIN_TRACE = trace
            with torch.no_grad():
                grp(x)
            ops.CHAIN_TRACE = None
        else:
            y = grp(x)
            ops.CHAIN_TRACE = trace
            with ops.deferred_wgrads() as q:
                y.backward(g)
                q.items.clear()
            ops.CHAIN_TRACE = None
        torch.cuda.synchronize()
        tr = trace[:ctas * n_ops * 32].cpu().numpy().reshape(ctas, n_ops, 32)
        if direction == "forward":
            kinds = ["conv1" if i % 2 == 0 else "conv2+CA" for i in range(40)] + ["tail"]
        else:
            kinds = ["tail+CAbwd"] + ["dgrad2+mask" if i % 2 == 0 else "dgrad1+CAbwd" for i in range(40)]
        report(tr, n_ops, ctas, f"RCAN ResidualGroup {direction}", kinds)
    # timings over graph replays
    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    def fwd():
        with torch.no_grad():
            grp(x)

    def fwd_bwd():
        with ops.deferred_wgrads() as q:
            grp(x).backward(g)
            q.items.clear()
    uf = timed(fwd)
    ub = timed(fwd_bwd)
    print(f"group forward {uf:.1f} us, forward+backward {ub:.1f} us (eager launches; includes host overhead)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "long"
    if what == "long":
        run_long()
    else:
        run_group()
