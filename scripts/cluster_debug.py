"""Stage-by-stage check of the per-sample cluster chain kernel (conv_cluster.cu) against the per-layer kernels.
One stage per process (a trap poisons the CUDA context):  python scripts/cluster_debug.py <stage>
Prints where mismatches sit (sample, row band, column tile) instead of just asserting."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from srb200 import lib as L, ops  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16


def rnd(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def where_bad(a, b, name):
    """a, b: [N,H,W,64]"""
    bad = (a != b) | (a.isnan() != b.isnan())
    nb = int(bad.sum())
    print(f"  {name}: {'OK bit-exact' if nb == 0 else 'MISMATCH'}  bad={nb}/{bad.numel()}  rel={rel(a.float(), b.float()):.3e}")
    if nb:
        n, h, w, _ = a.shape
        per = bad.any(dim=3)
        for s in range(min(n, 2)):
            rows = []
            for hb in range(0, h, 16):
                rows.append(" ".join(f"{int(per[s, hb:hb + 16, wb:wb + 8].sum()):3d}" for wb in range(0, w, 8)))
            print(f"    sample {s}: bad pixels per 16x8 tile (rows = bands): " + " | ".join(rows))
        idx = bad.nonzero()[:6]
        for i in idx:
            i = tuple(int(v) for v in i)
            print(f"    first bad {i}: got {float(a[i]):.5f} want {float(b[i]):.5f}")
    return nb == 0


def stage_convs(shape, depth=3):
    n, h, w = shape
    x = rnd((n, h, w, 64), seed=1).to(bf)
    m = rnd((n, h, w, 64), seed=2).to(bf)
    ws = [rnd((64, 64, 3, 3), 0.05, seed=10 + i).contiguous() for i in range(3)]
    bs = [rnd((64,), 0.1, seed=20 + i) for i in range(3)]
    packs = [ops.PackedWeights() for _ in range(3)]
    y1 = torch.empty_like(x)
    ops.conv(x, 0, 64, packs[0], ws[0], bs[0], y1, 0, 64, 3, relu=True)
    y2 = torch.empty_like(x)
    ops.conv(y1, 0, 64, packs[1], ws[1], bs[1], y2, 0, 64, 3, scale=0.5, res=(x, 0))
    y3 = torch.empty_like(x)
    cs = torch.zeros(64, device=DEV)
    ops.conv(y2, 0, 64, packs[2], ws[2], None, y3, 0, 64, 3, mask=(m, 0), colsum=cs, colsum_groups=1)
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.full((3, n, h, w, 64), float("nan"), dtype=bf, device=DEV)
    E = torch.stack([x, m]).contiguous()
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    cs2 = torch.zeros(64, device=DEV)
    ch.conv(ref(1, 0), ref(0, 0), 0, bs[0], relu=True)
    if depth > 1:
        ch.conv(ref(0, 0), ref(0, 1), 1, bs[1], scale=0.5, res=ref(1, 0))
    if depth > 2:
        ch.conv(ref(0, 1), ref(0, 2), 2, None, mask=ref(1, 1), colsum=cs2, colsum_groups=1)
    ch.run(bank)
    torch.cuda.synchronize()
    print(f"convs {shape} depth {depth}: cluster kernel used = {ch.used_cluster}")
    ok = where_bad(A[0], y1, "relu conv")
    if depth > 1:
        ok &= where_bad(A[1], y2, "residual conv x0.5")
    if depth > 2:
        ok &= where_bad(A[2], y3, "masked conv")
        print(f"  colsum rel {rel(cs2, cs):.2e}")
        ok &= rel(cs2, cs) < 1e-5
    return ok


def stage_long(n=16, depth=24, reps=3):
    h = w = 48
    x = rnd((n, h, w, 64), seed=3).to(bf)
    ws = [rnd((64, 64, 3, 3), 0.06, seed=100 + i).contiguous() for i in range(depth)]
    bs = [rnd((64,), 0.1, seed=200 + i) for i in range(depth)]
    packs = [ops.PackedWeights() for _ in range(depth)]
    cur, outs = x, []
    for i in range(depth):
        y = torch.empty_like(x)
        ops.conv(cur, 0, 64, packs[i], ws[i], bs[i], y, 0, 64, 3, relu=True)
        outs.append(y)
        cur = y
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.zeros((depth, n, h, w, 64), dtype=bf, device=DEV)
    ok = True
    for rep in range(reps):
        A.zero_()
        ch = ops.Chain(n, h, w, x.device)
        ch.space(0, A)
        ch.space(1, x.view(1, n, h, w, 64))
        ref = ops.Chain.ref
        for i in range(depth):
            ch.conv(ref(1, 0) if i == 0 else ref(0, i - 1), ref(0, i), i, bs[i], relu=True)
        ch.run(bank)
        torch.cuda.synchronize()
        first_bad = next((i for i in range(depth) if not torch.equal(A[i], outs[i])), None)
        print(f"long chain rep {rep}: cluster={ch.used_cluster} first bad op = {first_bad}")
        if first_bad is not None:
            where_bad(A[first_bad], outs[first_bad], f"op {first_bad}")
            ok = False
    # timing
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, x.view(1, n, h, w, 64))
    for i in range(depth):
        ch.conv(ops.Chain.ref(1, 0) if i == 0 else ops.Chain.ref(0, i - 1), ops.Chain.ref(0, i), i, bs[i], relu=True)
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
    print(f"long chain: {us:.1f} us per launch, {us / depth:.2f} us per layer, {fl / us / 1e6:.0f} TFLOP/s")
    return ok


def stage_ca(shape):
    n, h, w = shape
    cr = 4
    y1 = rnd((n, h, w, 64), seed=5).to(bf)
    x = rnd((n, h, w, 64), seed=6).to(bf)
    w2 = rnd((64, 64, 3, 3), 0.05, seed=8).contiguous()
    b2 = rnd((64,), 0.1, seed=9)
    cw1, cb1 = rnd((cr, 64), 0.2, seed=11), rnd((cr,), 0.1, seed=12)
    cw2, cb2 = rnd((64, cr), 0.2, seed=13), rnd((64,), 0.1, seed=14)
    w3 = rnd((64, 64, 3, 3), 0.05, seed=15).contiguous()
    b3 = rnd((64,), 0.1, seed=16)
    pk, pk3 = ops.PackedWeights(), ops.PackedWeights()
    t = torch.empty_like(x)
    pool = torch.zeros(n, 64, device=DEV)
    ops.conv(y1, 0, 64, pk, w2, b2, t, 0, 64, 3, colsum=pool, colsum_groups=n)
    out = torch.empty_like(x)
    s = torch.empty(n, 64, device=DEV)
    yg = torch.empty(n, 64, device=DEV)
    ops.ca_fwd(t, x, pool, False, cw1, cb1, cw2, cb2, out, s, yg)
    nxt = torch.empty_like(x)
    ops.conv(out, 0, 64, pk3, w3, b3, nxt, 0, 64, 3, relu=True)
    bank = ops.FilterBank().get([(w2, pk), (w3, pk3)], L.PACK_FWD)
    A = torch.zeros((3, n, h, w, 64), dtype=bf, device=DEV)
    E = torch.stack([y1, x]).contiguous()
    pool2 = torch.zeros(n, 64, device=DEV)
    s2 = torch.empty(n, 64, device=DEV)
    yg2 = torch.empty(n, 64, device=DEV)
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    ch.conv_ca(ref(1, 0), ref(0, 0), ref(0, 1), ref(1, 1), 0, b2, pool2, cw1, cb1, cw2, cb2, s2, yg2)
    ch.conv(ref(0, 1), ref(0, 2), 1, b3, relu=True)      # consumes the CA output through the halos
    ch.run(bank)
    torch.cuda.synchronize()
    print(f"CA forward {shape}: cluster={ch.used_cluster}")
    ok = where_bad(A[0], t, "t = conv2 + bias")
    print(f"  pool rel {rel(pool2, pool):.2e}  s rel {rel(s2, s):.2e}  gate rel {rel(yg2, yg):.2e}  out rel {rel(A[1], out):.2e}  "
          f"next conv rel {rel(A[2], nxt):.2e}")
    ok &= rel(pool2, pool) < 1e-5 and rel(s2, s) < 1e-5 and rel(yg2, yg) < 1e-5 and rel(A[1], out) < 2e-3 and rel(A[2], nxt) < 4e-3
    return ok


def stage_rcab(shape, blocks=2):
    """conv1+ReLU -> conv2+CALayer+skip, `blocks` times, then a plain conv — against the per-layer kernels."""
    n, h, w = shape
    cr = 4
    x = rnd((n, h, w, 64), seed=6).to(bf)
    ws = [rnd((64, 64, 3, 3), 0.05, seed=50 + i).contiguous() for i in range(2 * blocks + 1)]
    bs = [rnd((64,), 0.1, seed=70 + i) for i in range(2 * blocks + 1)]
    cas = [(rnd((cr, 64), 0.2, seed=90 + 4 * i), rnd((cr,), 0.1, seed=91 + 4 * i), rnd((64, cr), 0.2, seed=92 + 4 * i),
            rnd((64,), 0.1, seed=93 + 4 * i)) for i in range(blocks)]
    packs = [ops.PackedWeights() for _ in ws]
    cur = x
    want = []
    for b in range(blocks):
        r = torch.empty_like(x)
        ops.conv(cur, 0, 64, packs[2 * b], ws[2 * b], bs[2 * b], r, 0, 64, 3, relu=True)
        t = torch.empty_like(x)
        pool = torch.zeros(n, 64, device=DEV)
        ops.conv(r, 0, 64, packs[2 * b + 1], ws[2 * b + 1], bs[2 * b + 1], t, 0, 64, 3, colsum=pool, colsum_groups=n)
        out = torch.empty_like(x)
        s = torch.empty(n, 64, device=DEV)
        yg = torch.empty(n, 64, device=DEV)
        ops.ca_fwd(t, cur, pool, False, *cas[b], out, s, yg)
        want.append((r, t, out, pool, s, yg))
        cur = out
    last = torch.empty_like(x)
    ops.conv(cur, 0, 64, packs[-1], ws[-1], bs[-1], last, 0, 64, 3, relu=True)
    bank = ops.FilterBank().get(list(zip(ws, packs)), L.PACK_FWD)
    A = torch.zeros((3 * blocks + 1, n, h, w, 64), dtype=bf, device=DEV)
    pools = torch.zeros(blocks, n, 64, device=DEV)
    ss = torch.empty(blocks, n, 64, device=DEV)
    ys = torch.empty(blocks, n, 64, device=DEV)
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, x.view(1, n, h, w, 64))
    ref = ops.Chain.ref
    c = ref(1, 0)
    for b in range(blocks):
        ch.conv(c, ref(0, 3 * b), 2 * b, bs[2 * b], relu=True)
        ch.conv_ca(ref(0, 3 * b), ref(0, 3 * b + 1), ref(0, 3 * b + 2), c, 2 * b + 1, bs[2 * b + 1], pools[b], *cas[b], ss[b], ys[b])
        c = ref(0, 3 * b + 2)
    ch.conv(c, ref(0, 3 * blocks), 2 * blocks, bs[-1], relu=True)
    ch.run(bank)
    torch.cuda.synchronize()
    print(f"RCAB forward {shape} x{blocks}: cluster={ch.used_cluster}")
    ok = True
    for b in range(blocks):
        r, t, out, pool, s, yg = want[b]
        if b == 0:
            ok &= where_bad(A[0], r, "block 0 relu conv")
            ok &= where_bad(A[1], t, "block 0 t = conv2 + bias")
        e = dict(r=rel(A[3 * b], r), t=rel(A[3 * b + 1], t), out=rel(A[3 * b + 2], out), pool=rel(pools[b], pool), s=rel(ss[b], s),
                 gate=rel(ys[b], yg))
        print(f"  block {b}: " + "  ".join(f"{k} rel {v:.2e}" for k, v in e.items()))
        ok &= e["s"] < 5e-3 and e["gate"] < 1e-3 and e["out"] < 4e-3 and e["t"] < 4e-3
    e_last = rel(A[3 * blocks], last)
    print(f"  last conv rel {e_last:.2e}")
    return ok and e_last < 6e-3


def stage_cabwd(shape):
    n, h, w = shape
    cr = 4
    x = rnd((n, h, w, 64), 0.02, seed=31).to(bf)
    res = rnd((n, h, w, 64), 0.02, seed=32).to(bf)
    t = rnd((n, h, w, 64), seed=33).to(bf)
    m = rnd((n, h, w, 64), seed=41).to(bf)
    wt = rnd((64, 64, 3, 3), 0.05, seed=34).contiguous()
    w3 = rnd((64, 64, 3, 3), 0.05, seed=42).contiguous()
    cw1, cb1 = rnd((cr, 64), 0.2, seed=35), rnd((cr,), 0.1, seed=36)
    cw2, cb2 = rnd((64, cr), 0.2, seed=37), rnd((64,), 0.1, seed=38)
    s = rnd((n, 64), 0.3, seed=39)
    yg = torch.sigmoid(rnd((n, 64), seed=40))
    pk, pk3 = ops.PackedWeights(), ops.PackedWeights()
    g = torch.empty_like(x)
    ops.conv(x, 0, 64, pk, wt, None, g, 0, 64, 3, res=(res, 0))
    dt = torch.empty_like(x)
    gr = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2 = torch.zeros(64, device=DEV)
    ops.ca_bwd(g, t, s, yg, cw1, cb1, cw2, cb2, dt, gr[0], gr[1], gr[2], gr[3], db2, torch.zeros(n, 64, device=DEV),
               accumulate=True, scratch_is_zero=True)
    d1 = torch.empty_like(x)
    cs = torch.zeros(64, device=DEV)
    ops.conv(dt, 0, 64, pk3, w3, None, d1, 0, 64, 3, mask=(m, 0), colsum=cs, colsum_groups=1)
    bank = ops.FilterBank().get([(wt, pk), (w3, pk3)], L.PACK_FWD)
    A = torch.zeros((3, n, h, w, 64), dtype=bf, device=DEV)
    E = torch.stack([x, res, t, m]).contiguous()
    gr2 = [torch.zeros_like(p) for p in (cw1, cb1, cw2, cb2)]
    db2c = torch.zeros(64, device=DEV)
    cs2 = torch.zeros(64, device=DEV)
    ch = ops.Chain(n, h, w, x.device)
    ch.space(0, A)
    ch.space(1, E)
    ref = ops.Chain.ref
    ch.conv(ref(1, 0), ref(0, 0), 0, None, res=ref(1, 1),
            ca_bwd=dict(t=ref(1, 2), dt=ref(0, 1), w1=cw1, b1=cb1, w2=cw2, b2=cb2, s=s, y=yg, dw1=gr2[0], db1=gr2[1],
                        dw2=gr2[2], db2=gr2[3], scratch=torch.zeros(n, 64, device=DEV), colsum_dt=db2c))
    ch.conv(ref(0, 1), ref(0, 2), 1, None, mask=ref(1, 3), colsum=cs2, colsum_groups=1)
    ch.run(bank)
    torch.cuda.synchronize()
    print(f"CA backward fused {shape}: cluster={ch.used_cluster}")
    ok = where_bad(A[0], g, "g = dgrad + residual")
    print(f"  dt rel {rel(A[1], dt):.2e}  db2 rel {rel(db2c, db2):.2e}  next masked conv rel {rel(A[2], d1):.2e} colsum rel {rel(cs2, cs):.2e}  "
          + " ".join(f"{rel(a, b):.1e}" for a, b in zip(gr2, gr)))
    ok &= rel(A[1], dt) < 2e-3 and rel(db2c, db2) < 2e-3 and rel(A[2], d1) < 6e-3 and all(rel(a, b) < 1e-3 for a, b in zip(gr2, gr))
    return ok


def stage_model(name, n=2):
    import models
    torch.manual_seed(0)
    if name == "rcan":
        kw = dict(n_feats=64, reduction=16, scale_factor=4, n_resblocks=20, n_resgroups=2)
        cls = models.RCAN
    else:
        kw = dict(n_feats=64, scale_factor=4, n_resblocks=16, res_scale=0.1)
        cls = models.EDSR
    m0 = cls(**kw)
    sd = {k: v.clone() for k, v in m0.state_dict().items()}
    x = torch.rand(n, 3, 48, 48)
    hr = torch.rand(n, 3, 192, 192)
    res = {}
    for mode in ("cluster", "flags", "layers"):
        os.environ["SRB200_NO_CHAIN"] = "1" if mode == "layers" else "0"
        os.environ["SRB200_CHAIN_CLUSTER"] = "1" if mode == "cluster" else "0"
        os.environ["SRB200_CHAIN_FWD"] = "cluster" if mode == "cluster" else "flags"      # (the defaults pick per direction)
        os.environ["SRB200_CHAIN_BWD"] = "cluster" if mode == "cluster" else "flags"
        m = cls(**kw)
        m.load_state_dict(sd)
        m.compute_dtype = "bf16"
        m = m.to(DEV)
        out = m.training_step({"lr": x.to(DEV), "hr": hr.to(DEV)}, 0)
        out["loss"].backward()
        torch.cuda.synchronize()
        with torch.no_grad():
            sr = m.forward(x.to(DEV)).float().cpu()
        res[mode] = (sr, out["loss"].item(), {k: p.grad.double().cpu() for k, p in m.named_parameters() if p.requires_grad})
    ok = True
    for other in ("flags", "layers"):
        e_out = rel(res["cluster"][0], res[other][0])
        num = sum(float(((res["cluster"][2][k] - v) ** 2).sum()) for k, v in res[other][2].items())
        den = sum(float((v ** 2).sum()) for v in res[other][2].values())
        worst = max(((rel(res["cluster"][2][k], v), k) for k, v in res[other][2].items()))
        print(f"{name} cluster vs {other}: out rel {e_out:.2e}  loss {res['cluster'][1]:.6f} vs {res[other][1]:.6f}  "
              f"grad global {(num / den) ** 0.5:.2e}  worst tensor {worst[0]:.2e} ({worst[1]})")
        ok &= e_out < 5e-3 and (num / den) ** 0.5 < 2e-2
    return ok


STAGES = {
    "c1": lambda: stage_convs((1, 16, 24), 1),
    "c1x3": lambda: stage_convs((1, 16, 24), 3),
    "c2h": lambda: stage_convs((1, 16, 48), 3),
    "c2v": lambda: stage_convs((1, 32, 24), 3),
    "c4": lambda: stage_convs((3, 32, 48), 3),
    "c6": lambda: stage_convs((2, 48, 48), 3),
    "c6b": lambda: stage_convs((16, 48, 48), 3),
    "c8": lambda: stage_convs((2, 64, 48), 3),
    "long": lambda: stage_long(),
    "ca1": lambda: stage_ca((2, 16, 24)),
    "ca6": lambda: stage_ca((16, 48, 48)),
    "cab1": lambda: stage_cabwd((2, 16, 24)),
    "cab6": lambda: stage_cabwd((16, 48, 48)),
    "rcab1": lambda: stage_rcab((2, 16, 24)),
    "rcab2": lambda: stage_rcab((2, 32, 48)),
    "rcab6": lambda: stage_rcab((16, 48, 48), 3),
    "rcan": lambda: stage_model("rcan"),
    "edsr": lambda: stage_model("edsr"),
}

if __name__ == "__main__":
    name = sys.argv[1]
    ok = STAGES[name]()
    print(f"STAGE {name}: {'PASS' if ok else 'FAIL'}")
    sys.exit(0 if ok else 1)
