"""N>1 host logic on CPU with gloo, world_size 2 (no GPU): row-strip partition + per-layer halo
exchange reproduce the full-frame convolution stack; flat-gradient all-reduce + 1/world scaling
reproduces the gradient mean of DDP."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _strip_stack(x_full, weights, exchange, parts, my, t=1):
    """Reference emulation of TiledEDSR's buffer discipline with torch CPU convs (NHWC buffers)."""
    from srb200.tiled import partition_rows
    H = x_full.shape[2]
    rng = partition_rows(H, parts)
    bufs = []
    for i in my:
        r0, r1 = rng[i]
        buf = torch.zeros(1, (r1 - r0) + 2 * t, x_full.shape[3], x_full.shape[1], dtype=x_full.dtype)
        lo, hi = max(r0 - t, 0), min(r1 + t, H)
        buf[:, lo - (r0 - t): lo - (r0 - t) + hi - lo] = x_full[:, :, lo:hi].permute(0, 2, 3, 1)
        bufs.append(buf)
    for w in weights:
        nxt = []
        for b in bufs:
            y = F.conv2d(b.permute(0, 3, 1, 2), w, padding=1).relu()
            nxt.append(y.permute(0, 2, 3, 1).contiguous())
        exchange(nxt, t)
        bufs = nxt
    return {i: b[:, t:b.shape[1] - t] for i, b in zip(my, bufs)}, rng


def test_partition_rows():
    from srb200.tiled import partition_rows
    p = partition_rows(540, 8)
    assert p[0] == (0, 68) and p[-1][1] == 540 and sum(b - a for a, b in p) == 540
    assert {b - a for a, b in p} == {67, 68}
    assert partition_rows(5, 5) == [(i, i + 1) for i in range(5)]


def test_local_exchange_matches_full_frame():
    from srb200.tiled import LocalExchange
    torch.manual_seed(0)
    x = torch.rand(1, 4, 23, 9, dtype=torch.float64)
    ws = [torch.randn(4, 4, 3, 3, dtype=torch.float64) * 0.3 for _ in range(4)]
    full = x
    for w in ws:
        full = F.conv2d(full, w, padding=1).relu()
    ex = LocalExchange(3)
    out, rng = _strip_stack(x, ws, ex.exchange, 3, ex.strips())
    got = torch.cat([out[i] for i in range(3)], dim=1).permute(0, 3, 1, 2)
    assert torch.allclose(got, full, rtol=0, atol=1e-12)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from srb200.tiled import DistExchange
        torch.manual_seed(0)
        x = torch.rand(1, 4, 21, 9, dtype=torch.float64)
        ws = [torch.randn(4, 4, 3, 3, dtype=torch.float64) * 0.3 for _ in range(3)]
        full = x
        for w in ws:
            full = F.conv2d(full, w, padding=1).relu()
        ex = DistExchange()
        out, rng = _strip_stack(x, ws, ex.exchange, world, ex.strips())
        r0, r1 = rng[rank]
        ok_tile = torch.allclose(out[rank].permute(0, 3, 1, 2), full[:, :, r0:r1], rtol=0, atol=1e-12)
        # DDP semantics of the flat gradient buffer: SUM all-reduce, then 1/world in the optimizer
        g = torch.full((10,), float(rank + 1))
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ok_grad = torch.allclose(g / world, torch.full((10,), sum(range(1, world + 1)) / world))
        q.put((rank, bool(ok_tile), bool(ok_grad)))
    finally:
        dist.destroy_process_group()


def test_dist_exchange_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def _bucket_worker(rank, world, port, q):
    """TrainStep's bucket bookkeeping with real gloo all-reduces on a CPU flat buffer: buckets reduced early (as the
    weight-gradient side stream does per ResidualGroup) + _reduce_rest for the complement = one all-reduce of the whole
    buffer, every element exactly once."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from srb200.trainer import TrainStep
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 8, 3), torch.nn.Conv2d(8, 5, 3), torch.nn.Conv2d(5, 3, 1))
        ts = TrainStep(net, (2, 3, 8, 8), 1)
        ts._bucket_groups = 1
        assert ts.world == world
        n = ts.flat.numel
        ts.flat.grad.copy_(torch.arange(n, dtype=torch.float32) * (rank + 1))
        base = ts.flat.grad.data_ptr()
        # two disjoint buckets in "backward order" with unaligned edges; the complement is three separate ranges
        o = ts.flat.offsets
        ts._reduce_bucket(base + 4 * o[4], base + 4 * n)
        ts._reduce_bucket(base + 4 * (o[2] + 1), base + 4 * (o[4] - 3))
        covered = list(ts._reduced)
        ts._reduce_rest()
        want = torch.arange(n, dtype=torch.float32) * sum(range(1, world + 1))
        q.put((rank, covered == [(o[4], n), (o[2] + 1, o[4] - 3)], bool(torch.equal(ts.flat.grad, want)), ts._reduced == []))
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_cover_the_flat_buffer_exactly_once_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True, True), (1, True, True, True)]
