"""Strip-parallel EDSR forward with the peer-memory halo exchange (srb200.tiled.PeerExchange) under torchrun:
  torchrun --nproc-per-node N scripts/tiled_peer_check.py [small|large] [frames]
Every rank compares its SR rows with the same rows of the untiled forward on its own GPU (bit-identical expected), eager
and as a CUDA graph, and rank 0 prints the frame rate (max over ranks, CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    import models
    from srb200.tiled import PeerExchange, TiledEDSR, partition_rows
    torch.manual_seed(0)
    if what == "large":
        kw, (H, W) = dict(n_feats=256, n_resblocks=32, res_scale=0.1, scale_factor=4), (540, 960)
    else:
        kw, (H, W) = dict(n_feats=64, n_resblocks=4, res_scale=1.0, scale_factor=4), (70, 96)
    m = models.EDSR(**kw)
    m.compute_dtype = "bf16"
    m = m.to(f"cuda:{local}").eval()
    g = torch.Generator().manual_seed(1)
    xs = [torch.rand(1, 3, H, W, generator=g).cuda() for _ in range(2)]
    ok = True
    with torch.no_grad():
        want = [m.forward(x) for x in xs]
        ex = PeerExchange()
        runner = TiledEDSR(m, ex)
        r0, r1 = partition_rows(H, world)[rank]
        s = kw["scale_factor"]
        runner.prepare(xs[0], use_graph=False)
        for i in (0, 1, 0):
            got = runner.run(xs[i])[rank]
            same = torch.equal(got, want[i][:, :, r0 * s:r1 * s])
            ok &= same
            if not same:
                d = (got - want[i][:, :, r0 * s:r1 * s]).abs()
                print(f"rank {rank} eager frame {i}: MISMATCH max {float(d.max()):.3e}, bad rows {d.amax(dim=(0, 1, 3)).nonzero().flatten()[:8].tolist()}")
        runner.prepare(xs[0], use_graph=True)
        for i in (1, 0, 1, 1):
            got = runner.run(xs[i])[rank]
            same = torch.equal(got, want[i][:, :, r0 * s:r1 * s])
            ok &= same
            if not same:
                print(f"rank {rank} graph frame {i}: MISMATCH")
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(frames):
            runner.run(xs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / frames], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"PEER {what} world={world}: {'bit-identical to the untiled forward' if int(flag.item()) else 'MISMATCH'}; "
                  f"{float(ms.item()):.3f} ms/frame = {1e3 / float(ms.item()):.1f} frames/s (graph replay, max over ranks)")
        ex.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
