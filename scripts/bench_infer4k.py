"""EDSR x4 large (32 ResBlocks, 256 ch, res_scale 0.1) 960x540 -> 3840x2160 forward, frames/s
(BASELINE.json config 5).  Single GPU: whole frame through model.forward.  N GPUs (torchrun):
row strips with per-layer NCCL halo exchange (srb200/tiled.py).  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sr-pytorch-lightning_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--h", type=int, default=540)
    ap.add_argument("--w", type=int, default=960)
    ap.add_argument("--local-parts", type=int, default=0, help="single GPU: chop into this many strips")
    args = ap.parse_args()
    import models
    from srb200.tiled import DistExchange, LocalExchange, TiledEDSR
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    m = models.EDSR(n_feats=256, n_resblocks=32, res_scale=0.1, scale_factor=4)
    m.compute_dtype = "bf16"
    m = m.cuda()
    x = torch.rand(1, 3, args.h, args.w, generator=torch.Generator().manual_seed(0)).cuda()
    if world > 1:
        runner = TiledEDSR(m, DistExchange())
        fn = lambda: runner.forward(x)  # noqa: E731
    elif args.local_parts > 1:
        runner = TiledEDSR(m, LocalExchange(args.local_parts))
        fn = lambda: runner.forward(x)  # noqa: E731
    else:
        fn = lambda: m.forward(x)  # noqa: E731
    with torch.no_grad():
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        gflop = 231.5644 * (args.h * args.w) / (48 * 48)
        print(json.dumps({"metric": "EDSR-large x4 960x540->3840x2160 forward frames/sec", "value": 1e3 / ms, "unit": "frames/s",
                          "n_gpus": world, "ms_per_frame": ms, "tflops": gflop / ms, "local_parts": args.local_parts,
                          "dtype": "bf16", "data": "synthetic"}))
    if world > 1:
        sys.stdout.flush()
        dist.barrier()
        os._exit(0)          # leave together, without NCCL teardown ordering hazards


if __name__ == "__main__":
    main()
