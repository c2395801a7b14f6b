#!/usr/bin/env python
"""bench.py — RCAN x4 training throughput on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model rcan|edsr|rdn]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = forward + L1 loss + backward (+ gradient all-reduce when N>1) + Adam over one batch of
16 synthetic 48x48 LR patches per GPU (weak scaling), i.e. SRModel.training_step + optimizer.step
of the reference (models/srmodel.py:160-171,145-154) on config 3 of BASELINE.json:
RCAN(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4), bf16 tcgen05 path.

  value     patches/s over all ranks, inputs resident in HBM, CUDA-graph replay, CUDA events,
            max over ranks
  e2e       same through the public API with pinned HOST batches: H2D copy of lr+hr and D2H read
            of the loss inside the timed region every step
  roofline  the dominant kernel (tcgen05 3x3 conv 64->64 on [16,48,48,64] bf16) timed alone
            with CUDA events; algorithmic FLOPs = 2*N*H*W*Cout*Cin*9
  cpu_baseline  the oracle port (torch CPU fp32, same model) on a bounded sample, rank 0, N=1
  --impl reference  times that CPU port with all host threads and prints the same line shape
  --impl library    (informative, not a driver arm) the same torch ops on the GPU through stock PyTorch / cuDNN
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sr-pytorch-lightning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one conv_chain_kernel launch (ncu --set full,
# profiles/r01_ncu_conv_chain_v8.txt: 12.8 MB read + 232.2 MB written); None until that capture exists
CHAIN_TRAFFIC_BYTES = 245.0e6

METRIC = "RCAN x4 train patches/sec (16x 48x48 LR patches per GPU per step, fwd+L1+bwd+Adam)"


def metric_name(model_key):
    return METRIC.replace("RCAN", MODEL_CFG[model_key][0])
UNIT = "patches/s"
BATCH, LR = 16, 48

MODEL_CFG = {
    "rcan": ("RCAN", dict(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4), 220.04),
    "edsr": ("EDSR", dict(n_feats=64, n_resblocks=16, res_scale=1.0, scale_factor=4), 27.41),
    "rdn": ("RDN", dict(rdn_config="B", scale_factor=4), 314.20),
}   # last entry: algorithmic training GFLOP per patch


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d["bf16_tflops_sustained"]),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    except Exception:  # noqa: BLE001
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference's own path re-stated functionally, oracle/sr_oracle.py)
# ------------------------------------------------------------------------------------------------
def cpu_port_step_fn(model_key: str, batch: int, device="cpu", autocast_bf16=False):
    from oracle import sr_oracle
    import models
    cls, kw, _ = MODEL_CFG[model_key]
    torch.manual_seed(0)
    ref_shapes = getattr(models, cls)(**kw).state_dict()
    sd = {}
    params = []
    for k, v in ref_shapes.items():
        t = v.detach().clone().float().to(device)
        if not k.startswith(("sub_mean", "add_mean")):
            t.requires_grad_(True)
            params.append(t)
        sd[k] = t
    opt = torch.optim.Adam(params, lr=1e-3)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(batch, 3, LR, LR, generator=g).to(device)
    hr = torch.rand(batch, 3, LR * 4, LR * 4, generator=g).to(device)
    cfg = {"scale": 4}
    if cls == "RCAN":
        cfg.update(n_resblocks=kw["n_resblocks"], n_resgroups=kw["n_resgroups"])
    elif cls == "EDSR":
        cfg.update(n_resblocks=kw["n_resblocks"], res_scale=kw["res_scale"])
    else:
        cfg.update(rdn_config=kw["rdn_config"])
    fwd = sr_oracle.FORWARDS[cls]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
            sr = fwd(x, sd, **cfg)
        loss = sr_oracle.l1_loss(sr.float(), hr)
        loss.backward()
        opt.step()
        return loss.item()
    return step


def time_cpu_port(model_key: str, batch: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_port_step_fn(model_key, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size the per-step sample so that (warmup + steps) steps take ~2 minutes
    pps1, ms1 = time_cpu_port(args.model, 1, 1, 1)
    budget_s = 110.0
    per_step = budget_s / max(1, args.steps + args.warmup)
    b = int(max(1, min(BATCH, per_step / (ms1 / 1e3))))
    pps, ms = time_cpu_port(args.model, b, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": metric_name(args.model), "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL_CFG[args.model][0]} x4 train step, CPU port of the reference path "
                               f"(oracle/sr_oracle.py, torch {torch.__version__} CPU fp32), {b} of 16 patches per step",
                   "sample_batch": b},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {b} patches (48x48 LR), fwd+L1+bwd+Adam"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_library_arm(args):
    """NOT a driver arm (informative only, `--impl library`): the same torch ops the reference modules call, run on
    the GPU through stock PyTorch / cuDNN — fp32 (TF32 off) and bf16 autocast — i.e. what the reference itself would
    reach on this B200 (SURVEY §8d "library bar").  One GPU, eager launches, CUDA events."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    out = {"impl": "library", "metric": metric_name(args.model), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": max(3, args.warmup), "higher_is_better": True, "dtype": "f32 / bf16 autocast", "data": "synthetic",
           "config": {"workload": f"{MODEL_CFG[args.model][0]} x4 train step, oracle port on cuda:0 through torch "
                                  f"{torch.__version__} / cuDNN {torch.backends.cudnn.version()}, eager"}}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        step = cpu_port_step_fn(args.model, BATCH, device="cuda", autocast_bf16=ac)
        for _ in range(max(3, args.warmup)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"value": BATCH / (ms / 1e3), "ms_per_step": ms}
    out["value"] = out["bf16_autocast"]["value"]
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.tmp.read().splitlines():
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def time_dominant_kernel(dev, reps=10):
    """The kernel that carries the step: conv_chain_kernel, one persistent launch per RCAN ResidualGroup
    and direction (20 RCAB + tail conv = 41 tcgen05 3x3 64->64 convs on [16,48,48,64] bf16, CALayer
    forward / backward fused in).  Timed alone with CUDA events over graph replays; the activations of
    a launch (3.1 GB of arena per direction across the 10 groups of a step; 289 MB per launch) exceed
    the 126 MB L2, and `reps` distinct input/arena sets rotate.  Returns (us forward launch, us backward
    launch, algorithmic FLOP per launch)."""
    import models
    from srb200 import ops
    from srb200.trainer import FlatParams
    torch.manual_seed(0)
    grp = models.rcan.ResidualGroup(64, 3, 16, 1, 20).to(dev)
    # gradients land in a flat buffer that is "live" (accumulating), as inside a training step: without it
    # every small gradient tensor would get its own fill launch (~120 per group) and those, not the chain
    # launch, would be timed (r01 v4 lines reported 826 us for the 336 us backward launch for this reason)
    flat = FlatParams(grp)
    flat.begin_step(zero=True)
    arena = ops.ZeroArena(dev)
    bf = torch.bfloat16
    xs = [torch.randn(BATCH, LR, LR, 64, device=dev).to(bf).requires_grad_(True) for _ in range(reps)]
    gs = [(torch.randn(BATCH, LR, LR, 64, device=dev) * 0.01).to(bf) for _ in range(reps)]

    def timed(fn0):
        def fn():
            ops.set_arena(arena)
            try:
                arena.reset()                     # one memset per graph replay (pooled sums / CA scratch)
                fn0()
            finally:
                ops.set_arena(None)
        fn()
        fn()                                      # second pass: the arena is sized now
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph):
                fn()
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    def fwd():
        with torch.no_grad():
            for x in xs:
                grp(x)

    def fwd_bwd():
        # weight-gradient launches are queued and dropped: only the chain launches (+ one 5 us skip add) run
        with ops.deferred_wgrads() as q:
            for x, g in zip(xs, gs):
                grp(x).backward(g)
                q.items.clear()
    us_f = timed(fwd)
    us_fb = timed(fwd_bwd)
    flat.detach()
    flop = 41 * 2.0 * BATCH * LR * LR * 64 * 64 * 9
    return us_f, us_fb - us_f, flop


def run_ours(args):
    import torch.distributed as dist
    import models
    from srb200 import lib as L
    from srb200.trainer import TrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L.load()

    cls, kw, gflop_patch = MODEL_CFG[args.model]
    torch.manual_seed(0)                      # identical initial weights on every rank
    model = getattr(models, cls)(**kw)
    model.compute_dtype = "bf16"
    model = model.to(dev)
    step = TrainStep(model, (BATCH, 3, LR, LR), 4, lr=1e-3, use_graph=not args.no_graph)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    nb = 4
    host_lr = [torch.rand(BATCH, 3, LR, LR, generator=g).pin_memory() for _ in range(nb)]
    host_hr = [torch.rand(BATCH, 3, LR * 4, LR * 4, generator=g).pin_memory() for _ in range(nb)]
    dev_lr = [t.to(dev) for t in host_lr]
    dev_hr = [t.to(dev) for t in host_hr]
    step.load_batch(dev_lr[0], dev_hr[0])
    torch.cuda.reset_peak_memory_stats(dev)
    step.capture()
    peak_mem = torch.cuda.max_memory_allocated(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up -------------------------------------------------------------------------------
    for i in range(max(3, args.warmup)):
        step.step(dev_lr[i % nb], dev_hr[i % nb])
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = L.launch_count()
    barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step.step(dev_lr[i % nb], dev_hr[i % nb])
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = step.launches_per_step * args.steps if step.graph is not None else L.launch_count() - c0
    loss_dev = float(step.loss.item())

    # ---- timed region 2: end to end from pinned host memory ------------------------------------
    h2d = host_lr[0].numel() * 4 + host_hr[0].numel() * 4
    barrier()
    torch.cuda.synchronize()
    e0.record()
    last = 0.0
    for i in range(args.steps):
        loss = step.step(host_lr[i % nb], host_hr[i % nb])
        last = loss.item()                    # D2H read of the step's result, every step
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None

    def finish():
        """Leave together: every rank meets at one last barrier (rank 0 arrives after its rank-0-only
        measurements), then exits without running NCCL's teardown — a rank that tears its communicator
        down while a peer is still busy can block both (seen as a torchrun that never returns)."""
        if world > 1:
            sys.stdout.flush()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    k_us_f, k_us_b, k_flop = time_dominant_kernel(dev)
    k_us = 0.5 * (k_us_f + k_us_b)             # a step launches it 10x forward + 10x backward
    k_tflops = k_flop / (k_us * 1e-6) / 1e12
    ms_step = ms_total / args.steps
    value = BATCH * world * args.steps / (ms_total / 1e3)
    e2e_value = BATCH * world * args.steps / (ms_e2e / 1e3)
    step_tflops = gflop_patch * BATCH / (ms_step / 1e3) / 1e3      # per GPU
    line = {
        "metric": metric_name(args.model), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": f"{cls} x4 training step (BASELINE.json configs[2]): {kw}, batch {BATCH} x 3x{LR}x{LR} LR per GPU, "
                        f"L1 loss, Adam lr=1e-3, bf16 activations / fp32 accumulate+master weights; "
                        f"64-channel trunks run as layer-chain launches (srb_conv_chain){'' if os.environ.get('SRB200_NO_CHAIN', '0') in ('', '0') else ' [disabled]'}",
            "parallelism": f"dp{world}", "cuda_graph": step.graph is not None,
            "l2": f"no explicit flush: one step streams {peak_mem / 2**30:.2f} GiB of saved activations and gradients "
                  f"(>> 126 MB L2); 4 distinct input batches rotate",
            "loss_last": loss_dev,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss_last": last},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": k_tflops, "peak": peaks["burst"], "unit": "TFLOP/s",
                     "frac": k_tflops / peaks["burst"], "traffic": CHAIN_TRAFFIC_BYTES,
                     "kernel": "conv_chain_kernel: one RCAN ResidualGroup (20 RCAB + conv = 41 tcgen05 3x3 64->64 convs, "
                               "CALayer fused) per launch on [16,48,48,64] bf16; average of the forward and the backward "
                               "launch, timed alone over graph replays of 10 rotating arena sets (2.9 GB > L2)",
                     "us_per_launch": k_us, "us_forward_launch": k_us_f, "us_backward_launch": k_us_b,
                     "flop_per_launch": k_flop, "launches_per_step": 20,
                     "share_of_step": (10.0 * (k_us_f + k_us_b) / (ms_step * 1e3)) if args.model == "rcan" else None,
                     "traffic_note": "dram__bytes_read+write per launch from profiles/r01_ncu_conv_chain_v8.txt",
                     "peak_source": peaks["source"]},
        "roofline_step": {"bound": "tensor", "achieved": step_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                          "frac": step_tflops / peaks["sustained"],
                          "note": f"whole step: {gflop_patch} algorithmic GFLOP/patch x {BATCH} / ms_per_step, vs sustained bf16 peak"},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            pps, ms = time_cpu_port(args.model, BATCH, 5, 1)
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"5 steps x {BATCH} patches (48x48 LR) after 1 warm-up (~{6 * ms / 1e3:.0f} s of CPU work), "
                                              "fwd+L1+bwd+Adam, torch CPU fp32, all host threads"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "library"])
    ap.add_argument("--model", default="rcan", choices=list(MODEL_CFG))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "library":
        run_library_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
