#!/usr/bin/env python
"""bench.py — BASELINE.json metric on B200, one JSON line:
"RCAN x4 train patches/sec/GPU at 1/2/4/8 B200; 4K-output fwd frames/sec".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|library]
                  [--model rcan|edsr|rdn|srcnn] [--workload all|train|infer4k]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Training half (the line's metric/value).  A "step" = forward + L1 loss + backward (+ gradient all-reduce when N>1) +
Adam over one batch of 16 synthetic 48x48 LR patches per GPU (weak scaling), i.e. SRModel.training_step +
optimizer.step of the reference (models/srmodel.py:160-171,145-154) on BASELINE.json configs[2]:
RCAN(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4), bf16 tcgen05 path.

  value     patches/s over all ranks, inputs resident in HBM, CUDA-graph replay, CUDA events, max over ranks
  e2e       same through the public API with pinned HOST batches: H2D copy of lr+hr and D2H read of the loss inside
            the timed region every step
  roofline  the dominant kernel of --model, timed alone with CUDA events (algorithmic FLOPs = 2*N*H*W*Cout*Cin*k*k)
  sustained the same step replayed for >= 3 s (clocks / power settle), with its own clock samples
  cpu_baseline  the oracle port (torch CPU fp32, same model) on a bounded sample, rank 0, N=1
  library / eager_plugin  (N=1) the stock PyTorch/cuDNN port on this GPU, and this repo's modules driven eagerly the
            way Lightning would (training_step + torch.optim.Adam, no CUDA graph, no flat buffers)
  configs0_srcnn  BASELINE.json configs[0]: SRCNN x2 fwd+bwd, batch 16 of 48x48, on the host cores and on this GPU

4K half (key "infer4k").  EDSR x4 large (256 ch, 32 ResBlocks, res_scale 0.1, BASELINE.json configs[4]),
960x540 -> 3840x2160, bf16, the reference's predict path (srmodel.py:375-380: forward + clamp); at N>1 the frame is
split into row strips with per-layer halo exchange (srb200/tiled.py).  Timed >= 3 s; carries its own clocks,
roofline (conv_wide_kernel<128>), e2e (6 MB H2D frame + 100 MB D2H output per frame) and cpu_baseline.

  --impl reference  times the reference's CPU path (oracle port, all host threads) and prints the same line shape
  --impl library    the stock PyTorch / cuDNN port on the GPU only (informative)
"""
from __future__ import annotations

import argparse
import json
import math
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

# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/);
# None where no capture exists
TRAFFIC = {
    "chain_cluster": None,      # filled from profiles/r02_ncu_chain_cluster.txt when captured
    "chain_flags": 245.0e6,     # profiles/r01_ncu_conv_chain_v8.txt: 12.8 MB read + 232.2 MB written
    "conv_wide128": None,
    "conv_wide64": None,
}
TRAFFIC_NOTE = {
    "chain_flags": "profiles/r01_ncu_conv_chain_v8.txt",
}
try:   # measured values are kept beside the profiles so that bench.py needs no edit after a capture
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as _f:
        for _k, _v in json.load(_f).items():
            TRAFFIC[_k] = _v["bytes"]
            TRAFFIC_NOTE[_k] = _v["source"]
except Exception:  # noqa: BLE001
    pass

UNIT = "patches/s"
BATCH, LR = 16, 48

# class, ctor kwargs (this repo's plugin API), oracle kwargs, scale, algorithmic training GFLOP per patch,
# BASELINE.json configs index
MODEL_CFG = {
    "rcan": ("RCAN", dict(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4),
             dict(n_resblocks=20, n_resgroups=10, scale=4), 4, 220.04, 2),
    "edsr": ("EDSR", dict(n_feats=64, n_resblocks=16, res_scale=1.0, scale_factor=4),
             dict(n_resblocks=16, res_scale=1.0, scale=4), 4, 27.41, 1),
    "rdn": ("RDN", dict(rdn_config="B", scale_factor=4), dict(rdn_config="B", scale=4), 4, 314.20, 3),
    "srcnn": ("SRCNN", dict(scale_factor=2), dict(scale=2), 2, 0.8192, 0),
}


def metric_name(model_key):
    cls, _, _, s, _, _ = MODEL_CFG[model_key]
    return f"{cls} x{s} train patches/sec (16x 48x48 LR patches per GPU per step, fwd+L1+bwd+Adam)"


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
# CPU / library arms: the oracle port (the reference's own path re-stated functionally, oracle/sr_oracle.py).
# Shapes and initial weights come from the oracle itself — none of this repo's models or kernels are involved.
# ------------------------------------------------------------------------------------------------
def port_step_fn(model_key: str, batch: int, device="cpu", autocast_bf16=False):
    from oracle import sr_oracle
    cls, _, okw, scale, _, _ = MODEL_CFG[model_key]
    shape_kw = dict(okw)
    shape_kw.pop("res_scale", None)
    sd = {}
    params = []
    for k, v in sr_oracle.init_state(cls, seed=0, **shape_kw).items():
        t = v.to(device)
        if not k.startswith(("sub_mean", "add_mean")):
            t.requires_grad_(True)
            params.append(t)
        sd[k] = t
    opt = torch.optim.Adam(params, lr=1e-3)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(batch, 3, LR, LR, generator=g).to(device)
    hr = torch.rand(batch, 3, LR * scale, LR * scale, generator=g).to(device)
    fwd = sr_oracle.FORWARDS[cls]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
            sr = fwd(x, sd, **okw)
        loss = sr_oracle.l1_loss(sr.float(), hr)
        loss.backward()
        opt.step()
        return loss.item()
    return step


def time_cpu_port(model_key: str, batch: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    step = port_step_fn(model_key, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def time_cpu_infer4k(div: int = 4):
    """EDSR-large forward of the oracle port on a (540/div) x (960/div) crop of the frame, all host threads.
    Returns (frames/s extrapolated by area, seconds of CPU work, crop shape)."""
    from oracle import sr_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    sd = sr_oracle.init_state("EDSR", seed=0, n_feats=256, n_resblocks=32, scale=4)
    h, w = 540 // div, 960 // div
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        t0 = time.perf_counter()
        sr_oracle.edsr_forward(x, sd, n_resblocks=32, res_scale=0.1, scale=4)
        dt = time.perf_counter() - t0
    return 1.0 / (dt * (540 * 960) / (h * w)), dt, (h, w)


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (oracle port; the reference itself is not installable:
    DESIGN.md §1) on the box's host cores, same config / metric / unit as the GPU arm.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cls = MODEL_CFG[args.model][0]
    # 16 patches per step like the GPU arm, unless (warmup + steps) steps would exceed ~4 minutes on this host
    pps1, ms1 = time_cpu_port(args.model, 2, 1, 1)
    est_step_s = (ms1 / 1e3) * BATCH / 2
    budget_s = 240.0
    b = BATCH
    if est_step_s * (args.steps + args.warmup) > budget_s:
        b = int(max(1, min(BATCH, budget_s / (args.steps + args.warmup) / (ms1 / 1e3 / 2))))
    pps, ms = time_cpu_port(args.model, b, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": metric_name(args.model), "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cls} x{MODEL_CFG[args.model][3]} train step (BASELINE.json configs[{MODEL_CFG[args.model][5]}]), CPU port of the "
                               f"reference path (oracle/sr_oracle.py, torch {torch.__version__} CPU fp32), {b} of 16 patches per step",
                   "sample_batch": b},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {b} patches (48x48 LR), fwd+L1+bwd+Adam"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload in ("all", "infer4k") and args.model == "rcan":
        try:
            fps, secs, crop = time_cpu_infer4k()
            line["infer4k"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": f"EDSR-large forward on one {crop[0]}x{crop[1]} crop ({secs:.1f} s), frames/s scaled by area to 540x960"}
        except Exception as e:  # noqa: BLE001
            line["infer4k"] = {"value": None, "unit": "frames/s", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)


def time_library(model_key: str, steps: int, warmup: int):
    """The same torch ops the reference modules call, on the GPU through stock PyTorch / cuDNN: fp32 (TF32 off) and
    bf16 autocast — what the reference itself reaches on this B200 (SURVEY §8d "library bar").  Eager, CUDA events."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    out = {}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        step = port_step_fn(model_key, BATCH, device="cuda", autocast_bf16=ac)
        for _ in range(max(3, warmup)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
        del step
        torch.cuda.empty_cache()
    out["note"] = (f"oracle port on cuda:0 through torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, eager, "
                   f"{steps} steps after {max(3, warmup)} warm-ups, loss.item() every step")
    return out


def run_library_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    res = time_library(args.model, args.steps, args.warmup)
    out = {"impl": "library", "metric": metric_name(args.model), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": max(3, args.warmup), "higher_is_better": True, "dtype": "f32 / bf16 autocast", "data": "synthetic",
           "config": {"workload": res.pop("note")}, **res}
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
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.tmp.read().splitlines():
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
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
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def _graph_time(fn, reps_inside: int):
    """us per inner repetition of `fn` (which performs reps_inside repetitions), timed over one CUDA-graph replay."""
    fn()
    fn()
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
    return e0.elapsed_time(e1) * 1e3 / reps_inside


def time_chain_kernel(dev, model_key: str, reps=10):
    """The kernel that carries the RCAN / EDSR-baseline step: the layer-chain launch (srb_conv_chain; the per-sample
    cluster kernel of conv_cluster.cu on this shape) — one RCAN ResidualGroup (20 RCAB + tail conv = 41 tcgen05 3x3
    64->64 convs, CALayer forward / backward fused in) or the EDSR body (33 convs) per launch and direction, on
    [16,48,48,64] bf16.  Timed alone with CUDA events over graph replays; `reps` distinct input/arena sets rotate
    (10 x 289 MB > the 126 MB L2).  Returns dict(us_f, us_b, flop, n_convs, cluster)."""
    import models
    from srb200 import ops
    from srb200.trainer import FlatParams
    torch.manual_seed(0)
    if model_key == "rcan":
        mod = models.rcan.ResidualGroup(64, 3, 16, 1, 20).to(dev)
        n_convs = 41
        call = lambda x: mod(x)  # noqa: E731
    else:
        full = models.EDSR(n_feats=64, n_resblocks=16, res_scale=1.0, scale_factor=4).to(dev)
        from srb200 import functional as F200
        blocks = list(full.body)
        params = []
        for blk in blocks[:-1]:
            params += [blk.body[0].weight, blk.body[0].bias, blk.body[2].weight, blk.body[2].bias]
        params += [blocks[-1].weight, blocks[-1].bias]
        n_convs = 33
        mod = full
        call = lambda x: F200.ResTrunkFn.apply(x, full, 1.0, *params)  # noqa: E731
    # gradients land in a flat buffer that is "live" (accumulating), as inside a training step: without it every small
    # gradient tensor would get its own fill launch and those, not the chain launch, would be timed
    flat = FlatParams(mod)
    flat.begin_step(zero=True)
    arena = ops.ZeroArena(dev)
    bf = torch.bfloat16
    xs = [torch.randn(BATCH, LR, LR, 64, device=dev).to(bf).requires_grad_(True) for _ in range(reps)]
    gs = [(torch.randn(BATCH, LR, LR, 64, device=dev) * 0.01).to(bf) for _ in range(reps)]

    def wrap(fn0):
        def fn():
            ops.set_arena(arena)
            try:
                arena.reset()                     # one memset per graph replay (pooled sums / CA scratch)
                fn0()
            finally:
                ops.set_arena(None)
        return fn

    def fwd():
        with torch.no_grad():
            for x in xs:
                call(x)

    def fwd_bwd():
        # weight-gradient launches are queued and dropped: only the chain launches (+ one 5 us skip add) run
        with ops.deferred_wgrads() as q:
            for x, g in zip(xs, gs):
                call(x).backward(g)
                q.items.clear()
    # the backward launch is timed in the form the training step uses: RCAN = cluster kernel (weight gradients run beside it
    # there), EDSR (a single chain) = L2-flag kernel
    cluster_on = os.environ.get("SRB200_CHAIN_CLUSTER", "1") not in ("0",) and os.environ.get("SRB200_WGRAD_OVERLAP", "1") not in ("", "0")
    saved_bwd = os.environ.get("SRB200_CHAIN_BWD")
    if model_key == "rcan" and cluster_on and saved_bwd is None:
        os.environ["SRB200_CHAIN_BWD"] = "cluster"
    try:
        us_f = _graph_time(wrap(fwd), reps)
        us_fb = _graph_time(wrap(fwd_bwd), reps)
    finally:
        if saved_bwd is None:
            os.environ.pop("SRB200_CHAIN_BWD", None)
    flat.detach()
    flop = n_convs * 2.0 * BATCH * LR * LR * 64 * 64 * 9
    cluster = os.environ.get("SRB200_CHAIN_CLUSTER", "1") not in ("0",) and os.environ.get("SRB200_NO_CHAIN", "0") in ("", "0")
    return dict(us_f=us_f, us_b=us_fb - us_f, flop=flop, n_convs=n_convs, cluster=cluster)


def time_wide_kernel(dev, n, h, w, cin, cout, reps=6):
    """conv_wide_kernel alone: one 3x3 cin->cout conv (+bias+ReLU) on [n,h,w,cin] bf16, `reps` rotating buffer sets."""
    from srb200 import ops
    bf = torch.bfloat16
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.02).contiguous()
    b = torch.zeros(cout, device=dev)
    pk = ops.PackedWeights()
    xs = [torch.randn(n, h, w, cin, device=dev).to(bf) for _ in range(reps)]
    ys = [torch.empty(n, h, w, cout, device=dev, dtype=bf) for _ in range(reps)]

    def fn():
        for x, y in zip(xs, ys):
            ops.conv(x, 0, cin, pk, wt, b, y, 0, cout, 3, relu=True)
    us = _graph_time(fn, reps)
    return us, 2.0 * n * h * w * cin * cout * 9


def roofline_for(model_key: str, dev, peaks, ms_step):
    if model_key in ("rcan", "edsr"):
        k = time_chain_kernel(dev, model_key)
        us = 0.5 * (k["us_f"] + k["us_b"])
        tf = k["flop"] / (us * 1e-6) / 1e12
        per_step = 20 if model_key == "rcan" else 2
        fwd_cluster = k["cluster"] and os.environ.get("SRB200_CHAIN_FWD", "flags") == "cluster"
        key = "chain_cluster" if k["cluster"] else "chain_flags"
        what = ("one RCAN ResidualGroup (20 RCAB + conv = 41 tcgen05 3x3 64->64 convs, CALayer fused)" if model_key == "rcan"
                else "the EDSR body (16 ResBlocks + conv = 33 tcgen05 3x3 64->64 convs)")
        kern = ("conv_chain_kernel (L2 tile flags, 144 CTAs)" if (not k["cluster"] or model_key == "edsr") else
                "chain_cluster_kernel (conv_cluster.cu: one 6-CTA thread-block cluster per sample, 96 CTAs)" if fwd_cluster else
                "forward launch conv_chain_kernel (conv_chain.cu: L2 tile flags, 144 CTAs), backward launch chain_cluster_kernel "
                "(conv_cluster.cu: one 6-CTA cluster per sample, 96 CTAs; the other 52 SMs run the previous group's weight gradients "
                "in the step)")
        return {"bound": "tensor", "achieved": tf, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": tf / peaks["burst"],
                "traffic": TRAFFIC.get(key),
                "kernel": f"{kern}: {what} per launch on [16,48,48,64] bf16; average of the forward and the backward launch, timed alone "
                          f"over graph replays of 10 rotating arena sets (2.9 GB > L2)",
                "us_per_launch": us, "us_forward_launch": k["us_f"], "us_backward_launch": k["us_b"], "flop_per_launch": k["flop"],
                "launches_per_step": per_step,
                "share_of_step": (per_step / 2) * (k["us_f"] + k["us_b"]) / (ms_step * 1e3),
                "traffic_note": TRAFFIC_NOTE.get(key), "peak_source": peaks["source"]}
    if model_key == "rdn":
        # dense layers 128..512 -> 64 of the 16 RDBs (fwd and dgrad): conv_wide_kernel<64>; the middle one (320 -> 64) is timed
        us, flop = time_wide_kernel(dev, BATCH, LR, LR, 320, 64)
        tf = flop / (us * 1e-6) / 1e12
        return {"bound": "tensor", "achieved": tf, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": tf / peaks["burst"],
                "traffic": TRAFFIC.get("conv_wide64"),
                "kernel": "conv_wide_kernel<64>: RDN dense layer 3x3 320->64 (+bias+ReLU) on [16,48,48,320] bf16 (the dense layers "
                          "128..512->64 and their dgrads are the largest share of the RDN step), timed alone over graph replays of 6 rotating buffer sets",
                "us_per_launch": us, "flop_per_launch": flop, "traffic_note": TRAFFIC_NOTE.get("conv_wide64"),
                "peak_source": peaks["source"]}
    return None


def run_infer4k(dev, world, rank, local, peaks, min_seconds: float, with_cpu: bool):
    """EDSR x4 large 960x540 -> 3840x2160 through the reference's predict path (forward + clamp, srmodel.py:375-380).
    Collective at N>1 (every rank runs it).  Returns the dict for the line's "infer4k" key on rank 0."""
    import torch.distributed as dist
    import models
    from srb200 import lib as L
    from srb200.tiled import DistExchange, PeerExchange, TiledEDSR, partition_rows
    H, W, S = 540, 960, 4
    torch.manual_seed(0)
    m = models.EDSR(n_feats=256, n_resblocks=32, res_scale=0.1, scale_factor=S)
    m.compute_dtype = "bf16"
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    host_x = [torch.rand(1, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    dev_x = [t.to(dev) for t in host_x]
    peer = None
    if world > 1:
        r0, r1 = partition_rows(H, world)[rank]
        out_rows = (r1 - r0) * S
        if os.environ.get("SRB200_TILED_EXCHANGE", "peer") == "peer":
            # halo rows by peer stores over NVLink from one kernel per layer, the strip forward replayed as one CUDA graph
            peer = PeerExchange(device=dev)
            runner = TiledEDSR(m, peer)
            with torch.no_grad():
                runner.prepare(dev_x[0], use_graph=True)

            def fn(x):
                return runner.run(x)[rank].clamp_(0, 1)
        else:
            runner = TiledEDSR(m, DistExchange())

            def fn(x):
                return runner.forward(x)[rank].clamp_(0, 1)
    else:
        out_rows = H * S

        def fn(x):
            return m.predict_step({"lr": x}, 0)
    host_out = torch.empty(1, 3, out_rows, W * S, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for i in range(3):
            fn(dev_x[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(dev_x[0])
        e1.record()
        torch.cuda.synchronize()
        ms1 = max_over_ranks(e0.elapsed_time(e1))
        frames = int(max(10, math.ceil(min_seconds * 1e3 / ms1)))
        sampler = ClockSampler(local).start() if rank == 0 else None
        if rank == 0:
            time.sleep(0.3)
        c0 = L.launch_count()
        barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(frames):
            fn(dev_x[i % 2])
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = L.launch_count() - c0
        if peer is not None and runner.graph is not None:
            launches = runner.launches_per_frame * frames      # graph replays do not pass through the host-side counter
        # end to end: pinned host frame in, pinned host SR frame (this rank's strip) out, every frame
        xin = torch.empty_like(dev_x[0])
        frames_e = max(5, frames // 2)
        barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(frames_e):
            xin.copy_(host_x[i % 2], non_blocking=True)
            out = fn(xin)
            host_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the caller holds the frame before asking for the next
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sampler is not None else None
    if peer is not None:
        peer.close()
    if rank != 0:
        return None
    gflop_frame = 231.5644 * (H * W) / (48 * 48)
    ms = ms_total / frames
    res = {
        "metric": "EDSR x4 large 960x540 -> 3840x2160 forward frames/sec", "value": 1e3 / ms, "unit": "frames/s", "n_gpus": world,
        "frames": frames, "seconds": ms_total / 1e3, "ms_per_frame": ms, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "EDSR(n_feats=256, n_resblocks=32, res_scale=0.1, scale_factor=4) predict path (forward + clamp) on "
                               "torch.rand(1,3,540,960) (BASELINE.json configs[4])",
                   "parallelism": "single GPU, whole frame" if world == 1 else
                   (f"{world} row strips; per layer ONE kernel pushes the border rows into the neighbours' halo rows through NVLink peer "
                    f"memory and synchronises by flags (csrc/halo.cu); strip forward = one CUDA graph" if peer is not None else
                    f"{world} row strips, per-layer NCCL send/recv halo exchange (srb200/tiled.py DistExchange)"),
                   "l2": "activations of one layer (265 MB bf16) exceed the 126 MB L2; 2 input frames rotate"},
        "tflops_algorithmic": gflop_frame / ms / world, "frac_of_burst_peak_per_gpu": gflop_frame / ms / world / peaks["burst"],
        "frac_of_sustained_peak_per_gpu": gflop_frame / ms / world / peaks["sustained"],
        "e2e": {"value": 1e3 / (ms_e2e / frames_e), "unit": "frames/s", "frames": frames_e, "ms_per_frame": ms_e2e / frames_e,
                "h2d_bytes_per_step": host_x[0].numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
                "note": "per rank: H2D of the LR frame, D2H of this rank's SR rows, stream synchronised every frame"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    del m
    return res


def infer4k_roofline(dev, peaks):
    us, flop = time_wide_kernel(dev, 1, 540, 960, 256, 256, reps=3)
    tf = flop / (us * 1e-6) / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": peaks["burst"], "peak_sustained": peaks["sustained"], "unit": "TFLOP/s",
            "frac": tf / peaks["burst"], "frac_of_sustained": tf / peaks["sustained"], "traffic": TRAFFIC.get("conv_wide128"),
            "kernel": "conv_wide_kernel<128>: 3x3 256->256 (+bias+ReLU) on [1,540,960,256] bf16 = 65 of the 69 convs and 76 % of the "
                      "52.1 TFLOP of a frame; timed alone over graph replays of 3 rotating buffer sets (265 MB each > L2)",
            "us_per_launch": us, "flop_per_launch": flop, "traffic_note": TRAFFIC_NOTE.get("conv_wide128"),
            "peak_source": peaks["source"]}


def time_eager_plugin(dev, model_key: str, steps=5, warmup=3):
    """The drop-in path as Lightning would drive it: SRModel.training_step + torch.optim.Adam, eager launches."""
    import models
    cls, kw, _, scale, _, _ = MODEL_CFG[model_key]
    torch.manual_seed(0)
    m = getattr(models, cls)(**kw)
    m.compute_dtype = "bf16"
    m = m.to(dev)
    opt = m.configure_optimizers()[0]
    g = torch.Generator().manual_seed(0)
    x = torch.rand(BATCH, 3, LR, LR, generator=g).to(dev)
    hr = torch.rand(BATCH, 3, LR * scale, LR * scale, generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        out = m.training_step({"lr": x, "hr": hr}, 0)
        out["loss"].backward()
        opt.step()
        return out["loss"]
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "loss_last": float(last),
            "note": f"{cls}.training_step + loss.backward() + torch.optim.Adam.step(), eager (no CUDA graph, per-parameter "
                    f"gradients and optimizer launches), {steps} steps after {warmup} warm-ups"}


def time_srcnn_config0(dev):
    """BASELINE.json configs[0]: SRCNN x2 forward+backward, batch 16 of 48x48 LR patches, L1 loss — on the host cores
    (the reference's CPU-runnable case, oracle port) and through this repo's SRCNN on the GPU."""
    out = {"workload": "SRCNN x2 fwd+L1+bwd+Adam, batch 16 of 48x48 LR (BASELINE.json configs[0])"}
    try:
        pps, ms = time_cpu_port("srcnn", BATCH, 20, 3)
        out["cpu"] = {"value": pps, "unit": UNIT, "ms_per_step": ms, "cores": torch.get_num_threads(), "kind": "port",
                      "sample": "20 steps x 16 patches after 3 warm-ups, torch CPU fp32, all host threads"}
    except Exception as e:  # noqa: BLE001
        out["cpu"] = {"value": None, "sample": f"failed: {e}"}
    try:
        import models
        from srb200.trainer import TrainStep
        torch.manual_seed(0)
        m = models.SRCNN(scale_factor=2)
        m.compute_dtype = "fp32"
        m = m.to(dev)
        step = TrainStep(m, (BATCH, 3, LR, LR), 2, lr=1e-3, use_graph=True)
        g = torch.Generator().manual_seed(0)
        x = torch.rand(BATCH, 3, LR, LR, generator=g).to(dev)
        hr = torch.rand(BATCH, 3, LR * 2, LR * 2, generator=g).to(dev)
        step.load_batch(x, hr)
        step.capture()
        for _ in range(5):
            step.step(x, hr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step.step(x, hr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        out["gpu"] = {"value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "dtype": "f32",
                      "note": "models.SRCNN through TrainStep (CUDA-core 9x9 / 1x1 / 5x5 kernels, fp32), 50 graph replays"}
        step.close()
    except Exception as e:  # noqa: BLE001
        out["gpu"] = {"value": None, "note": f"failed: {type(e).__name__}: {e}"}
    return out


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
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def finish():
        """Leave together: every rank meets at one last barrier (rank 0 arrives after its rank-0-only measurements), then
        exits without running NCCL's teardown — a rank that tears its communicator down while a peer is still busy can
        block both (seen as a torchrun that never returns)."""
        if world > 1:
            sys.stdout.flush()
            dist.barrier()
            os._exit(0)

    cls, kw, _, scale, gflop_patch, cfg_idx = MODEL_CFG[args.model]
    line = None
    if args.workload in ("all", "train"):
        torch.manual_seed(0)                      # identical initial weights on every rank
        model = getattr(models, cls)(**kw)
        model.compute_dtype = "fp32" if args.model == "srcnn" else "bf16"
        model = model.to(dev)
        step = TrainStep(model, (BATCH, 3, LR, LR), scale, lr=1e-3, use_graph=not args.no_graph)
        g = torch.Generator(device="cpu").manual_seed(1000 + rank)
        nb = 4
        host_lr = [torch.rand(BATCH, 3, LR, LR, generator=g).pin_memory() for _ in range(nb)]
        host_hr = [torch.rand(BATCH, 3, LR * scale, LR * scale, generator=g).pin_memory() for _ in range(nb)]
        dev_lr = [t.to(dev) for t in host_lr]
        dev_hr = [t.to(dev) for t in host_hr]
        step.load_batch(dev_lr[0], dev_hr[0])
        torch.cuda.reset_peak_memory_stats(dev)
        step.capture()
        peak_mem = torch.cuda.max_memory_allocated(dev)

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
        ncu_range = os.environ.get("SRB200_NCU_RANGE", "0") not in ("", "0")      # `ncu --profile-from-start off`: timed region only
        if ncu_range:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(args.steps):
            step.step(dev_lr[i % nb], dev_hr[i % nb])
        e1.record()
        torch.cuda.synchronize()
        if ncu_range:
            torch.cuda.profiler.stop()
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
            # the next batch's H2D copy is started behind this batch's take-over and runs under this step (Runner.fit does the
            # same with its look-ahead); every batch still crosses PCIe inside the timed region
            nxt = (host_lr[(i + 1) % nb], host_hr[(i + 1) % nb]) if i + 1 < args.steps else None
            t = step.step_async(host_lr[i % nb], host_hr[i % nb], prefetch=nxt)      # loss -> pinned host slot behind the step
            if i:
                last = step.loss_of(t - 1)        # D2H read of every step's result, one step behind the launch
        last = step.loss_of(t)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None

        # ---- timed region 2b: the same step fed by the GPU data path (srb200/data.py): synthetic uint8 images
        # resident in HBM, batches (random crop / rot90 / flips / to_tensor as srdata.py:57-169) built by one kernel straight
        # into the step's static buffers; per step the host uploads 56 bytes per patch and reads the loss back
        gpu_data = None
        if True:
            from srb200.data import PatchSampler
            import numpy as _np
            ps = PatchSampler(scale, LR, device=dev, augment=True, seed=rank)
            rs = _np.random.RandomState(rank)
            for _ in range(8):
                hr_img = rs.randint(0, 256, (LR * scale * 4, LR * scale * 4, 3), dtype=_np.uint8)
                lr_img = rs.randint(0, 256, (LR * 4, LR * 4, 3), dtype=_np.uint8)      # synthetic: no real down-scaling needed here
                ps.add(hr_img, lr_img)
            for _ in range(3):
                ps.fill(step.x, step.hr)
                step.run()
            barrier()
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                ps.fill(step.x, step.hr)
                t = step.run_async()
                if i:
                    last_g = step.loss_of(t - 1)
            last_g = step.loss_of(t)
            e1.record()
            torch.cuda.synchronize()
            barrier()
            ms_g = max_over_ranks(e0.elapsed_time(e1))
            gpu_data = {"value": BATCH * world * args.steps / (ms_g / 1e3), "unit": UNIT, "ms_per_step": ms_g / args.steps,
                        "h2d_bytes_per_step": 56 * BATCH, "d2h_bytes_per_step": 4, "loss_last": last_g,
                        "note": "batches built on the GPU by srb_patch_batch from 8 resident synthetic uint8 images per rank "
                                "(random crop, rot90, flips, to_tensor: the reference's srdata.py:57-169 semantics)"}

        # ---- timed region 3: sustained (>= 3 s of the same step, clocks and power settled) ----------
        ms_step = ms_total / args.steps
        sus_steps = int(max(args.steps, math.ceil(args.sustain_seconds * 1e3 / ms_step)))
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(sus_steps):
            step.step(dev_lr[i % nb], dev_hr[i % nb])
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_sus = max_over_ranks(e0.elapsed_time(e1))
        clocks_sus = sampler2.stop() if rank == 0 else None

        value = BATCH * world * args.steps / (ms_total / 1e3)
        e2e_value = BATCH * world * args.steps / (ms_e2e / 1e3)
        step_tflops = gflop_patch * BATCH / (ms_step / 1e3) / 1e3      # per GPU
        chain_note = ""
        if args.model in ("rcan", "edsr"):
            off = os.environ.get("SRB200_NO_CHAIN", "0") not in ("", "0")
            ov = getattr(step, "overlap", None)
            chain_note = ("; 64-channel trunks run as layer-chain launches (srb_conv_chain: forward = L2-flag kernel on all SMs, "
                          "backward = per-sample cluster kernel on 96 SMs"
                          + (f" with the previous group's weight gradients on <= {ov.sm_budget} CTAs of a side stream "
                             f"({ov.max_sections} groups)" if ov is not None else "") + ")"
                          + (" [disabled]" if off else ""))
        line = {
            "metric": metric_name(args.model), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.model == "srcnn" else "bf16", "data": "synthetic",
            "config": {
                "workload": f"{cls} x{scale} training step (BASELINE.json configs[{cfg_idx}]): {kw}, batch {BATCH} x 3x{LR}x{LR} LR per GPU, "
                            f"L1 loss, Adam lr=1e-3, bf16 activations / fp32 accumulate+master weights{chain_note}",
                "parallelism": f"dp{world}", "cuda_graph": step.graph is not None,
                "l2": f"no explicit flush: one step streams {peak_mem / 2**30:.2f} GiB of saved activations and gradients "
                      f"(>> 126 MB L2); 4 distinct input batches rotate",
                "loss_last": loss_dev,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "loss_last": last,
                    "note": "every step: pinned-host lr+hr cross PCIe (the next batch's copy runs under the current step) and the "
                            "loss is copied to pinned host memory and read by the host, one step behind the launch"},
            "e2e_gpu_data": gpu_data,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "sustained": {"value": BATCH * world * sus_steps / (ms_sus / 1e3), "unit": UNIT, "steps": sus_steps,
                          "seconds": ms_sus / 1e3, "ms_per_step": ms_sus / sus_steps, "clocks": clocks_sus,
                          "frac_of_sustained_peak": gflop_patch * BATCH / (ms_sus / sus_steps / 1e3) / 1e3 / peaks["sustained"]},
            "roofline_step": {"bound": "tensor", "achieved": step_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                              "frac": step_tflops / peaks["sustained"], "frac_of_burst": step_tflops / peaks["burst"],
                              "note": f"whole step: {gflop_patch} algorithmic GFLOP/patch x {BATCH} / ms_per_step, vs sustained bf16 peak"},
        }
        step.close()
        del step, model
        torch.cuda.empty_cache()

    infer = None
    if args.workload in ("all", "infer4k") and args.model == "rcan":
        try:
            infer = run_infer4k(dev, world, rank, local, peaks, args.sustain_seconds, with_cpu=(world == 1))
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise
            infer = {"value": None, "unit": "frames/s", "error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if rank != 0:
        finish()
        return

    if line is None:      # --workload infer4k: the 4K numbers are the line
        line = dict(infer or {})
        line.setdefault("higher_is_better", True)
        line.setdefault("scaling", "strong")
        line["vs_baseline"] = None
        infer_target = line
    else:
        if infer is not None:
            line["infer4k"] = infer
        infer_target = infer
        try:
            rl = roofline_for(args.model, dev, peaks, line["ms_per_step"])
            if rl is not None:
                line["roofline"] = rl
        except Exception as e:  # noqa: BLE001
            line["roofline"] = {"error": f"{type(e).__name__}: {e}"}
    if infer_target is not None and infer_target.get("value"):
        try:
            infer_target["roofline"] = infer4k_roofline(dev, peaks)
        except Exception as e:  # noqa: BLE001
            infer_target["roofline"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if world == 1 and not args.no_cpu_baseline:
        if "ms_per_step" in line:
            try:
                b = BATCH
                pps, ms = time_cpu_port(args.model, b, 5, 1)
                line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"5 steps x {b} patches (48x48 LR) after 1 warm-up (~{6 * ms / 1e3:.0f} s of CPU work), "
                                                  "fwd+L1+bwd+Adam, torch CPU fp32, all host threads"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        if infer_target is not None and infer_target.get("value"):
            try:
                fps, secs, crop = time_cpu_infer4k()
                infer_target["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                                "sample": f"oracle EDSR-large forward on one {crop[0]}x{crop[1]} crop ({secs:.1f} s of CPU work), "
                                                          "frames/s scaled by area to 540x960, torch CPU fp32, all host threads"}
            except Exception as e:  # noqa: BLE001
                infer_target["cpu_baseline"] = {"value": None, "unit": "frames/s", "kind": "port", "sample": f"failed: {e}"}
    if world == 1 and args.workload == "all" and not args.no_extras and "ms_per_step" in line:
        try:
            line["library"] = time_library(args.model, 5, 3)
        except Exception as e:  # noqa: BLE001
            line["library"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            line["eager_plugin"] = time_eager_plugin(dev, args.model)
        except Exception as e:  # noqa: BLE001
            line["eager_plugin"] = {"error": f"{type(e).__name__}: {e}"}
        if args.model == "rcan" and not args.no_cpu_baseline:
            line["configs0_srcnn"] = time_srcnn_config0(dev)
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "library"])
    ap.add_argument("--model", default="rcan", choices=list(MODEL_CFG))
    ap.add_argument("--workload", default="all", choices=["all", "train", "infer4k"])
    ap.add_argument("--sustain-seconds", type=float, default=3.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the library / eager-plugin / SRCNN side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "library":
        run_library_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
