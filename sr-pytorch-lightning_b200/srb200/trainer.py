"""Training-step runner for the B200 path: flat parameter / gradient / Adam-state buffers, one
CUDA graph per step (forward + L1 + backward + gradient all-reduce + Adam), data-parallel over
`torch.distributed` (one process per GPU, NCCL over NVLink).

What it replaces in the reference: Lightning's fit loop around `SRModel.training_step`
(/root/reference/models/srmodel.py:160-171), `configure_optimizers` (srmodel.py:145-154, Adam
defaults) and the implicit DDP gradient all-reduce.  Python runs once, at capture;
a replayed step is a single `cudaGraphLaunch` — the >5000 eager launches per RCAN step the
reference issues never touch the host again.

Per step, outside the conv/CA kernels themselves, the graph holds: one table-driven re-pack of all
weights (srb_pack_table), one memset of the flat gradient buffer, one memset of the zero arena
(pooled sums / CA scratch), a handful of batched weight-gradient launches, one Adam launch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import functional as F200
from . import ops


class FlatParams:
    """Re-homes every trainable parameter of `model` into one contiguous fp32 buffer (16-byte
    aligned slices) with matching flat gradient and Adam-moment buffers.  `state_dict()` is
    unaffected (parameters become views).  Installs `_srb_grad` views so the backward kernels
    write gradients in place (srb200.functional._grad_target)."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p._srb_grad = self.grad[o:o + p.numel()].view_as(p)
                p._srb_grad_live = False
        self.offsets = offs

    def begin_step(self, zero: bool = True):
        """zero=True: clear the whole gradient buffer with one memset and let every kernel
        accumulate (no per-layer fills).  zero=False: the first write of each gradient overwrites."""
        if zero:
            self.grad.zero_()
        for p in self.params:
            p._srb_grad_live = zero

    def collect_autograd_grads(self):
        """Parameters whose gradient torch autograd produced itself (weight-norm's g and v: the kernels only see the derived
        filter) have it in `.grad`: move it into the flat buffer the fused Adam reads."""
        for p in self.params:
            if p.grad is not None:
                p._srb_grad.add_(p.grad)
                p.grad = None

    def grads_by_name(self, model):
        return {k: p._srb_grad for k, p in model.named_parameters() if p.requires_grad}

    def detach(self):
        for p in self.params:
            if hasattr(p, "_srb_grad"):
                del p._srb_grad
                del p._srb_grad_live


def check_supported(model):
    """The fused step implements exactly what the reference's default configuration trains with: ONE unit-weight L1
    loss (`losses: l1`, srmodel.py:37,549) and Adam (srmodel.py:57,145-154).  Anything else must fail loudly instead
    of silently training as L1/Adam; such configurations can still use `SRModel.training_step` with a torch optimizer."""
    losses = getattr(model, "_losses", None)
    if losses is not None:
        if len(losses) != 1 or losses[0].name not in ("l1", "mae") or float(losses[0].weight) != 1.0:
            desc = "+".join(f"{l.weight}*{l.name}" for l in losses)
            raise NotImplementedError(f"TrainStep fuses a single unit-weight L1 loss; the model was built with losses='{desc}'")
    opt = getattr(model, "_optim", None)
    if opt is not None and opt is not torch.optim.Adam:
        raise NotImplementedError(f"TrainStep fuses Adam; the model was built with optimizer {opt.__name__}")


class TrainStep:
    """One training step of an SRModel subclass on fixed-shape batches.

    step(lr_batch, hr_batch) -> loss tensor (device).  Inputs may be host (pinned) or device
    tensors; they are copied into static device buffers, then the captured graph is replayed."""

    def __init__(self, model, lr_shape, scale: int, *, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, use_graph: bool = True, process_group=None):
        check_supported(model)
        self.model = model
        self.flat = FlatParams(model)
        dev = self.flat.flat.device
        self.device = dev
        n, c, h, w = lr_shape
        self.x = torch.zeros(lr_shape, dtype=torch.float32, device=dev)
        self.hr = torch.zeros((n, c, h * scale, w * scale), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.hp = dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = use_graph
        self.graph = None
        self.launches_per_step = 0
        self.arena = ops.ZeroArena(dev)
        self.pack_table = None
        # weight gradients of a ResidualGroup under the next group's backward chain (ops.WgradOverlap): needs the cluster
        # chain kernel, which leaves 52 of the 148 SMs free.  SRB200_WGRAD_OVERLAP=0 turns it off;
        # SRB200_WGRAD_OVERLAP_SMS / _GROUPS set the CTA budget of the overlapped launches / how many groups overlap.
        import os
        self.overlap = None
        if dev.type == "cuda" and os.environ.get("SRB200_WGRAD_OVERLAP", "1") not in ("", "0") \
                and os.environ.get("SRB200_CHAIN_CLUSTER", "1") not in ("0",) \
                and os.environ.get("SRB200_NO_CHAIN", "0") in ("", "0"):
            self.overlap = ops.WgradOverlap(dev, sm_budget=int(os.environ.get("SRB200_WGRAD_OVERLAP_SMS", "52")),
                                            max_sections=int(os.environ.get("SRB200_WGRAD_OVERLAP_GROUPS", "9")),
                                            on_section_done=self._reduce_bucket if self.world > 1 and
                                            os.environ.get("SRB200_ALLREDUCE_BUCKETS", "0") not in ("", "0") else None)
        self._stage = None          # staging buffers / copy stream of prefetch()
        self._prefetched = None
        self._reduced = []          # [lo, hi) element ranges of flat.grad already all-reduced in this step
        self._pending = None        # (lo, hi, groups) collected for the next bucket
        # SRB200_ALLREDUCE_BUCKETS: "0" (default) = one all-reduce of the whole buffer after backward; "tail" = the gradients of the groups whose weight gradients ran on the side
        # stream are all-reduced as ONE bucket on that stream once it has finished, i.e. under the weight-gradient launches
        # that are left for the main stream (which then leave SRB200_ALLREDUCE_SMS SMs to NCCL); "1" = a bucket every
        # SRB200_ALLREDUCE_BUCKET_GROUPS groups during backward (measured slower: the side stream is already as long as the
        # chain stream).  Measured at N=2: 7.66 ms ("0"), 7.71 ("tail"), 7.75 ("1", 2 groups), 7.88 ("1", 1 group).
        self._bucket_mode = os.environ.get("SRB200_ALLREDUCE_BUCKETS", "0")
        self._bucket_groups = (1 << 30) if self._bucket_mode == "tail" else int(os.environ.get("SRB200_ALLREDUCE_BUCKET_GROUPS", "2"))
        self._nccl_sms = int(os.environ.get("SRB200_ALLREDUCE_SMS", "16"))
        self.main = torch.cuda.Stream(device=dev, priority=-1) if dev.type == "cuda" else None     # the step's own (higher-priority) stream
        self.sync_from_rank0()

    def sync_from_rank0(self):
        """Replicas start from rank 0's parameters and Adam state, as Lightning DDP broadcasts them for the reference
        (call again after loading a checkpoint on one rank only)."""
        if self.world > 1:
            src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
            for t in (self.flat.flat, self.flat.m, self.flat.v, self.flat.step_dev):
                dist.broadcast(t, src=src, group=self.pg)

    def _reduce_bucket(self, lo_ptr: int, hi_ptr: int):
        """All-reduce the slice of the flat gradient buffer a ResidualGroup's backward has completed (called on the
        weight-gradient side stream right behind that group's launches): the bucket travels over NVLink under the following
        groups' backward chains, as Lightning DDP's buckets do for the reference (configs/all.yml:125-127)."""
        base = self.flat.grad.data_ptr()
        lo = max(0, (lo_ptr - base) // 4)
        hi = min(self.flat.numel, (hi_ptr - base + 3) // 4)
        if hi <= lo:
            return
        # Several groups per bucket: the all-reduce sits on the weight-gradient stream, which is as long as the chain stream
        # beside it; one launch per group (8 x ~50 us) made the step LONGER at N=2 (7.71 -> 7.88 ms), fewer and larger ones
        # amortise the launch latency.  Consecutive groups are adjacent parameter ranges, so a bucket stays one slice.
        if self._pending and (self._pending[1] == lo or self._pending[0] == hi):
            self._pending = (min(lo, self._pending[0]), max(hi, self._pending[1]), self._pending[2] + 1)
        else:
            self._flush_bucket()
            self._pending = (lo, hi, 1)
        if self._pending[2] >= self._bucket_groups:
            self._flush_bucket()

    def _flush_bucket(self):
        if self._pending:
            lo, hi, _ = self._pending
            dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            self._reduced.append((lo, hi))
        self._pending = None

    def _reduce_rest(self):
        """All-reduce what no bucket covered (head, up-sampling, tail, the groups whose weight gradients ran last)."""
        if self._pending:            # (a partial bucket: reduce it with the rest, on this stream)
            self._pending = None
        cur = 0
        for lo, hi in sorted(self._reduced) + [(self.flat.numel, self.flat.numel)]:
            if lo > cur:
                dist.all_reduce(self.flat.grad[cur:lo], op=dist.ReduceOp.SUM, group=self.pg)
            cur = max(cur, hi)
        self._reduced = []

    # the work of one step; captured once
    def _body(self):
        ops.set_arena(self.arena)
        try:
            self.arena.reset()
            self.flat.begin_step(zero=True)
            if self.pack_table is not None:
                self.pack_table.run()                 # all packed weight copies, one launch
            else:
                ops.invalidate_packed()               # (warm-up) re-pack lazily, conv by conv
            sr = self.model.forward(self.x)
            loss = F200.l1_loss(sr, self.hr)
            # weight gradients batched, off the dgrad chain; ONE flush for the whole step, so that the launch planner
            # (wgrad_umma.cu plan_launches) sees every layer and leaves at most one partially filled launch
            ops.set_wgrad_overlap(self.overlap)
            if self.overlap is not None:
                self.overlap.begin_pass()
            tail_bucket = self.world > 1 and self.overlap is not None and self.overlap.on_section_done is not None \
                and self._bucket_mode == "tail"
            try:
                from . import lib as L
                import ctypes as C
                ctx = C.c_void_p(L.ctx(self.device.index)) if tail_bucket else None
                try:
                    with ops.deferred_wgrads(max_items=4096):
                        loss.backward()
                        if tail_bucket and self._pending:
                            with torch.cuda.stream(self.overlap.side):
                                self._flush_bucket()          # behind the side stream's last weight-gradient launch
                            # the launches flushed on leaving this block share the GPU with that all-reduce
                            L.load().srb_set_wgrad_sm_budget(ctx, max(1, L.load().srb_num_sms(ctx) - self._nccl_sms))
                finally:
                    if tail_bucket:
                        L.load().srb_set_wgrad_sm_budget(ctx, 0)
                if self.overlap is not None:
                    self.overlap.join()
                self.flat.collect_autograd_grads()
            finally:
                ops.set_wgrad_overlap(None)
            if self.world > 1:
                self._reduce_rest()
            ops.inc_counter(self.flat.step_dev)
            ops.adam_step(self.flat.flat, self.flat.grad, self.flat.m, self.flat.v, step=0,
                          step_dev=self.flat.step_dev, grad_scale=1.0 / self.world, **self.hp)
            self.loss.copy_(loss.detach())
        finally:
            ops.set_arena(None)

    def prepare(self, warmup: int = 2):
        """Warm-up steps (sizes the arena, fills the packed-weight caches), then build the pack
        table and, if enabled, capture the step into a CUDA graph."""
        from . import lib as L
        side = self.main
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                self._body()
            self.pack_table = ops.PackTable(self.device)
            self.pack_table.install()
            self._body()                              # one more eager step through the table path
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if not self.use_graph:
            return
        c0 = L.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.main):
            self._body()
        self.launches_per_step = L.launch_count() - c0
        torch.cuda.synchronize()

    capture = prepare

    def load_batch(self, lr_batch, hr_batch):
        pf = self._prefetched
        if pf is not None and pf[0] is lr_batch and pf[1] is hr_batch:
            # staged by prefetch() while the previous step ran: two device-to-device copies behind the copy stream's event
            self._prefetched = None
            torch.cuda.current_stream(self.device).wait_event(pf[2])
            self.x.copy_(self._stage[0], non_blocking=True)
            self.hr.copy_(self._stage[1], non_blocking=True)
            return
        self.x.copy_(lr_batch, non_blocking=True)
        self.hr.copy_(hr_batch, non_blocking=True)

    def prefetch(self, lr_batch, hr_batch):
        """Start the host-to-device copy of the NEXT batch (pinned host tensors) on a copy stream, into staging buffers; the
        step() that is later called with the same two tensors takes it from there.  Call it before step() of the current
        batch: the DMA then runs under that step instead of in front of the next one (the reference gets the same from its
        DataLoader's pin_memory + non_blocking transfers, srdata.py:514-516)."""
        if self._stage is None:
            self._stage = (torch.empty_like(self.x), torch.empty_like(self.hr))
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_free = None
        if self._stage_free is not None:
            self._copy_stream.wait_event(self._stage_free)      # the previous staged batch has been taken over
        with torch.cuda.stream(self._copy_stream):
            self._stage[0].copy_(lr_batch, non_blocking=True)
            self._stage[1].copy_(hr_batch, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._prefetched = (lr_batch, hr_batch, ev)

    def run(self):
        # the step's Adam kernel rewrites the parameters without bumping tensor versions: packed copies that the
        # pack table does not refresh (created after capture, e.g. by a validation pass on another shape) are stale now
        ops.invalidate_packed()
        if self.graph is not None:
            self.graph.replay()
        else:
            from . import lib as L
            c0 = L.launch_count()
            self._body()
            self.launches_per_step = L.launch_count() - c0
        return self.loss

    def step(self, lr_batch, hr_batch, prefetch=None):
        """One step on (lr_batch, hr_batch).  prefetch = (lr, hr) of the NEXT batch (pinned host tensors): its host-to-device
        copy is started on the copy stream right after this batch has been taken over and runs under this step's graph."""
        staged = self._prefetched is not None
        self.load_batch(lr_batch, hr_batch)
        if staged and self._prefetched is None:      # the staging buffers may be refilled once these copies have run
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.device))
        if prefetch is not None:
            self.prefetch(prefetch[0], prefetch[1])
        return self.run()

    def step_async(self, lr_batch, hr_batch, prefetch=None) -> int:
        """step() whose loss travels to a pinned host slot behind the step instead of being read synchronously: returns a
        ticket for loss_of().  A loop that reads ticket i-1 after launching step i keeps one step queued on the device, so
        the host's per-step work (launch, the read's wake-up) is hidden; at most len(ring)-1 tickets may be outstanding."""
        self.step(lr_batch, hr_batch, prefetch=prefetch)
        return self._post_loss()

    def run_async(self) -> int:
        """run() on the batch already in the static buffers, loss through the pinned ring (see step_async)."""
        self.run()
        return self._post_loss()

    def _post_loss(self) -> int:
        if getattr(self, "_loss_ring", None) is None:
            self._loss_ring = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._loss_events = [torch.cuda.Event() for _ in range(4)]
            self._ticket = 0
        t = self._ticket
        self._ticket += 1
        self._loss_ring[t % 4: t % 4 + 1].copy_(self.loss.reshape(1), non_blocking=True)
        self._loss_events[t % 4].record(torch.cuda.current_stream(self.device))
        return t

    def loss_of(self, ticket: int) -> float:
        if not (self._ticket - 4 < ticket < self._ticket):
            raise ValueError(f"ticket {ticket} is no longer (or not yet) in the loss ring")
        self._loss_events[ticket % 4].synchronize()
        return float(self._loss_ring[ticket % 4])

    def close(self):
        ops.PackTable.uninstall()
        self.flat.detach()
