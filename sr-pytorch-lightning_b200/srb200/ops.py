"""Host-side operator layer: thin, non-differentiable launch wrappers over the C ABI
(include/srb200.h).  Tensors are torch CUDA tensors used purely as device memory; every launch
goes to `torch.cuda.current_stream()` so Lightning/AMP/DDP stream ordering and CUDA-graph capture
work (SURVEY §8b "Threading / streams").

Activations: NHWC contiguous `[N, H, W, Cs]` tensors (bf16 or fp32).  A *channel slice* is
addressed as (tensor, channel_offset, channels) — the replacement for torch.cat in RDN
(/root/reference/models/rdn.py:21,108).
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _ctx(t):
    if not t.is_cuda:
        raise RuntimeError("srb200 kernels need CUDA tensors (there is no CPU fallback)")
    return C.c_void_p(L.ctx(t.device.index if t.device.index is not None else torch.cuda.current_device()))


def _nhwc(t):
    assert t.dim() == 4 and t.is_contiguous(), "activation must be a contiguous NHWC tensor"
    return t.shape


def dtype_code(t):
    return _DT[t.dtype]


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def pack_weight(w: torch.Tensor, packing: int, mode: int, shuffle: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 OIHW parameter -> packed device buffer (uint8 tensor; `out` = caller-provided slice)."""
    lib = L.load()
    assert w.dtype == torch.float32 and w.is_contiguous()
    cout, cin, k, _ = w.shape
    nbytes = lib.srb_packed_weight_bytes(cout, cin, k, packing, mode)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    else:
        assert out.dtype == torch.uint8 and out.numel() == nbytes and out.is_contiguous() and out.device == w.device
    L.check(lib.srb_pack_weight(_ctx(w), _p(w), cout, cin, k, packing, mode, shuffle, _p(out), _stream()), "srb_pack_weight")
    return out


def pack_bias(b: torch.Tensor, shuffle: int) -> torch.Tensor:
    if shuffle <= 1:
        return b
    lib = L.load()
    out = torch.empty_like(b)
    L.check(lib.srb_pack_bias(_ctx(b), _p(b), b.numel(), shuffle, _p(out), _stream()), "srb_pack_bias")
    return out


_generation = 0


def invalidate_packed():
    """Mark every packed-weight cache stale (parameters were updated by a kernel that does not
    bump tensor versions, e.g. the flat Adam step)."""
    global _generation
    _generation += 1


_all_packs = weakref.WeakSet()
# Addresses of the packed buffers the installed PackTable refreshes once per step (None: no table).  ONLY those
# entries are trusted without a version check; everything else (keys first used after the table was built: a
# validation image that takes the per-layer path, fp32 / SIMT packings, a second model) falls back to the
# (data_ptr, _version, generation) tag, and `invalidate_packed()` — called by TrainStep.run after every step,
# because the flat Adam kernel does not bump tensor versions — makes them re-pack on their next use.
_managed = None


class PackedWeights:
    """Cache of packed copies of one conv weight, rebuilt when the parameter changes
    (keyed on data_ptr and the tensor version counter, SURVEY §8b "Ownership")."""

    def __init__(self):
        self._cache = {}
        _all_packs.add(self)

    def get(self, w: torch.Tensor, packing: int, mode: int, shuffle: int = 0, out: torch.Tensor | None = None):
        """`out`: pack into this caller-owned slice (a layer of a chain's filter bank) instead of a
        private buffer; such entries live under their own cache key."""
        key = (packing, mode, shuffle) if out is None else (packing, mode, shuffle, "bank")
        hit = self._cache.get(key)
        same_home = hit is not None and (out is None or hit[1].data_ptr() == out.data_ptr())
        if same_home and _managed is not None and hit[1].data_ptr() in _managed and hit[2].data_ptr() == w.data_ptr():
            return hit[1]
        tag = (w.data_ptr(), w._version, w.device, _generation)
        if same_home and hit[0] == tag:
            return hit[1]
        packed = pack_weight(w.detach(), packing, mode, shuffle, out=out)
        self._cache[key] = (tag, packed, w.detach())
        return packed

    def get_bias(self, b: torch.Tensor, shuffle: int):
        if b is None or shuffle <= 1:
            return b.detach() if b is not None else None
        key = ("bias", shuffle)
        hit = self._cache.get(key)
        if hit is not None and _managed is not None and hit[1].data_ptr() in _managed and hit[2].data_ptr() == b.data_ptr():
            return hit[1]
        tag = (b.data_ptr(), b._version, b.device, _generation)
        if hit is not None and hit[0] == tag:
            return hit[1]
        packed = pack_bias(b.detach(), shuffle)
        self._cache[key] = (tag, packed, b.detach())
        return packed


class PackTable:
    """Device table of every (parameter -> packed buffer) pair currently cached by the model's
    PackedWeights objects; `run()` re-packs them all in one launch (srb_pack_table).  Built after
    a warm-up step, when every conv has been through forward and backward once.  While a table is
    installed, the buffers it lists (`_managed`) are trusted: addresses are stable and the table refreshes
    their contents each step.  Entries created later are NOT in the table and keep the version/generation check."""

    def __init__(self, device):
        import numpy as np
        rows = []
        self.keep = []
        max_elems = 1
        for pw in list(_all_packs):
            for key, (tag, packed, src) in pw._cache.items():
                if src.device != device:
                    continue
                if key[0] == "bias":
                    rows.append((src.data_ptr(), packed.data_ptr(), src.numel(), 1, 0, 0, 0, key[1]))
                else:
                    packing, mode, shuffle = key[:3]
                    cout, cin, k, _ = src.shape
                    rows.append((src.data_ptr(), packed.data_ptr(), cout, cin, k, packing, mode, shuffle))
                    max_elems = max(max_elems, src.numel())
                self.keep.append((packed, src))
        dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("ksize", "<i4"),
                       ("packing", "<i4"), ("mode", "<i4"), ("shuffle", "<i4")])
        assert dt.itemsize == C.sizeof(L.PackItem)
        arr = np.array(rows, dtype=dt)
        self.n = len(rows)
        # grid rows are sized for the median item (srb_pack_table: one CTA per 2048 of max_elems; larger items loop).  A 3x3
        # tensor-core item wants one CTA per 32 x 64 x 9 staged elements (misc.cu pack_umma3_staged), the others one per 2048.
        def ctas(r):
            cout, cin, k, packing, mode = r[2], r[3], r[4], r[5], r[6]
            if packing == L.PACK_UMMA and k == 3:
                kdim, rows_ = (cin, cout) if mode == L.PACK_FWD else (cout, cin)
                return -(-kdim // 64) * (-(-rows_ // 32) if mode == L.PACK_FWD else 2 * -(-rows_ // 64))
            return -(-(cout * cin * k * k) // 2048)
        demand = sorted(ctas(r) for r in rows if r[4] > 0)
        self.max_elems = 2048 * demand[len(demand) // 2] if demand else max_elems
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        self.device = device

    def install(self):
        global _managed
        _managed = frozenset(packed.data_ptr() for packed, _ in self.keep)

    @staticmethod
    def uninstall():
        global _managed
        _managed = None

    def run(self):
        L.check(L.load().srb_pack_table(C.c_void_p(L.ctx(self.device.index)), _p(self.table), self.n, self.max_elems,
                                        _stream()), "srb_pack_table")


class ZeroArena:
    """One fp32 buffer cleared once per step; kernels that accumulate with atomics (pooled sums,
    CA scratch) take zero-initialised slices from it instead of launching a fill each."""

    def __init__(self, device):
        self.device = device
        self.buf = None
        self.off = 0
        self.need = 0

    def reset(self):
        if self.buf is None or self.buf.numel() < self.need:
            self.buf = torch.zeros(max(self.need, 1024), dtype=torch.float32, device=self.device)
        else:
            self.buf.zero_()
        self.off = 0
        self.need = 0

    def take(self, shape):
        n = 1
        for s in shape:
            n *= int(s)
        n4 = (n + 3) // 4 * 4
        self.need += n4
        if self.buf is None or self.off + n4 > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float32, device=self.device)   # sizing pass
        out = self.buf[self.off:self.off + n].view(shape)
        self.off += n4
        return out


_arena = None


def set_arena(arena):
    global _arena
    _arena = arena


def zeros_f32(shape, device):
    """Zero-initialised fp32 scratch: a slice of the per-step arena when one is active."""
    if _arena is not None and _arena.device == device:
        return _arena.take(shape)
    return torch.zeros(shape, dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------
def make_conv_desc(x, x_co, cin, y, y_co, cout, k, *, flags=0, scale=1.0, shuffle=0, res=None, mask=None, y2=None,
                   colsum_groups=0, backend=L.BACKEND_AUTO) -> L.ConvDesc:
    n, h, w, xcs = _nhwc(x)
    d = L.ConvDesc()
    d.N, d.H, d.W = n, h, w
    d.Cin, d.Cout, d.ksize = cin, cout, k
    d.dtype = dtype_code(x)
    d.flags = flags
    d.scale = float(scale)
    d.shuffle = shuffle
    d.colsum_groups = colsum_groups
    d.backend = backend
    d.x_cs, d.x_co = xcs, x_co
    d.y_cs, d.y_co = y.shape[3], y_co
    if res is not None:
        d.r_cs, d.r_co = res[0].shape[3], res[1]
    if mask is not None:
        d.m_cs, d.m_co = mask[0].shape[3], mask[1]
    if y2 is not None:
        d.y2_cs, d.y2_co = y2[0].shape[3], y2[1]
    return d


def conv_uses_umma(desc: L.ConvDesc) -> bool:
    return bool(L.load().srb_conv_uses_umma(C.byref(desc)))


def conv(x, x_co, cin, weights: PackedWeights, w_param, bias, y, y_co, cout, k, *, mode=L.PACK_FWD, relu=False,
         scale=1.0, shuffle=0, res=None, mask=None, y2=None, colsum=None, colsum_groups=0, backend=L.BACKEND_AUTO):
    """y[..., y_co:y_co+cout'] = epilogue(conv(x[..., x_co:x_co+cin])).

    mode PACK_DGRAD: `w_param` is the forward parameter [cin_fwd=cout, ...]; the kernel computes
    the input gradient (here `cin` = forward Cout, `cout` = forward Cin).
    res / mask / y2: (tensor, channel_offset) on the output pixel grid."""
    lib = L.load()
    flags = (L.RELU if relu else 0) | (L.RESIDUAL if res is not None else 0) | (L.MASK if mask is not None else 0) | \
            (L.OUT2 if y2 is not None else 0) | (L.COLSUM if colsum is not None else 0)
    d = make_conv_desc(x, x_co, cin, y, y_co, cout, k, flags=flags, scale=scale, shuffle=shuffle, res=res, mask=mask,
                       y2=y2, colsum_groups=colsum_groups, backend=backend)
    packing = L.PACK_UMMA if conv_uses_umma(d) else L.PACK_SIMT
    if backend == L.BACKEND_UMMA:
        packing = L.PACK_UMMA
    # for DGRAD the pixel-shuffle permutation applies to the (forward) output channels = our inputs
    wp = weights.get(w_param, packing, mode, shuffle if mode == L.PACK_FWD else 0)
    L.check(lib.srb_conv(_ctx(x), C.byref(d), _p(x), _p(wp), _p(bias), _p(res[0]) if res else None,
                         _p(mask[0]) if mask else None, _p(y), _p(y2[0]) if y2 else None, _p(colsum), _stream()),
            "srb_conv")
    return y


def conv_dgrad_shuffled(gu, weights: PackedWeights, w_param, y, shuffle, **kw):
    """Input gradient of a conv whose output went through PixelShuffle(r): `gu` is the
    un-shuffled gradient in (ij, c') channel order, so the DGRAD packing must permute its input
    channels the same way."""
    lib = L.load()
    cout_f, cin_f, k, _ = w_param.shape
    res = kw.get("res")
    flags = (L.RESIDUAL if res is not None else 0)
    d = make_conv_desc(gu, 0, cout_f, y, 0, cin_f, k, flags=flags, res=res)
    packing = L.PACK_UMMA if conv_uses_umma(d) else L.PACK_SIMT
    wp = weights.get(w_param, packing, L.PACK_DGRAD, shuffle)
    L.check(lib.srb_conv(_ctx(gu), C.byref(d), _p(gu), _p(wp), None, _p(res[0]) if res else None, None, _p(y), None,
                         None, _stream()), "srb_conv(dgrad)")
    return y


class _WgradQueue:
    """Deferred weight gradients.  Inside `deferred_wgrads()` every conv_wgrad call is recorded
    (tensors kept alive) instead of launched; the flush hands the whole list to
    srb_conv_wgrad_batched, where the tcgen05-eligible layers share a few persistent launches."""

    def __init__(self):
        self.active = False
        self.items = []
        self.max_items = 296

    def push(self, desc, x, gy, dw, dbias):
        self.items.append((desc, x, gy, dw, dbias))
        if len(self.items) >= self.max_items:
            self.flush()

    def flush(self):
        if not self.items:
            return
        items, self.items = self.items, []
        self.flushed = items          # (WgradOverlap keeps these alive until the side stream has been joined)
        arr = (L.WgradItem * len(items))()
        for i, (d, x, gy, dw, dbias) in enumerate(items):
            arr[i].d = d
            arr[i].x = x.data_ptr()
            arr[i].gy = gy.data_ptr()
            arr[i].dw = dw.data_ptr() if dw is not None else None
            arr[i].dbias = dbias.data_ptr() if dbias is not None else None
        L.check(L.load().srb_conv_wgrad_batched(_ctx(items[0][1]), arr, len(items), _stream()), "srb_conv_wgrad_batched")


_wq = _WgradQueue()


class deferred_wgrads:
    """Context manager: defer and batch the weight-gradient launches issued inside (used around
    `loss.backward()` by srb200.trainer).  All launches happen on the current stream at exit, so
    gradients are complete before anything enqueued later on that stream reads them."""

    def __init__(self, max_items: int = 296):
        self.max_items = max_items

    def __enter__(self):
        _wq.active = True
        _wq.max_items = self.max_items
        return _wq

    def __exit__(self, *exc):
        _wq.active = False
        if exc[0] is None:
            _wq.flush()
        else:
            _wq.items = []
        return False


class WgradOverlap:
    """Weight gradients of a layer chain under the NEXT chain's backward launch.

    The per-sample cluster chain kernel (conv_cluster.cu) occupies 6 SMs per sample — 96 of 148 for the 16-patch batch — and
    is latency-bound; the batched weight-gradient kernel is throughput-bound and does not care which SMs it gets.  A chain's
    backward function wraps its conv_wgrad calls in `section()`: the calls are collected and launched on a lower-priority
    side stream right behind that chain's launch, with at most `sm_budget` CTAs (srb_set_wgrad_sm_budget), while the main
    stream goes on with the next chain.  Only the first `max_sections` sections of a backward pass are overlapped — the side
    stream falls behind (52 SMs do less than 148), and what it has not finished when the last chain ends would run on 52 SMs
    only — the rest joins the ordinary deferred queue and runs on all SMs.  `join()` makes the main stream wait for the side
    stream and drops the references that kept the operands alive.

    Reference behaviour replaced: autograd runs each layer's weight gradient right behind its input gradient on the one
    stream; the result is the same sum, accumulated into the same flat gradient buffer."""

    def __init__(self, device, sm_budget: int = 52, max_sections: int = 9, on_section_done=None):
        self.device = device
        # on_section_done(lo_ptr, hi_ptr): called on the side stream behind a section's launches with the address range of the
        # gradient buffers they wrote (data-parallel training: all-reduce that bucket under the following chains)
        self.on_section_done = on_section_done
        self.sm_budget = int(sm_budget)
        self.max_sections = int(max_sections)
        self.side = torch.cuda.Stream(device=device, priority=0)
        self.keep = []
        self.pending = None          # (queue, event) of the last section, launched by kick()
        self.adopted = False         # the layers behind the trunk have been taken over in this pass (wgrad_overlap_adopt)
        import os
        self.head_delay_ns = int(os.environ.get("SRB200_WGRAD_OVERLAP_DELAY_NS", "8000"))
        self.used = 0
        self.sections_run = 0

    def begin_pass(self):
        self.used = 0
        self.adopted = False

    def take(self) -> bool:
        if self.used >= self.max_sections:
            return False
        self.used += 1
        return True

    class _Section:
        def __init__(self, ov):
            self.ov = ov

        def __enter__(self):
            global _wq
            self.saved = _wq
            _wq = _WgradQueue()
            _wq.active = True
            _wq.max_items = 1 << 30
            return self

        def __exit__(self, *exc):
            global _wq
            q, _wq = _wq, self.saved
            if exc[0] is not None or not q.items:
                return False
            ov = self.ov
            ov.kick()                        # (an earlier section nobody kicked)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(ov.device))
            ov.keep.append(list(q.items))    # operands stay allocated until join(): the side stream still reads them
            ov.pending = (q, ev)
            return False

    def section(self):
        return WgradOverlap._Section(self)

    def kick(self):
        """Launch the pending section on the side stream.  Called right AFTER the next chain has been enqueued on the main
        stream: both only wait for the previous chain, and whichever reaches the GPU first takes its SMs first — the
        weight-gradient CTAs, spread over all GPCs, once left the cluster kernel without room for all 16 of its 6-SM
        clusters until they had finished (chain launch 375 -> 580 us, profiles/r02_step_timeline_v2_race.txt)."""
        if self.pending is None:
            return
        (q, ev), self.pending = self.pending, None
        self.side.wait_event(ev)
        lib = L.load()
        ctx = C.c_void_p(L.ctx(self.device.index))
        with torch.cuda.stream(self.side):
            if self.head_delay_ns > 0:       # let the chain launch that became runnable at the same instant take its SMs first
                L.check(lib.srb_delay(ctx, self.head_delay_ns, _stream()), "srb_delay")
            L.check(lib.srb_set_wgrad_sm_budget(ctx, self.sm_budget), "srb_set_wgrad_sm_budget")
            try:
                q.flush()
            finally:
                lib.srb_set_wgrad_sm_budget(ctx, 0)
            if self.on_section_done is not None:
                items = q.flushed
                lo = min(min(it[3].data_ptr(), it[4].data_ptr() if it[4] is not None else it[3].data_ptr()) for it in items)
                hi = max(max(it[3].data_ptr() + it[3].numel() * it[3].element_size(),
                             (it[4].data_ptr() + it[4].numel() * it[4].element_size()) if it[4] is not None else 0)
                         for it in items)
                self.on_section_done(lo, hi)
        self.sections_run += 1

    def join(self):
        self.kick()
        if self.keep:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            self.keep = []


_overlap = None


def set_wgrad_overlap(ov):
    """Install (or clear, with None) the WgradOverlap the chain backward functions use."""
    global _overlap
    _overlap = ov


def wgrad_overlap_installed() -> bool:
    return _overlap is not None and _wq.active


def wgrad_overlap_adopt():
    """Before the FIRST backward chain of a pass: the weight gradients queued so far (tail, up-sampling convs — the layers
    behind the trunk, whose backward has just run) become a side-stream section of their own, so that they run beside that
    first chain, which otherwise has nothing beside it, instead of after the last one."""
    ov = _overlap
    if ov is None or not _wq.active or ov.adopted or not _wq.items or ov.pending is not None:
        return
    import os
    # Neutral while the side stream was as long as the chain stream (N = 64 weight-gradient kernel: 7.48 vs 7.51 ms); with the
    # N = 128 kernel the side stream has ~110 us of slack per group and this takes the tail / up-sampling / head weight and bias
    # gradients off the end of the step: 7.22 -> 7.14 ms.  SRB200_WGRAD_ADOPT=0 turns it off.
    if os.environ.get("SRB200_WGRAD_ADOPT", "1") in ("", "0"):
        return
    ov.adopted = True
    q = _WgradQueue()
    q.items, _wq.items = _wq.items, []
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(ov.device))
    ov.keep.append(list(q.items))
    ov.pending = (q, ev)


def wgrad_overlap_kick():
    """To be called right after a chain launch: starts the previous chain's weight gradients beside it."""
    if _overlap is not None:
        _overlap.kick()


def wgrad_overlap_section(used_cluster: bool):
    """Context for the conv_wgrad calls of one chain: overlapped on the side stream if an overlap object is installed, the
    chain ran as the cluster kernel (it leaves SMs free) and the pass has sections left; otherwise a no-op."""
    import contextlib
    if _overlap is not None and used_cluster and _wq.active and _overlap.take():
        return _overlap.section()
    return contextlib.nullcontext()


def bn_stats(x, c, eps, momentum, running_mean=None, running_var=None):
    """Batch statistics of an NHWC tensor's first c channels: (mean [c], rstd [c]) fp32; the running statistics are updated
    in place as nn.BatchNorm2d does in training mode."""
    n, h, w, cs = _nhwc(x)
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = torch.empty(2 * c, dtype=torch.float32, device=x.device)
    L.check(L.load().srb_bn_stats(_ctx(x), _p(x), cs, 0, c, n * h * w, dtype_code(x), float(eps), float(momentum), _p(ws), _p(mean),
                                  _p(rstd), _p(running_mean), _p(running_var), _stream()), "srb_bn_stats")
    return mean, rstd


def bn_act_fwd(x, c, mean, rstd, gamma, beta, prelu_a, res, y):
    n, h, w, cs = _nhwc(x)
    L.check(L.load().srb_bn_act_fwd(_ctx(x), _p(x), cs, 0, c, n * h * w, dtype_code(x), _p(mean), _p(rstd), _p(gamma), _p(beta),
                                    _p(prelu_a), _p(res), res.shape[3] if res is not None else 0, 0, _p(y), y.shape[3], 0, _stream()),
            "srb_bn_act_fwd")
    return y


def bn_act_bwd(g, x, c, mean, rstd, gamma, beta, prelu_a, dx, dgamma, dbeta, da, accumulate):
    n, h, w, cs = _nhwc(x)
    ws = torch.empty(3 * c, dtype=torch.float32, device=x.device)
    L.check(L.load().srb_bn_act_bwd(_ctx(x), _p(g), g.shape[3], 0, _p(x), cs, 0, c, n * h * w, dtype_code(x), _p(mean), _p(rstd),
                                    _p(gamma), _p(beta), _p(prelu_a), _p(ws), _p(dx), dx.shape[3], 0, _p(dgamma), _p(dbeta), _p(da),
                                    1 if accumulate else 0, _stream()), "srb_bn_act_bwd")
    return dx


def bias_grad(gy, g_co, cout, dbias, *, accumulate=False, alpha=1.0, shuffle=0):
    """dbias (+)= alpha * per-channel sums of gy (`shuffle` > 1: gy's channels are in (sub-pixel, c') order, dbias in the
    parameter's), as a deferred item when a queue is active (it then shares the batch's one column-sum launch), else at once."""
    n, h, w, gcs = _nhwc(gy)
    d = L.WgradDesc()
    d.N, d.H, d.W = n, h, w
    d.Cin, d.Cout, d.ksize = cout, cout, 1
    d.dtype = dtype_code(gy)
    d.accumulate = 1 if accumulate else 0
    d.shuffle = shuffle
    d.backend = L.BACKEND_AUTO
    d.x_cs, d.x_co = gcs, g_co
    d.g_cs, d.g_co = gcs, g_co
    d.alpha = float(alpha)
    if _wq.active:
        _wq.push(d, gy, gy, None, dbias)
        return
    arr = (L.WgradItem * 1)()
    arr[0].d, arr[0].x, arr[0].gy, arr[0].dw, arr[0].dbias = d, gy.data_ptr(), gy.data_ptr(), None, dbias.data_ptr()
    L.check(L.load().srb_conv_wgrad_batched(_ctx(gy), arr, 1, _stream()), "srb_conv_wgrad_batched")


def conv_wgrad(x, x_co, cin, gy, g_co, cout, k, dw, dbias, *, accumulate=False, shuffle=0, alpha=1.0,
               backend=L.BACKEND_AUTO, defer=True):
    """dw (fp32 OIHW) (+)= alpha * correlation(x, gy); dbias (+)= alpha * sum(gy).
    defer=False: launch now even inside `deferred_wgrads()` — for a gradient that autograd goes on to consume (the filter
    of a weight-normalised conv is a derived tensor: its gradient feeds the g / v backward right away)."""
    lib = L.load()
    n, h, w, xcs = _nhwc(x)
    d = L.WgradDesc()
    d.N, d.H, d.W = n, h, w
    d.Cin, d.Cout, d.ksize = cin, cout, k
    d.dtype = dtype_code(x)
    d.accumulate = 1 if accumulate else 0
    d.shuffle = shuffle
    d.backend = backend
    d.x_cs, d.x_co = xcs, x_co
    d.g_cs, d.g_co = gy.shape[3], g_co
    d.alpha = float(alpha)
    if _wq.active and defer:
        _wq.push(d, x, gy, dw, dbias)
        return
    L.check(lib.srb_conv_wgrad(_ctx(x), C.byref(d), _p(x), _p(gy), _p(dw), _p(dbias), _stream()), "srb_conv_wgrad")


# --------------------------------------------------------------------------------------------
# channel attention
# --------------------------------------------------------------------------------------------
def ca_fwd(t, skip, pooled_sum, compute_pool, w1, b1, w2, b2, out, s_out, y_out):
    n, h, w, c = _nhwc(t)
    cr = w1.shape[0]
    L.check(L.load().srb_ca_fwd(_ctx(t), n, h, w, c, cr, dtype_code(t), _p(t), _p(skip), _p(pooled_sum),
                                1 if compute_pool else 0, _p(w1), _p(b1), _p(w2), _p(b2), _p(out), _p(s_out), _p(y_out),
                                _stream()), "srb_ca_fwd")


def ca_bwd(g, t, s, y, w1, b1, w2, b2, dt, dw1, db1, dw2, db2, colsum_dt, scratch, accumulate=False,
           scratch_is_zero=False):
    n, h, w, c = _nhwc(t)
    cr = w1.shape[0]
    L.check(L.load().srb_ca_bwd(_ctx(t), n, h, w, c, cr, dtype_code(t), _p(g), _p(t), _p(s), _p(y), _p(w1), _p(b1),
                                _p(w2), _p(b2), _p(dt), _p(dw1), _p(db1), _p(dw2), _p(db2), _p(colsum_dt), _p(scratch),
                                1 if scratch_is_zero else 0, 1 if accumulate else 0, _stream()), "srb_ca_bwd")


# --------------------------------------------------------------------------------------------
# layout / elementwise
# --------------------------------------------------------------------------------------------
def nchw_to_nhwc(x, chan_add, dtype, out=None, out_co=0):
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty((n, h, w, c), dtype=dtype, device=x.device)
    L.check(L.load().srb_nchw_to_nhwc(_ctx(x), _p(x), n, c, h, w, _p(chan_add), dtype_code(out), _p(out),
                                      out.shape[3], out_co, _stream()), "srb_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, x_co, c, chan_add):
    n, h, w, cs = _nhwc(x)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    L.check(L.load().srb_nhwc_to_nchw(_ctx(x), _p(x), cs, x_co, dtype_code(x), n, c, h, w, _p(chan_add), _p(out),
                                      _stream()), "srb_nhwc_to_nchw")
    return out


def copy_channels(src, s_co, dst, d_co, c):
    npix = src.shape[0] * src.shape[1] * src.shape[2]
    L.check(L.load().srb_copy_channels(_ctx(src), _p(src), src.shape[3], s_co, _p(dst), dst.shape[3], d_co, c, npix,
                                       dtype_code(src), _stream()), "srb_copy_channels")


def add_channels(a, a_co, b, b_co, out, o_co, c):
    npix = a.shape[0] * a.shape[1] * a.shape[2]
    L.check(L.load().srb_add_channels(_ctx(a), _p(a), a.shape[3], a_co, _p(b), b.shape[3], b_co, _p(out), out.shape[3],
                                      o_co, c, npix, dtype_code(a), _stream()), "srb_add_channels")


def relu_bwd(g, g_co, act, a_co, out, o_co, c):
    npix = g.shape[0] * g.shape[1] * g.shape[2]
    L.check(L.load().srb_relu_bwd(_ctx(g), _p(g), g.shape[3], g_co, _p(act), act.shape[3], a_co, _p(out), out.shape[3],
                                  o_co, c, npix, dtype_code(g), _stream()), "srb_relu_bwd")


def pixel_unshuffle(g, r, cp=None):
    """cp: logical channels of g (default: all of them); g may be wider (bf16 tensors keep pixels 16-byte aligned, so the
    3-channel output of WDSR's shuffled tail / skip convs is stored 8 wide)."""
    n, hr, wr, gcs = _nhwc(g)
    padded_out = cp is not None
    cp = gcs if cp is None else cp
    h, w = hr // r, wr // r
    ocs = cp * r * r
    if padded_out and g.dtype == torch.bfloat16 and ocs % 8:     # the consumer is a conv kernel: its tensors keep 16-byte pixels
        ocs = (ocs + 7) // 8 * 8
    out = torch.empty((n, h, w, ocs), dtype=g.dtype, device=g.device)
    L.check(L.load().srb_pixel_unshuffle(_ctx(g), _p(g), gcs, 0, _p(out), ocs, 0, n, h, w, cp, r, dtype_code(g),
                                         _stream()), "srb_pixel_unshuffle")
    return out


def colsum(x, x_co, c, out, accumulate=False):
    npix = x.shape[0] * x.shape[1] * x.shape[2]
    L.check(L.load().srb_colsum(_ctx(x), _p(x), x.shape[3], x_co, c, npix, dtype_code(x), _p(out),
                                1 if accumulate else 0, _stream()), "srb_colsum")


def l1_loss(sr, hr, want_grad=True):
    """(loss[1] fp32, grad or None): mean |sr - hr| and sign(sr-hr)/n (srmodel.py:37,549)."""
    assert sr.dtype == torch.float32 and hr.dtype == torch.float32 and sr.is_contiguous() and hr.is_contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=sr.device)
    grad = torch.empty_like(sr) if want_grad else None
    L.check(L.load().srb_l1_loss(_ctx(sr), _p(sr), _p(hr), sr.numel(), _p(loss), _p(grad), _stream()), "srb_l1_loss")
    return loss, grad


def adam_step(param, grad, m, v, *, lr, beta1, beta2, eps, weight_decay, step, step_dev=None, grad_scale=1.0):
    L.check(L.load().srb_adam_step(_ctx(param), _p(param), _p(grad), _p(m), _p(v), param.numel(), lr, beta1, beta2, eps,
                                   weight_decay, int(step), _p(step_dev), grad_scale, _stream()), "srb_adam_step")


def inc_counter(counter):
    L.check(L.load().srb_inc_counter(_ctx(counter), _p(counter), _stream()), "srb_inc_counter")


# --------------------------------------------------------------------------------------------
# layer chains (srb_conv_chain): many dependent 64-channel layers in one persistent launch
# --------------------------------------------------------------------------------------------
CHAIN_LAYER_BYTES = 9 * 64 * 64 * 2


class FilterBank:
    """Contiguous SRB_PACK_UMMA copies of the 3x3 64->64 filters a chain uses, one bank per pack
    mode (forward / DGRAD).  Layer i is a slice of one uint8 tensor; the slices are registered in
    the convs' PackedWeights caches so a PackTable refreshes them with every other packed copy."""

    def __init__(self):
        self.banks = {}

    def get(self, convs, mode: int) -> torch.Tensor:
        """convs: list of (weight parameter, PackedWeights).  Returns the bank, (re)packing stale layers."""
        w0 = convs[0][0]
        bank = self.banks.get(mode)
        if bank is None or bank.device != w0.device or bank.numel() != len(convs) * CHAIN_LAYER_BYTES:
            bank = torch.empty(len(convs) * CHAIN_LAYER_BYTES, dtype=torch.uint8, device=w0.device)
            self.banks[mode] = bank
        for i, (w, packs) in enumerate(convs):
            assert tuple(w.shape) == (64, 64, 3, 3), "chain filter banks hold 3x3 64->64 convs only"
            packs.get(w, L.PACK_UMMA, mode, 0, out=bank[i * CHAIN_LAYER_BYTES:(i + 1) * CHAIN_LAYER_BYTES])
        return bank


def chain_forward_hint() -> int:
    """Kernel for FORWARD chains (srb_chain_desc.kernel_hint).  Nothing else runs beside a forward chain, so the L2-flag
    kernel, which spreads the tiles over all SMs, is the faster one there (RCAN ResidualGroup on [16,48,48,64]: 271 us
    against 279 us); backward chains take the cluster kernel, whose free SMs run the weight gradients (WgradOverlap).
    SRB200_CHAIN_FWD=cluster|flags overrides."""
    import os
    return 0 if os.environ.get("SRB200_CHAIN_FWD", "flags") == "cluster" else 1


def chain_backward_hint() -> int:
    """Kernel for BACKWARD chains of a multi-chain model (RCAN): the cluster kernel when weight gradients run beside it
    (WgradOverlap installed, i.e. inside TrainStep), else the L2-flag kernel.  SRB200_CHAIN_BWD=cluster|flags overrides."""
    import os
    e = os.environ.get("SRB200_CHAIN_BWD", "auto")
    if e == "cluster":
        return 0
    if e == "flags":
        return 1
    return 0 if wgrad_overlap_installed() else 1


def chain_tile_flags() -> bool:
    """Layers of a chain are ordered with per-tile flags: a tile waits for the 3x3 neighbourhood of tiles of
    the previous op instead of for the slowest of its sample's tiles (RCAN step 8.25 -> 7.94 ms).
    SRB200_CHAIN_TILEFLAGS=0 selects the per-sample counters."""
    import os
    return os.environ.get("SRB200_CHAIN_TILEFLAGS", "1") not in ("", "0")


CHAIN_TRACE = None   # diagnostics: int64 device tensor that receives the event times of the next chain launches


class Chain:
    """Builder for one srb_conv_chain call.  Spaces are [slots, N, H, W, 64] bf16 tensors; `ref`
    names a slot.  Ops run in order; an op may read what earlier ops wrote."""

    def __init__(self, n: int, h: int, w: int, device):
        self.n, self.h, self.w, self.device = n, h, w, device
        self.spaces = [None] * 4
        self.ops = []
        self.keep = []
        self.used_cluster = None      # set by run(): the per-sample cluster kernel (conv_cluster.cu) took the last launch

    def space(self, index: int, t: torch.Tensor):
        assert t.dim() == 5 and t.is_contiguous() and t.dtype == torch.bfloat16 and t.shape[4] == 64
        assert tuple(t.shape[1:4]) == (self.n, self.h, self.w) and t.shape[0] < (1 << 14)
        self.spaces[index] = t
        return index

    @staticmethod
    def ref(space: int, slot: int) -> int:
        return (space << 14) | slot

    def _op(self, kind, x, y):
        o = L.ChainOp()
        o.kind, o.x, o.y, o.e, o.y2, o.e2 = kind, x, y, L.CHAIN_NONE, L.CHAIN_NONE, L.CHAIN_NONE
        o.scale, o.w_layer = 1.0, 0
        self.ops.append(o)
        return o

    def _ptr(self, t):
        if t is None:
            return None
        assert t.is_contiguous() and t.dtype == torch.float32 and t.device == self.device
        self.keep.append(t)
        return t.data_ptr()

    def conv(self, x, y, w_layer, bias=None, *, relu=False, scale=1.0, res=None, mask=None, colsum=None, colsum_groups=1,
             colsum_scale=1.0, ca_bwd=None, y_scratch=False):
        """ca_bwd: dict(t=ref, dt=ref, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch, colsum_dt) — fuse the
        CALayer backward of the block whose dL/dout this conv produces (SRB_CHAIN_CA_BWD_FUSED)."""
        o = self._op(L.CHAIN_CONV, x, y)
        if ca_bwd is not None:
            self._fill_ca_bwd(o, **ca_bwd)
        o.w_layer, o.scale = w_layer, float(scale)
        o.flags = (L.RELU if relu else 0) | (L.RESIDUAL if res is not None else 0) | (L.MASK if mask is not None else 0) | \
                  (L.COLSUM if colsum is not None else 0)
        if res is not None:
            o.e = res
        if mask is not None:
            o.e = mask
        if ca_bwd is not None:
            o.flags |= L.CHAIN_CA_BWD_FUSED
        if y_scratch:      # nothing outside the chain reads slot y (SRB_CHAIN_Y_SCRATCH)
            o.flags |= L.CHAIN_Y_SCRATCH
        o.bias = self._ptr(bias)
        o.colsum = self._ptr(colsum)
        o.colsum_groups = colsum_groups
        o.colsum_scale = float(colsum_scale)
        return o

    def conv_ca(self, x, t, out, skip, w_layer, bias, pool, w1, b1, w2, b2, s_out, y_out):
        """RCAB second conv + CALayer + skip: t = conv(x)+bias -> slot t; out = t*gate + skip."""
        o = self.conv(x, t, w_layer, bias, res=skip, colsum=pool, colsum_groups=self.n)
        o.flags |= L.CHAIN_CA
        o.y2 = out
        o.ca_cr = w1.shape[0]
        o.ca_w1, o.ca_b1, o.ca_w2, o.ca_b2 = self._ptr(w1), self._ptr(b1), self._ptr(w2), self._ptr(b2)
        o.ca_s, o.ca_y = self._ptr(s_out), self._ptr(y_out)
        return o

    def _fill_ca_params(self, o, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch):
        o.ca_cr = w1.shape[0]
        o.ca_w1, o.ca_b1, o.ca_w2, o.ca_b2 = self._ptr(w1), self._ptr(b1), self._ptr(w2), self._ptr(b2)
        o.ca_s, o.ca_y = self._ptr(s), self._ptr(y)
        o.ca_dw1, o.ca_db1, o.ca_dw2, o.ca_db2 = self._ptr(dw1), self._ptr(db1), self._ptr(dw2), self._ptr(db2)
        o.ca_scratch = self._ptr(scratch)

    def _fill_ca_bwd(self, o, t, dt, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch, colsum_dt=None):
        self._fill_ca_params(o, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch)
        o.e2, o.y2 = t, dt
        o.colsum2 = self._ptr(colsum_dt)

    def ca_bwd(self, t, g, dt, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch, colsum_dt=None):
        o = self._op(L.CHAIN_CA_BWD, t, dt)
        o.e = g
        self._fill_ca_params(o, w1, b1, w2, b2, s, y, dw1, db1, dw2, db2, scratch)
        o.colsum = self._ptr(colsum_dt)
        return o

    def run(self, bank: torch.Tensor | None, trace: torch.Tensor | None = None, hint: int = 0):
        """Launch; more than CHAIN_MAX_OPS ops are split into consecutive launches (stream order
        carries the dependency across the split)."""
        lib = L.load()
        n_layers = bank.numel() // CHAIN_LAYER_BYTES if bank is not None else 0
        for start in range(0, len(self.ops), L.CHAIN_MAX_OPS):
            seg = self.ops[start:start + L.CHAIN_MAX_OPS]
            arr = (L.ChainOp * len(seg))(*seg)
            d = L.ChainDesc()
            d.N, d.H, d.W, d.n_ops = self.n, self.h, self.w, len(seg)
            d.ops = arr
            for i, sp in enumerate(self.spaces):
                d.space_base[i] = sp.data_ptr() if sp is not None else None
                d.space_slots[i] = sp.shape[0] if sp is not None else 0
            d.weights = bank.data_ptr() if bank is not None else None
            d.n_layers = n_layers
            counters = zeros_f32((len(seg) * 2 * self.n,), self.device)     # fp32 zero bits == int32 zero
            self.keep.append(counters)
            d.counters = counters.data_ptr()
            if chain_tile_flags():
                tiles = self.n * ((self.h + 15) // 16) * ((self.w + 7) // 8)
                flags = zeros_f32((len(seg) * tiles,), self.device)
                self.keep.append(flags)
                d.tile_flags = flags.data_ptr()
            else:
                d.tile_flags = None
            if trace is None:
                trace = CHAIN_TRACE
            d.trace = trace.data_ptr() if trace is not None else None
            d.kernel_hint = hint
            self.used_cluster = bool(lib.srb_conv_chain_uses_cluster(C.byref(d)))
            L.check(lib.srb_conv_chain(C.c_void_p(L.ctx(self.device.index)), C.byref(d), _stream()), "srb_conv_chain")


def chain_grid(device, n, h, w) -> int:
    return int(L.load().srb_conv_chain_grid(C.c_void_p(L.ctx(device.index)), n, h, w))
