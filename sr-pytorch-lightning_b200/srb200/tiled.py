"""Large-image inference, spatially tiled across GPUs (BASELINE.json config 5: EDSR x4,
960x540 -> 3840x2160 on 8 GPUs).  The reference has no tiling at all (`SRModel.predict_step`,
/root/reference/models/srmodel.py:375-380, pushes the whole frame through `forward` on one device);
this is the multi-GPU form of that same `forward` (edsr.py:40-54).

Partition: row strips of the LR frame (W kept whole so strip rows are contiguous NHWC slabs).
Every strip buffer carries `t` halo rows above and below its owned rows.  A conv runs over the whole
buffer; the outermost output rows are incomplete and are overwritten by the **per-layer halo
exchange**: each strip sends its first/last `t` owned rows to its neighbours.  True image borders
keep zeros in the halo (the conv's zero padding), so the result is bit-identical to the untiled
forward (same kernels, same per-pixel summation order) — tests/test_tiled_gpu.py checks equality.

Exchange back-ends: `LocalExchange` (all strips on this GPU; used for testing and for single-GPU
chopping of frames that do not fit), `DistExchange` (one strip per rank, NCCL send/recv of the
contiguous row slabs over NVLink, batched per layer, eager) and `PeerExchange` (one strip per rank; the
layer buffers live in CUDA-IPC memory that the neighbours map, ONE kernel per layer pushes the border rows
straight into the neighbours' halo rows through NVLink and synchronises with flags in peer memory —
csrc/halo.cu — so that the whole strip forward is a CUDA graph without any host or NCCL call per layer).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

import ctypes as C

from . import lib as L
from . import ops
from .functional import _padded


def partition_rows(h: int, parts: int):
    """[(r0, r1)] contiguous row ranges, sizes differing by at most one, larger strips first."""
    base, extra = divmod(h, parts)
    out, r = [], 0
    for i in range(parts):
        n = base + (1 if i < extra else 0)
        out.append((r, r + n))
        r += n
    return out


class LocalExchange:
    """All strips live in this process: the exchange is row copies between their buffers."""

    def __init__(self, parts: int):
        self.parts = parts

    def strips(self):
        return list(range(self.parts))

    def exchange(self, bufs, t: int):
        """bufs[i]: [1, hs_i + 2t, W, C]; fills each buffer's halo rows from its neighbours."""
        for i in range(self.parts):
            b = bufs[i]
            if i > 0:
                up = bufs[i - 1]
                b[:, :t].copy_(up[:, up.shape[1] - 2 * t: up.shape[1] - t])
            else:
                b[:, :t].zero_()
            if i < self.parts - 1:
                dn = bufs[i + 1]
                b[:, b.shape[1] - t:].copy_(dn[:, t:2 * t])
            else:
                b[:, b.shape[1] - t:].zero_()


class DistExchange:
    """One strip per rank of `group`; neighbours trade halo rows with batched P2P ops."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.parts = dist.get_world_size(group)

    def strips(self):
        return [self.rank]

    def exchange(self, bufs, t: int):
        b = bufs[0]
        ops_ = []
        hs2 = b.shape[1]
        up, dn = self.rank - 1, self.rank + 1
        if up >= 0:
            ops_.append(dist.P2POp(dist.isend, b[:, t:2 * t], dist.get_global_rank(self.group, up) if self.group else up, self.group))
            ops_.append(dist.P2POp(dist.irecv, b[:, :t], dist.get_global_rank(self.group, up) if self.group else up, self.group))
        else:
            b[:, :t].zero_()
        if dn < self.parts:
            ops_.append(dist.P2POp(dist.isend, b[:, hs2 - 2 * t: hs2 - t], dist.get_global_rank(self.group, dn) if self.group else dn, self.group))
            ops_.append(dist.P2POp(dist.irecv, b[:, hs2 - t:], dist.get_global_rank(self.group, dn) if self.group else dn, self.group))
        else:
            b[:, hs2 - t:].zero_()
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()


class _RawCuda:
    """Device memory that torch did not allocate, exposed through __cuda_array_interface__ (uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerExchange:
    """One strip per rank; halo rows travel by peer stores over NVLink (csrc/halo.cu, srb_halo_exchange).

    Every exchanged layer buffer is carved, in the fixed order of the forward pass, out of one CUDA-IPC allocation per rank
    (`alloc`); buffers are laid out for the LARGEST strip so that a buffer has the same offset on every rank, and each rank
    maps its two neighbours' allocations.  `exchange` = one kernel: handshake ("my layer buffer is complete, you may
    write its halo rows"), push of the first / last t owned rows, release flags carrying the frame number, acquire of the
    neighbours' flags.  Nothing else: no NCCL call, no host synchronisation — `TiledEDSR.capture` records the strip forward
    as one CUDA graph.  The arena is sized by a first pass that counts what `alloc` is asked for (`TiledEDSR` does it)."""

    FLAG_BYTES = 256 * 32        # 256 layer slots x {from above: ready, data; from below: ready, data} int64

    def __init__(self, group=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.parts = dist.get_world_size(group)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = L.load()
        self.ctx = C.c_void_p(L.ctx(self.device.index))
        self.base = None             # this rank's arena (device pointer), uint8 tensor view in self.bytes
        self.up = self.dn = None     # neighbours' arenas as mapped here
        self.planning = True
        self.plan_bytes = 0
        self.cursor = 0
        self.layer = 0
        self.live = {}               # data_ptr -> offset of the buffers handed out in this frame
        self.frame = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.done = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.rows = None             # owned LR rows per rank (set by TiledEDSR)
        self.halo = 1                # halo rows at LR resolution (set by TiledEDSR)

    def strips(self):
        return [self.rank]

    # ---- arena ----
    def begin_frame(self):
        self.cursor = self.FLAG_BYTES
        self.layer = 0
        self.live = {}
        if not self.planning:
            L.check(self.lib.srb_inc_counter64(self.ctx, C.c_void_p(self.frame.data_ptr()), ops._stream()), "srb_inc_counter64")

    def alloc(self, shape, dtype, device, max_rows: int):
        """A [1, rows, W, C] buffer; `max_rows` = rows of the largest strip's buffer for this layer (layout is per layer, not
        per rank).  Planning pass: an ordinary tensor, and the size is counted."""
        n, rows, w, c = shape
        esz = torch.empty((), dtype=dtype).element_size()
        nbytes = (n * max_rows * w * c * esz + 255) // 256 * 256
        off = self.cursor
        self.cursor += nbytes
        if self.planning:
            self.plan_bytes = max(self.plan_bytes, self.cursor)
            return torch.empty(shape, dtype=dtype, device=device)
        assert self.cursor <= self.bytes.numel(), "PeerExchange: the forward pass asks for more than the planning pass did"
        t = self.bytes[off:off + n * rows * w * c * esz].view(dtype).view(shape)
        self.live[t.data_ptr()] = off
        return t

    def commit(self):
        """End of the planning pass: allocate the arena, trade IPC handles, map the neighbours."""
        assert self.planning
        total = torch.tensor([self.plan_bytes], dtype=torch.int64, device=self.device)
        dist.all_reduce(total, op=dist.ReduceOp.MAX, group=self.group)
        nbytes = int(total.item())
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        L.check(self.lib.srb_ipc_alloc(self.ctx, C.c_size_t(nbytes), C.byref(ptr), handle), "srb_ipc_alloc")
        self.base = ptr.value
        self.bytes = torch.as_tensor(_RawCuda(self.base, nbytes), device=self.device)
        handles = [None] * self.parts
        dist.all_gather_object(handles, handle.raw, group=self.group)

        def open_(r):
            q = C.c_void_p()
            L.check(self.lib.srb_ipc_open(self.ctx, C.c_char_p(handles[r]), C.byref(q)), "srb_ipc_open")
            return q.value
        self.up = open_(self.rank - 1) if self.rank > 0 else None
        self.dn = open_(self.rank + 1) if self.rank + 1 < self.parts else None
        self.planning = False
        dist.barrier(group=self.group)

    # ---- per layer ----
    def exchange(self, bufs, t: int):
        b = bufs[0]
        if self.planning:
            self.layer += 1
            return
        off = self.live.get(b.data_ptr())
        assert off is not None, "PeerExchange.exchange: the buffer does not come from alloc()"
        _, hs2, w, c = b.shape
        row_bytes = w * c * b.element_size()
        slab = t * row_bytes
        up, dn = self.rank > 0, self.rank + 1 < self.parts
        if not up:
            b[:, :t].zero_()
        if not dn:
            b[:, hs2 - t:].zero_()
        assert self.layer < 256
        d = L.HaloDesc()
        d.src_top = b.data_ptr() + t * row_bytes
        d.src_bot = b.data_ptr() + (hs2 - 2 * t) * row_bytes
        d.slab_bytes = slab
        fl = self.layer * 32
        if up:
            scale = t // self.halo      # t grows with the up-sampling factor exactly as the owned rows do
            up_rows = self.rows[self.rank - 1] * scale
            d.dst_up = self.up + off + (up_rows + t) * row_bytes       # its bottom halo rows
            d.flag_up = self.up + fl + 16                              # its "from below" pair
            d.wait_up = self.base + fl                                 # my "from above" pair
        if dn:
            d.dst_dn = self.dn + off                                   # its top halo rows
            d.flag_dn = self.dn + fl                                   # its "from above" pair
            d.wait_dn = self.base + fl + 16                            # my "from below" pair
        d.frame = self.frame.data_ptr()
        d.done = self.done.data_ptr()
        L.check(self.lib.srb_halo_exchange(self.ctx, C.byref(d), ops._stream()), "srb_halo_exchange")
        self.layer += 1

    def close(self):
        if self.base is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for q in (self.up, self.dn):
            if q:
                self.lib.srb_ipc_close(self.ctx, C.c_void_p(q))
        dist.barrier(group=self.group)
        del self.bytes
        self.lib.srb_ipc_free(self.ctx, C.c_void_p(self.base))
        self.base = self.up = self.dn = None


class TiledEDSR:
    """Strip-parallel `EDSR.forward` (no autograd).  `model` is a models.EDSR on this device."""

    def __init__(self, model, exchange, halo: int | None = None):
        """halo: rows of halo each strip buffer carries at LR resolution.  A 3x3 conv uses up one VALID halo row per side, an
        exchange restores all `halo` of them, so with halo = 2 only every second layer needs an exchange (33 instead of 69
        for EDSR-large; none at the up-sampled resolutions, where the rows are 2-4x longer); the extra halo row is computed
        redundantly by both neighbours (same kernel, same inputs: same bits).  68 + 4 rows are still nine 8-row tiles."""
        self.m = model
        self.ex = exchange
        self.halo = halo            # None: chosen per frame height, see forward()
        self.graph = None

    def _conv(self, conv, xs, *, relu=False, scale=1.0, residual=None, shuffle=0, exchanged=True):
        outs = []
        for i, x in enumerate(xs):
            n, h, w, _ = x.shape
            cout, cin = conv.weight.shape[0], conv.weight.shape[1]
            r = shuffle if shuffle > 1 else 1
            shape = (n, h * r, w * r, _padded(cout // (r * r), x.dtype))
            if hasattr(self.ex, "alloc") and exchanged:
                y = self.ex.alloc(shape, x.dtype, x.device, max_rows=(self._max_rows * (self._t // self.halo) + 2 * self._t) * r)
            else:
                y = torch.empty(shape, dtype=x.dtype, device=x.device)
            b = conv.packs.get_bias(conv.bias, shuffle)
            # The first / last strip end at the image border: their outer halo rows lie outside the image.  The conv runs on the
            # row slab WITHOUT them, so that its own zero padding is the image border (no clearing of those rows after every layer,
            # and fewer rows to compute); nothing ever reads them.
            top = self._t if self._mine[i] == 0 else 0
            bot = self._t if self._mine[i] == self._nparts - 1 else 0
            xv, yv = x[:, top:h - bot], y[:, top * r:(h - bot) * r]
            rv = residual[i][:, top:h - bot] if residual is not None else None
            ops.conv(xv, 0, cin, conv.packs, conv.weight, b, yv, 0, cout, conv.kernel_size, relu=relu, scale=scale,
                     shuffle=shuffle, res=(rv, 0) if rv is not None else None)
            outs.append(y)
        return outs

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """x: full LR frame [1,3,H,W] fp32 on this device (every rank holds it: it is 6 MB).
        Returns {strip index: SR rows [1,3,(r1-r0)*s,W*s] fp32} for the strips this process owns."""
        m = self.m
        assert x.shape[0] == 1, "tiled inference takes one frame at a time"
        H = x.shape[2]
        parts = partition_rows(H, self.ex.parts)
        mine = self.ex.strips()
        rows_max, rows_min = max(b - a for a, b in parts), min(b - a for a, b in parts)
        if self.halo is None:
            # two halo rows halve the number of exchanges; take them when they do not add an 8-row tile to the largest strip
            self.halo = 2 if rows_min >= 2 and -(-(rows_max + 4) // 8) == -(-(rows_max + 2) // 8) else 1
        t = self.halo
        assert t >= 1 and rows_min >= t, "strips must own at least `halo` rows"
        self._max_rows = max(b - a for a, b in parts)
        self._t = t
        self._mine, self._nparts = list(mine), len(parts)
        if hasattr(self.ex, "begin_frame"):
            self.ex.rows = [b - a for a, b in parts]
            self.ex.halo = self.halo
            self.ex.begin_frame()
        rgb = m._channels == 3
        add_in = m.sub_mean.channel_add() if rgb else None
        # input strips with halo straight from the frame (zero outside the image, AFTER mean shift)
        xs = []
        for i in mine:
            r0, r1 = parts[i]
            lo, hi = max(r0 - t, 0), min(r1 + t, H)
            cs = _padded(x.shape[1], m.act_dtype)
            buf = torch.zeros((1, (r1 - r0) + 2 * t, x.shape[3], cs), dtype=m.act_dtype, device=x.device)
            rows = buf[:, (lo - (r0 - t)):(lo - (r0 - t)) + (hi - lo)]     # contiguous row slab of buf
            ops.nchw_to_nhwc(x[:, :, lo:hi].contiguous(), add_in, m.act_dtype, out=rows)
            xs.append(buf)
        # valid halo rows per side of a buffer list: the input strips carry `t` rows of real data (or the image border's zeros)
        valid = {id(xs): t}

        def conv(cv, src, *, residual=None, **kw):
            if valid[id(src)] == 0:                      # the conv would read stale halo rows: refresh all t of them first
                self.ex.exchange(src, self._t)
                valid[id(src)] = self._t
            out = self._conv(cv, src, residual=residual, **kw)
            v = valid[id(src)] - 1
            if residual is not None:
                v = min(v, valid[id(residual)])
            r = kw.get("shuffle", 0)
            if r > 1:                                    # PixelShuffle: every row becomes r rows
                v *= r
                self._t *= r
            valid[id(out)] = v
            return out
        h0 = conv(m.head[0], xs)
        res = h0
        blocks = list(m.body)
        for blk in blocks[:-1]:
            c1, c2 = blk.body[0], blk.body[2]
            y1 = conv(c1, res, relu=True)
            res = conv(c2, y1, scale=blk.res_scale, residual=res)
        res = conv(blocks[-1], res, residual=h0)
        y = res
        up = list(m.tail[0])
        for cv, ps in zip(up[0::2], up[1::2]):
            y = conv(cv, y, shuffle=ps.upscale_factor)
        y = conv(m.tail[1], y, exchanged=False)
        t = self._t
        out = {}
        add_out = m.add_mean.channel_add() if rgb else None
        for i, yy in zip(mine, y):
            owned = yy[:, t:yy.shape[1] - t].contiguous()
            out[i] = ops.nhwc_to_nchw(owned, 0, m._channels, add_out)
        return out

    @torch.no_grad()
    def prepare(self, x: torch.Tensor, use_graph: bool = True):
        """PeerExchange only: planning pass (sizes the IPC arena; no data moves), arena + neighbour mapping, warm-up, and —
        if `use_graph` — capture of the strip forward on a static copy of `x` as one CUDA graph.  Every rank calls it."""
        ex = self.ex
        assert isinstance(ex, PeerExchange)
        if ex.planning:
            self.forward(x)
            ex.commit()
        self.static_x = x.clone()
        for _ in range(2):
            self.static_out = self.forward(self.static_x)
        torch.cuda.synchronize()
        if use_graph:
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            c0 = L.launch_count()
            with torch.cuda.graph(g, stream=side):
                self.static_out = self.forward(self.static_x)
            self.launches_per_frame = L.launch_count() - c0      # this library's kernels inside one replay
            torch.cuda.current_stream().wait_stream(side)
            self.graph = g
        return self

    @torch.no_grad()
    def run(self, x: torch.Tensor):
        """After prepare(): one frame; returns the same dict as forward() (static tensors, valid until the next run)."""
        self.static_x.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.static_out = self.forward(self.static_x)
        return self.static_out

    @torch.no_grad()
    def forward_gathered(self, x: torch.Tensor):
        """Full SR frame on every rank (all_gather of the strips) / on this process (local)."""
        out = self.forward(x)
        if isinstance(self.ex, LocalExchange):
            return torch.cat([out[i] for i in sorted(out)], dim=2)
        s = self.m._scale_factor
        parts = partition_rows(x.shape[2], self.ex.parts)
        mine = out[self.ex.rank]
        bufs = [torch.empty((1, mine.shape[1], (r1 - r0) * s, mine.shape[3]), dtype=mine.dtype, device=mine.device)
                for r0, r1 in parts]
        dist.all_gather(bufs, mine, group=self.ex.group) if len({b.shape for b in bufs}) == 1 else \
            [dist.broadcast(bufs[i] if i != self.ex.rank else mine,
                            src=dist.get_global_rank(self.ex.group, i) if self.ex.group is not None else i, group=self.ex.group)
             for i in range(self.ex.parts)]
        bufs[self.ex.rank] = mine
        return torch.cat(bufs, dim=2)
