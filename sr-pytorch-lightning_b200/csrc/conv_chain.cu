// Layer chains: many dependent 64-channel layers of a 16 x 48x48 batch in ONE persistent launch
// (include/srb200.h "layer chains"; RCAN ResidualGroup rcan.py:59-74 forward and backward).
//
// Why: profiles/r01_trace_conv_c64_v1.txt — a dependent 3x3 64->64 layer costs ~6.5 us as its own
// kernel (launch gap + prologue ~2.2 us, filter + window import ~0.8 us, two serial 0.9 us MMA
// tiles, ~1 us epilogue, store) for 1.8 us of tensor work, and the CALayer adds two more latency-
// bound launches per RCAB.  Here:
//   * every CTA owns fixed output tiles (tile t of op i and tile t of op i+1 are the same pixels),
//     so there is no launch gap, no prologue and no TMEM allocation per layer;
//   * the CTA's tiles alternate between two CHAINS (two window buffers, two TMEM accumulators, two
//     epilogue warpgroups).  Consecutive tiles of a CTA belong to different samples, so while one
//     sample's tile is in its epilogue -> TMA store -> release flag -> neighbour's acquire -> TMA
//     load round trip (~3 us through L2), the other chain's MMAs keep the tensor pipe busy;
//   * the 72 KB filter bank of op i+1 streams in (three 24 KB kw-slabs) behind the last MMAs of op i;
//   * the CALayer is part of the conv epilogue: pooled sums by red.global, a per-sample counter, the
//     64->4->64 gate recomputed by every tile, out = t*gate + skip written from the same staging
//     buffer.  CALayer backward is a tile op of the same kernel.
// Ordering between CTAs: tile_flags[op][tile] is raised when the tile's op-`op` outputs are complete in
// global memory (TMA store completed, then a gpu-scope release); a tile's producer warp acquires the flags
// of the 3x3 neighbourhood of tiles of op-1 (nine lanes, one flag each) before it requests its input
// window, so a tile never waits for the slowest of its sample's 18 tiles (RCAN step 8.25 -> 7.94 ms against
// the older per-sample form: counters[op][0][n] == tiles_per_sample, still selectable with a NULL
// tile_flags).  CALayer ops additionally meet on counters[op][1][n] inside the op.  All CTAs walk
// (op, chain) in the same order and only wait on strictly earlier (op, chain) pairs, so there is no
// cycle; the grid is <= the SM count with one CTA per SM, i.e. all CTAs are co-resident.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 384;                    // 12 warps: 2 producers, MMA, filters, 2 x 4 epilogue
constexpr int kTW = 8, kTH = 16;                 // output tile: 128 pixels = UMMA M
constexpr int kP = kTW + 2, kRows = kTH + 2;     // input window with halo
constexpr uint32_t kWinBytes = kRows * kP * 128; // 23040
constexpr uint32_t kWinStride = 23u * 1024u;
constexpr uint32_t kSlabBytes = 3u * 64u * 128u; // one kw slab: [kh][cout][cin]
constexpr uint32_t kWBytes = 3u * kSlabBytes;    // 73728
constexpr uint32_t kTileBytes = 128u * 128u;
constexpr uint32_t kChainStride = kWinStride + 3u * kTileBytes;   // window | staging | operand tile | 2nd operand tile
constexpr uint32_t kSmemBytes = kWBytes + 2u * kChainStride + 1024u;
constexpr uint32_t kTmemCols = 128;
constexpr int kMaxCr = 16;
constexpr uint32_t kScaled = 1u << 16;    // epilogue specialisation keys: scale != 1 / no specialisation
constexpr uint32_t kGeneric = 1u << 17;

struct ChainMaps {
  CUtensorMap win[4];    // 5-D (c, w, h, n, slot), box 64 x 10 x 18
  CUtensorMap tile[4];   // same tensors, box 64 x 8 x 16
  CUtensorMap w;         // 4-D (cin, cout, tap, layer), box 64 x 64 x 3
};

struct ChainParams {
  int N, H, W;
  int tiles_w, tiles_h, tiles_per_sample, total_tiles;
  int n_ops;
  int* counters;         // [n_ops][2][N]
  int* tile_flags;       // NULL, or [n_ops][total_tiles]: per-tile completion flags instead of per-sample counts
  long long* trace;
  srb_chain_op ops[SRB_CHAIN_MAX_OPS];
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic <-> async proxy ordering for GLOBAL memory only (the all-space form also synchronises the
// shared-memory traffic of every TMA / MMA in flight in the CTA and costs several hundred ns)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Bounded spin on a device counter (a protocol bug traps instead of hanging the GPU).
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  if (ld_acquire_gpu(p) >= target) return;
  const uint64_t t0 = ptx::globaltimer_ns();
  uint32_t spins = 0;
  while (ld_acquire_gpu(p) < target) {
    if ((++spins & 0x3FFu) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
      printf("srb200: chain counter wait timed out (block %d thread %d target %d have %d)\n", blockIdx.x, threadIdx.x,
             target, ld_acquire_gpu(p));
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// diagnostics: trace[((cta * 2 + chain) * n_ops + op) * 8 + event] = globaltimer (ns)
enum { TR_DEP = 0, TR_AFULL, TR_ISSUED, TR_ACC, TR_STAGED, TR_STORED, TR_RELEASED, TR_POOL };
#define CH_TRACE(chain, op, ev)                                                                                   \
  do {                                                                                                            \
    if (p.trace) p.trace[(((size_t)blockIdx.x * 2 + (chain)) * p.n_ops + (op)) * 8 + (ev)] = (long long)ptx::globaltimer_ns(); \
  } while (0)

__device__ __forceinline__ int ref_space(uint16_t r) { return r >> 14; }
__device__ __forceinline__ int ref_slot(uint16_t r) { return r & 0x3FFF; }

// 36 MMAs of one 128-pixel tile: D[tmem_d] = sum over the nine taps (kw slab, kh) and four 16-channel
// k-steps of window(kh, kw) x filter(kw, kh).  first: the filter slabs of this op are awaited slab by
// slab; last: each slab is handed back to the filter producer once its MMAs have been issued.
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_lo, uint32_t w_lo, uint64_t* w_full,
                                         uint64_t* w_empty, bool first, bool last, uint32_t w_parity) {
  constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 0, 0);
  constexpr uint32_t hi_a = ptx::smem_desc_hi_sw128((uint32_t)kP * 128u);
  constexpr uint32_t hi_b = ptx::smem_desc_hi_sw128(1024u);
  // kw is a real loop (12 MMAs per trip): the body is executed once per tile from a cold instruction
  // cache, so a compact loop whose second and third trips hit in L0 beats 36 unrolled MMAs
#pragma unroll 1
  for (int kw = 0; kw < 3; ++kw) {
    if (first) {
      ptx::mbar_wait(&w_full[kw], w_parity);
      ptx::tc_fence_after();
    }
    const uint32_t a_kw = a_lo + (uint32_t)(kw * 8);
    const uint32_t b_kw = w_lo + (uint32_t)(kw * 3 * 512);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::umma_bf16_lohi(tmem_d, a_kw + (uint32_t)(kh * kP * 8 + k * 2), hi_a, b_kw + (uint32_t)(kh * 512 + k * 2), hi_b,
                            idesc, (kh | k) != 0 ? 1u : (kw != 0 ? 1u : 0u));
    }
    if (last) ptx::umma_commit(&w_empty[kw]);   // slab free once these MMAs have read it
  }
}

// Column sums of a staged tile ([128 pixels][64 ch] bf16, 16-byte chunks XOR-swizzled by pixel & 7) read
// back from shared memory: warp q sums its 32 rows; lane = (chunk k = lane & 7, row phase lane >> 3), eight
// conflict-free 16-byte loads per lane, two shuffle levels; lanes 0-7 then hold channels 8k..8k+7 and write
// them to out[64].  Used AFTER the tile's TMA store has been issued, i.e. off the op's critical path
// (a butterfly transpose-reduce over the epilogue registers cost ~1 us per tile).
__device__ __forceinline__ void tile_colsum_lds(uint32_t tile, int q, int lane, float* out) {
  const int k = lane & 7, rp = lane >> 3;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = q * 32 + rp + 4 * i;
    const uint4 v = ptx::lds128(tile + (uint32_t)r * 128u + (uint32_t)((k ^ (r & 7)) << 4));
    const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
    a[0] += f0.x; a[1] += f0.y; a[2] += f1.x; a[3] += f1.y; a[4] += f2.x; a[5] += f2.y; a[6] += f3.x; a[7] += f3.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
  }
  if (rp == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[k * 8 + j] = a[j];
  }
}

// Same access pattern over TWO staged tiles: column sums of the element-wise product (CALayer backward:
// sum over pixels of g * t per channel).  Out-of-image rows are zero in both tiles (TMA zero fill / the
// epilogue stages zeros), so no validity mask is needed.
__device__ __forceinline__ void tile_prod_colsum_lds(uint32_t tile_a, uint32_t tile_b, int q, int lane, float* out) {
  const int k = lane & 7, rp = lane >> 3;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = q * 32 + rp + 4 * i;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((k ^ (r & 7)) << 4);
    const uint4 u = ptx::lds128(tile_a + off);
    const uint4 v = ptx::lds128(tile_b + off);
    const float2 u0 = unpack_bf16x2(u.x), u1 = unpack_bf16x2(u.y), u2 = unpack_bf16x2(u.z), u3 = unpack_bf16x2(u.w);
    const float2 v0 = unpack_bf16x2(v.x), v1 = unpack_bf16x2(v.y), v2 = unpack_bf16x2(v.z), v3 = unpack_bf16x2(v.w);
    a[0] = fmaf(u0.x, v0.x, a[0]); a[1] = fmaf(u0.y, v0.y, a[1]); a[2] = fmaf(u1.x, v1.x, a[2]); a[3] = fmaf(u1.y, v1.y, a[3]);
    a[4] = fmaf(u2.x, v2.x, a[4]); a[5] = fmaf(u2.y, v2.y, a[5]); a[6] = fmaf(u3.x, v3.x, a[6]); a[7] = fmaf(u3.y, v3.y, a[7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
  }
  if (rp == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[k * 8 + j] = a[j];
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[3], w_empty[3];
  __shared__ uint64_t a_full[2], a_empty[2], e_full[2], e_empty[2], e2_full[2], e2_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[2][64];
  __shared__ float colsum_s[2][4][64];
  __shared__ float ca_s[2][64], ca_y[2][64], ca_du[2][64], ca_ds[2][64], ca_z[2][kMaxCr], ca_dv[2][kMaxCr];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = base;
  const int grid = (int)gridDim.x, bid = (int)blockIdx.x;
  const int my_tiles = (p.total_tiles - bid + grid - 1) / grid;
  const int N = p.N;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      ptx::mbar_init(&w_full[i], 1);
      ptx::mbar_init(&w_empty[i], 1);
    }
    for (int c = 0; c < 2; ++c) {
      ptx::mbar_init(&a_full[c], 1);
      ptx::mbar_init(&a_empty[c], 1);
      ptx::mbar_init(&e_full[c], 1);
      ptx::mbar_init(&e_empty[c], 1);
      ptx::mbar_init(&e2_full[c], 1);
      ptx::mbar_init(&e2_empty[c], 1);
      ptx::mbar_init(&acc_full[c], 1);
      ptx::mbar_init(&acc_empty[c], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;
  ptx::griddep_wait();
  if (p.trace && threadIdx.x == 0) {   // (globaltimer, clock64) pairs at start / end: the SM clock the kernel ran at
    long long* tail = p.trace + (size_t)gridDim.x * 2 * p.n_ops * 8 + (size_t)blockIdx.x * 4;
    tail[0] = (long long)ptx::globaltimer_ns();
    tail[1] = clock64();
  }

  if (warp < 2) {
    // ===================== activation producer of chain `warp` =====================
    if (p.tile_flags != nullptr) {
      // Per-tile flags: a tile's window only needs the 3x3 neighbourhood of tiles of the previous op, so it does
      // not wait for the slowest of the sample's 18 tiles.  Lanes 0-8 poll one neighbour's flag each; lane 0
      // then issues the loads exactly as in the per-sample form below.
      const int c = warp;
      const uint32_t win = base + kWBytes + (uint32_t)c * kChainStride;
      const uint32_t ebuf = win + kWinStride + kTileBytes;
      const uint32_t e2buf = ebuf + kTileBytes;
      uint32_t a_k = 0, e_k = 0, e2_k = 0;
      for (int op = 0; op < p.n_ops; ++op) {
        const srb_chain_op& o = p.ops[op];
        for (int j = c; j < my_tiles; j += 2) {
          const int t = bid + j * grid;
          const int n = t / p.tiles_per_sample;
          const int r = t - n * p.tiles_per_sample;
          const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
          const int h0 = th * kTH, w0 = tw * kTW;
          if (op > 0) {
            const int* fl = nullptr;
            if (lane < 9) {
              const int nh = th + lane / 3 - 1, nw = tw + lane % 3 - 1;
              if (nh >= 0 && nh < p.tiles_h && nw >= 0 && nw < p.tiles_w)
                fl = p.tile_flags + (size_t)(op - 1) * p.total_tiles + (size_t)n * p.tiles_per_sample + nh * p.tiles_w + nw;
            }
            bool done = fl == nullptr;
            const uint64_t t0 = ptx::globaltimer_ns();
            uint32_t spins = 0;
            while (true) {
              if (!done) done = ld_acquire_gpu(fl) >= 1;
              if (__all_sync(0xffffffffu, done)) break;
              if ((++spins & 0x3FFu) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
                if (!done) printf("srb200: chain tile-flag wait timed out (block %d op %d tile %d lane %d)\n", blockIdx.x, op, t, lane);
                __trap();
              }
            }
            if (lane == 0) fence_proxy_async_all();
          }
          if (lane == 0) {
            CH_TRACE(c, op, TR_DEP);
            ptx::mbar_wait(&a_empty[c], (a_k & 1u) ^ 1u);
            if (o.kind == SRB_CHAIN_CONV) {
              ptx::mbar_arrive_expect_tx(&a_full[c], kWinBytes);
              tma_load_5d(win, &maps.win[ref_space(o.x)], &a_full[c], 0, w0 - 1, h0 - 1, n, ref_slot(o.x));
            } else {
              ptx::mbar_arrive_expect_tx(&a_full[c], kTileBytes);
              tma_load_5d(win, &maps.tile[ref_space(o.x)], &a_full[c], 0, w0, h0, n, ref_slot(o.x));
            }
            if (o.e != SRB_CHAIN_NONE) {
              ptx::mbar_wait(&e_empty[c], (e_k & 1u) ^ 1u);
              ptx::mbar_arrive_expect_tx(&e_full[c], kTileBytes);
              tma_load_5d(ebuf, &maps.tile[ref_space(o.e)], &e_full[c], 0, w0, h0, n, ref_slot(o.e));
            }
            if (o.e2 != SRB_CHAIN_NONE) {
              ptx::mbar_wait(&e2_empty[c], (e2_k & 1u) ^ 1u);
              ptx::mbar_arrive_expect_tx(&e2_full[c], kTileBytes);
              tma_load_5d(e2buf, &maps.tile[ref_space(o.e2)], &e2_full[c], 0, w0, h0, n, ref_slot(o.e2));
            }
          }
          ++a_k;
          if (o.e != SRB_CHAIN_NONE) ++e_k;
          if (o.e2 != SRB_CHAIN_NONE) ++e2_k;
          __syncwarp();
        }
      }
    } else if (lane == 0) {
      const int c = warp;
      const uint32_t win = base + kWBytes + (uint32_t)c * kChainStride;
      const uint32_t ebuf = win + kWinStride + kTileBytes;
      const uint32_t e2buf = ebuf + kTileBytes;
      uint32_t a_k = 0, e_k = 0, e2_k = 0;
      for (int op = 0; op < p.n_ops; ++op) {
        const srb_chain_op& o = p.ops[op];
        for (int j = c; j < my_tiles; j += 2) {
          const int t = bid + j * grid;
          const int n = t / p.tiles_per_sample;
          const int r = t - n * p.tiles_per_sample;
          const int h0 = (r / p.tiles_w) * kTH, w0 = (r % p.tiles_w) * kTW;
          if (op > 0) {
            wait_counter(p.counters + ((size_t)(op - 1) * 2) * N + n, p.tiles_per_sample);
            fence_proxy_async_all();
          }
          CH_TRACE(c, op, TR_DEP);
          ptx::mbar_wait(&a_empty[c], (a_k & 1u) ^ 1u);
          if (o.kind == SRB_CHAIN_CONV) {
            ptx::mbar_arrive_expect_tx(&a_full[c], kWinBytes);
            tma_load_5d(win, &maps.win[ref_space(o.x)], &a_full[c], 0, w0 - 1, h0 - 1, n, ref_slot(o.x));
          } else {
            ptx::mbar_arrive_expect_tx(&a_full[c], kTileBytes);
            tma_load_5d(win, &maps.tile[ref_space(o.x)], &a_full[c], 0, w0, h0, n, ref_slot(o.x));
          }
          ++a_k;
          if (o.e != SRB_CHAIN_NONE) {
            ptx::mbar_wait(&e_empty[c], (e_k & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(&e_full[c], kTileBytes);
            tma_load_5d(ebuf, &maps.tile[ref_space(o.e)], &e_full[c], 0, w0, h0, n, ref_slot(o.e));
            ++e_k;
          }
          if (o.e2 != SRB_CHAIN_NONE) {
            ptx::mbar_wait(&e2_empty[c], (e2_k & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(&e2_full[c], kTileBytes);
            tma_load_5d(e2buf, &maps.tile[ref_space(o.e2)], &e2_full[c], 0, w0, h0, n, ref_slot(o.e2));
            ++e2_k;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one_sync()) {
      const uint32_t w_lo = ptx::smem_desc_lo(wbase, 16u);
      const uint32_t a_lo0 = ptx::smem_desc_lo(base + kWBytes, 16u);
      const uint32_t a_lo1 = ptx::smem_desc_lo(base + kWBytes + kChainStride, 16u);
      uint32_t a_k0 = 0, a_k1 = 0, acc_k0 = 0, acc_k1 = 0, w_k = 0;
      for (int op = 0; op < p.n_ops; ++op) {
        const srb_chain_op& o = p.ops[op];
        if (o.kind != SRB_CHAIN_CONV) {
          // Tile op without MMA.  The buffer's phase must still be OBSERVED here: a parity wait only
          // distinguishes the current phase from the one before it, so skipping a phase would let the
          // next conv's wait return on a stale completion.
          for (int j = 0; j < my_tiles; j += 2) {
            ptx::mbar_wait(&a_full[0], a_k0 & 1u);
            ++a_k0;
            if (j + 1 < my_tiles) {
              ptx::mbar_wait(&a_full[1], a_k1 & 1u);
              ++a_k1;
            }
          }
          continue;
        }
        // the two chains are written out separately so that every descriptor is a loop-invariant
        // (uniform-register) base plus an immediate: the issuing thread has ~48 cycles per MMA
        for (int j = 0; j < my_tiles; j += 2) {
          ptx::mbar_wait(&acc_empty[0], (acc_k0 & 1u) ^ 1u);
          ptx::mbar_wait(&a_full[0], a_k0 & 1u);
          ptx::tc_fence_after();
          CH_TRACE(0, op, TR_AFULL);
          mma_tile(tmem_acc, a_lo0, w_lo, w_full, w_empty, j == 0, j == my_tiles - 1, w_k & 1u);
          ptx::umma_commit(&a_empty[0]);
          ptx::umma_commit(&acc_full[0]);
          CH_TRACE(0, op, TR_ISSUED);
          ++a_k0;
          ++acc_k0;
          if (j + 1 < my_tiles) {
            ptx::mbar_wait(&acc_empty[1], (acc_k1 & 1u) ^ 1u);
            ptx::mbar_wait(&a_full[1], a_k1 & 1u);
            ptx::tc_fence_after();
            CH_TRACE(1, op, TR_AFULL);
            mma_tile(tmem_acc + 64u, a_lo1, w_lo, w_full, w_empty, false, j + 1 == my_tiles - 1, w_k & 1u);
            ptx::umma_commit(&a_empty[1]);
            ptx::umma_commit(&acc_full[1]);
            CH_TRACE(1, op, TR_ISSUED);
            ++a_k1;
            ++acc_k1;
          }
        }
        ++w_k;
      }
    }
  } else if (warp == 3) {
    // ===================== filter producer =====================
    if (lane == 0) {
      ptx::prefetch_tensormap(&maps.w);
      uint32_t w_k = 0;
      for (int op = 0; op < p.n_ops; ++op) {
        const srb_chain_op& o = p.ops[op];
        if (o.kind != SRB_CHAIN_CONV) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          ptx::mbar_wait(&w_empty[kw], (w_k & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&w_full[kw], kSlabBytes);
          ptx::tma_load_4d(wbase + (uint32_t)kw * kSlabBytes, &maps.w, &w_full[kw], 0, 0, kw * 3, o.w_layer);
        }
        ++w_k;
      }
    }
  } else {
    // ===================== epilogue warpgroup of chain c (thread = one pixel of the tile) =====================
    const int c = (warp - 4) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const bool store_thread = (q == 0 && lane == 0);
    const int bar_id = 1 + c;
    const uint32_t win = base + kWBytes + (uint32_t)c * kChainStride;
    const uint32_t stg = win + kWinStride;
    const uint32_t ebuf = stg + kTileBytes;
    const uint32_t orow = stg + (uint32_t)row * 128u;
    const uint32_t erow = ebuf + (uint32_t)row * 128u;
    const uint32_t trow = win + (uint32_t)row * 128u;   // CA_BWD: the t tile lands in the window buffer
    const uint32_t e2buf = ebuf + kTileBytes;
    const uint32_t e2row = e2buf + (uint32_t)row * 128u;
    const float inv_hw = 1.f / (float)(p.H * p.W);
    uint32_t a_k = 0, e_k = 0, e2_k = 0, acc_k = 0;

    for (int op = 0; op < p.n_ops; ++op) {
      const srb_chain_op& o = p.ops[op];
      const uint32_t flags = o.flags;
      const int Cr = o.ca_cr;
      for (int j = c; j < my_tiles; j += 2) {
        const int t = bid + j * grid;
        const int n = t / p.tiles_per_sample;
        const int r = t - n * p.tiles_per_sample;
        const int h0 = (r / p.tiles_w) * kTH, w0 = (r % p.tiles_w) * kTW;
        const bool valid = (h0 + (row >> 3) < p.H) && (w0 + (row & 7) < p.W);
        int* cnt_done = p.counters + ((size_t)op * 2) * N + n;
        int* cnt_part = p.counters + ((size_t)op * 2 + 1) * N + n;

        // CALayer + skip backward on one tile (rcan.py:10-29,54): g_row / t_row = this thread's 128-byte rows
        // of dL/dout and of the saved pre-attention tensor; dt = g*gate + ds/HW is written to out_row.
        auto ca_bwd_tile = [&](const uint32_t g_row, const uint32_t t_row, const uint32_t out_row) {
          // sum over this warp's 32 pixels of g*t per channel, read back from the two staged tiles (eight
          // conflict-free 16-byte loads per tile and lane + two shuffle levels; the register butterfly this
          // replaces was ~250 instructions per thread on the critical path of every RCAB)
          tile_prod_colsum_lds(g_row - (uint32_t)row * 128u, t_row - (uint32_t)row * 128u, q, lane, colsum_s[c][q]);
          ptx::named_bar_sync(bar_id, 128);
          if (row < 64) {
            const float tot = (colsum_s[c][0][row] + colsum_s[c][1][row]) + (colsum_s[c][2][row] + colsum_s[c][3][row]);
            atomicAdd(o.ca_scratch + (int64_t)n * 64 + row, tot);
          }
          ptx::named_bar_sync(bar_id, 128);      // all partial sums issued; the release below is cumulative over them
          if (store_thread) red_release_gpu(cnt_part, 1);
          // everything the gate backward needs that does not depend on the sample sums (cold in L1)
          const bool fast = Cr <= 4;
          float w1a = 0.f, w1b = 0.f, b1q = 0.f, w2a = 0.f, w2b = 0.f;     // this warp's hidden unit jj = q
          if (fast && q < Cr) {
            w1a = __ldg(o.ca_w1 + q * 64 + lane);
            w1b = __ldg(o.ca_w1 + q * 64 + lane + 32);
            b1q = __ldg(o.ca_b1 + q);
            w2a = __ldg(o.ca_w2 + lane * Cr + q);
            w2b = __ldg(o.ca_w2 + (lane + 32) * Cr + q);
          }
          float w2r[4] = {0.f, 0.f, 0.f, 0.f}, w1c[4] = {0.f, 0.f, 0.f, 0.f}, b2v = 0.f, sv = 0.f, yv = 0.f;
          if (row < 64) {
            b2v = __ldg(o.ca_b2 + row);
            sv = __ldg(o.ca_s + (int64_t)n * 64 + row);
            yv = __ldg(o.ca_y + (int64_t)n * 64 + row);
            if (fast)
              for (int jj = 0; jj < Cr; ++jj) {
                w2r[jj] = __ldg(o.ca_w2 + row * Cr + jj);
                w1c[jj] = __ldg(o.ca_w1 + jj * 64 + row);
              }
            ca_s[c][row] = sv;
            ca_y[c][row] = yv;
          }
          ptx::named_bar_sync(bar_id, 128);
          // hidden layer again (dW2 and the ReLU mask need it); independent of the sample sums
          if (fast) {
            if (q < Cr) {
              float a = warp_sum(w1a * ca_s[c][lane] + w1b * ca_s[c][lane + 32]);
              if (lane == 0) {
                a += b1q;
                ca_z[c][q] = fmaxf(a, 0.f);
                ca_dv[c][q] = a > 0.f ? 1.f : 0.f;
              }
            }
          } else {
            for (int jj = q; jj < Cr; jj += 4) {
              float a = __ldg(o.ca_w1 + jj * 64 + lane) * ca_s[c][lane] + __ldg(o.ca_w1 + jj * 64 + lane + 32) * ca_s[c][lane + 32];
              a = warp_sum(a);
              if (lane == 0) {
                a += __ldg(o.ca_b1 + jj);
                ca_z[c][jj] = fmaxf(a, 0.f);
                ca_dv[c][jj] = a > 0.f ? 1.f : 0.f;
              }
            }
          }
          ptx::named_bar_sync(bar_id, 128);
          float spsn = 0.f;
          if (row < 64) {
            float u = b2v;
            for (int jj = 0; jj < Cr; ++jj) u += (fast ? w2r[jj] : __ldg(o.ca_w2 + row * Cr + jj)) * ca_z[c][jj];
            const float sp = 1.f / (1.f + expf(-u)), sn = 1.f / (1.f + expf(u));
            spsn = sp * sn;       // sigmoid'(u) from u itself: no (1 - y) cancellation near saturation
          }
          if (store_thread) {
            wait_counter(cnt_part, p.tiles_per_sample);
            CH_TRACE(c, op, TR_POOL);
          }
          ptx::named_bar_sync(bar_id, 128);
          if (row < 64) ca_du[c][row] = __ldcg(o.ca_scratch + (int64_t)n * 64 + row) * spsn;
          ptx::named_bar_sync(bar_id, 128);
          if (fast) {
            if (q < Cr) {
              const float dz = warp_sum(w2a * ca_du[c][lane] + w2b * ca_du[c][lane + 32]);
              if (lane == 0) ca_dv[c][q] *= dz;
            }
          } else {
            for (int jj = q; jj < Cr; jj += 4) {
              float dz = __ldg(o.ca_w2 + lane * Cr + jj) * ca_du[c][lane] + __ldg(o.ca_w2 + (lane + 32) * Cr + jj) * ca_du[c][lane + 32];
              dz = warp_sum(dz);
              if (lane == 0) ca_dv[c][jj] *= dz;
            }
          }
          ptx::named_bar_sync(bar_id, 128);
          if (row < 64) {
            float d = 0.f;
            for (int jj = 0; jj < Cr; ++jj) d += (fast ? w1c[jj] : __ldg(o.ca_w1 + jj * 64 + row)) * ca_dv[c][jj];
            ca_ds[c][row] = d * inv_hw;
          }
          if (r == 0) {                              // parameter gradients, once per sample
            for (int i = row; i < 64 * Cr; i += 128) {
              atomicAdd(o.ca_dw2 + i, ca_du[c][i / Cr] * ca_z[c][i % Cr]);    // w2 [64][Cr]
              atomicAdd(o.ca_dw1 + i, ca_dv[c][i / 64] * ca_s[c][i % 64]);    // w1 [Cr][64]
            }
            if (row < 64) atomicAdd(o.ca_db2 + row, ca_du[c][row]);
            if (row < Cr) atomicAdd(o.ca_db1 + row, ca_dv[c][row]);
          }
          ptx::named_bar_sync(bar_id, 128);
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t packed[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t off = (((uint32_t)(c0 >> 3) + g) ^ sw) << 4;
              const uint4 gv = ptx::lds128(g_row + off);
              const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 fg = unpack_bf16x2(gw[e]);
                const int ch = c0 + g * 8 + e * 2;
                packed[g * 4 + e] = valid ? pack_bf16x2(fmaf(fg.x, ca_y[c][ch], ca_ds[c][ch]),
                                                        fmaf(fg.y, ca_y[c][ch + 1], ca_ds[c][ch + 1]))
                                          : 0u;
              }
              ptx::sts128(out_row + off, make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]));
            }
          }
        };
        // column sums of a tile already handed to a TMA store (read-only here), accumulated into dst[64]
        auto late_colsum = [&](const uint32_t tile, float* dst, const float factor) {
          tile_colsum_lds(tile, q, lane, colsum_s[c][q]);
          ptx::named_bar_sync(bar_id, 128);
          if (row < 64) {
            const float tot = (colsum_s[c][0][row] + colsum_s[c][1][row]) + (colsum_s[c][2][row] + colsum_s[c][3][row]);
            atomicAdd(dst + row, factor != 0.f ? tot * factor : tot);
          }
        };

        if (o.kind == SRB_CHAIN_CONV) {
          const bool has_e = o.e != SRB_CHAIN_NONE;
          const bool ca = (flags & SRB_CHAIN_CA) != 0;
          // the bias of this op is cold in L1 (every op has its own): fetch it into shared memory
          // while the MMAs run instead of stalling each 32-column chunk on an L2 round trip
          if (row < 64) bias_s[c][row] = o.bias ? __ldg(o.bias + row) : 0.f;
          ptx::named_bar_sync(bar_id, 128);
          ptx::mbar_wait(&acc_full[c], acc_k & 1u);
          ptx::tc_fence_after();
          if (store_thread) CH_TRACE(c, op, TR_ACC);
          if (has_e) ptx::mbar_wait(&e_full[c], e_k & 1u);
          // Epilogue arithmetic, specialised at compile time on the op's flag combination: the block runs
          // once per tile from a cold instruction cache, and every skipped `if (flags & ...)` section was
          // a taken branch plus an instruction refetch (ncu: no_inst + branch_resolving = half the stalls).
          auto epilogue = [&](auto FC) {
            constexpr uint32_t F = decltype(FC)::value;
            uint32_t acc2[2][32];
            ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 64), acc2[0]);
            ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 64 + 32), acc2[1]);
            ptx::tmem_ld_wait();
            const uint32_t fl = (F == kGeneric) ? flags : F;
            const float scale = o.scale;
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
              const int c0 = hc * 32;
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc2[hc][i]);
#pragma unroll
              for (int i = 0; i < 32; i += 4) {     // bias_s holds zeros when the op has no bias
                const float4 bq = *reinterpret_cast<const float4*>(&bias_s[c][c0 + i]);
                v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
              }
              if (fl & SRB_RELU) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if ((F == kGeneric || (F & kScaled)) && scale != 1.f) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= scale;
              }
              if (fl & SRB_MASK) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const uint4 m = ptx::lds128(erow + ((((uint32_t)(c0 >> 3) + g) ^ sw) << 4));
                  const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16x2(mw[e]);
                    if (!(f.x > 0.f)) v[g * 8 + e * 2] = 0.f;
                    if (!(f.y > 0.f)) v[g * 8 + e * 2 + 1] = 0.f;
                  }
                }
              }
              if ((fl & SRB_RESIDUAL) && !(fl & SRB_CHAIN_CA)) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const uint4 m = ptx::lds128(erow + ((((uint32_t)(c0 >> 3) + g) ^ sw) << 4));
                  const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16x2(mw[e]);
                    v[g * 8 + e * 2] += f.x;
                    v[g * 8 + e * 2 + 1] += f.y;
                  }
                }
              }
              uint32_t packed[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) packed[i] = valid ? pack_bf16x2(v[2 * i], v[2 * i + 1]) : 0u;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                ptx::sts128(orow + ((((uint32_t)(c0 >> 3) + g) ^ sw) << 4),
                            make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]));
            }
          };
          switch ((flags & 47u) | (o.scale != 1.f ? kScaled : 0u)) {
            case SRB_RELU: epilogue(std::integral_constant<uint32_t, SRB_RELU>{}); break;
            case SRB_RESIDUAL: epilogue(std::integral_constant<uint32_t, SRB_RESIDUAL>{}); break;
            case SRB_RESIDUAL | kScaled: epilogue(std::integral_constant<uint32_t, SRB_RESIDUAL | kScaled>{}); break;
            case SRB_MASK | SRB_COLSUM: epilogue(std::integral_constant<uint32_t, SRB_MASK | SRB_COLSUM>{}); break;
            case SRB_MASK | SRB_COLSUM | kScaled: epilogue(std::integral_constant<uint32_t, SRB_MASK | SRB_COLSUM | kScaled>{}); break;
            case SRB_RESIDUAL | SRB_COLSUM | SRB_CHAIN_CA:
              epilogue(std::integral_constant<uint32_t, SRB_RESIDUAL | SRB_COLSUM | SRB_CHAIN_CA>{});
              break;
            case 0: epilogue(std::integral_constant<uint32_t, 0>{}); break;
            default: epilogue(std::integral_constant<uint32_t, kGeneric>{}); break;
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (ca) {
            // CALayer pool (needed before anything else can happen): sums of the STORED (bf16-rounded) values,
            // read back from the staged tile (the register butterfly this replaces cost ~1 us per tile); the
            // four warps' partial sums meet in shared memory so that a tile issues 64 atomics, not 256
            tile_colsum_lds(stg, q, lane, colsum_s[c][q]);
            ptx::named_bar_sync(bar_id, 128);
            if (row < 64) {
              const float tot = (colsum_s[c][0][row] + colsum_s[c][1][row]) + (colsum_s[c][2][row] + colsum_s[c][3][row]);
              atomicAdd(o.colsum + (int64_t)n * 64 + row, tot);
            }
            ptx::named_bar_sync(bar_id, 128);   // pooled-sum contributions precede the cumulative release below
          }
          if (!ca) {
            if (store_thread) {
              ptx::mbar_arrive(&acc_empty[c]);
              if (has_e) ptx::mbar_arrive(&e_empty[c]);
              tma_store_5d(&maps.tile[ref_space(o.y)], stg, 0, w0, h0, n, ref_slot(o.y));
              ptx::bulk_commit_group();
              CH_TRACE(c, op, TR_STAGED);
            }
            if (flags & SRB_COLSUM)
              late_colsum(stg, o.colsum + (int64_t)(o.colsum_groups > 1 ? n : 0) * 64, o.colsum_scale);
            if (flags & SRB_CHAIN_CA_BWD_FUSED) {
              // y (just staged, bf16) is dL/dout of the previous RCAB: run its CALayer backward here
              // instead of as a dependent op.  The staging buffer is only READ (the store of y may
              // still be reading it too); dt overwrites the t tile and is stored from there.
              ptx::mbar_wait(&e2_full[c], e2_k & 1u);
              ca_bwd_tile(orow, e2row, e2row);
              ptx::fence_proxy_async_smem();
              ptx::named_bar_sync(bar_id, 128);
              if (store_thread) {
                tma_store_5d(&maps.tile[ref_space(o.y2)], e2buf, 0, w0, h0, n, ref_slot(o.y2));
                ptx::bulk_commit_group();
              }
              if (o.colsum2) late_colsum(e2buf, o.colsum2, 0.f);
              ++e2_k;
            }
          } else {
            // ---- CALayer gate + RCAB skip (rcan.py:10-29,54) on the tile still in the staging buffer ----
            // The pooled-sum contributions of all 128 threads precede the barrier above; the release
            // below (one thread, gpu scope) is cumulative over them — the cutlass::Barrier::arrive_inc
            // pattern — so no per-thread fence is needed.  It is issued BEFORE the store of t so that it
            // does not wait behind it.
            if (store_thread) {
              ptx::mbar_arrive(&acc_empty[c]);
              red_release_gpu(cnt_part, 1);
              tma_store_5d(&maps.tile[ref_space(o.y)], stg, 0, w0, h0, n, ref_slot(o.y));
              ptx::bulk_commit_group();
              CH_TRACE(c, op, TR_STAGED);
            }
            // gate operands do not depend on the pool: fetch them (cold in L1 — every acquire poll
            // invalidates it) while the other tiles of the sample arrive
            float w1a[2] = {0.f, 0.f}, w1b[2] = {0.f, 0.f}, b1v[2] = {0.f, 0.f};
            {
              int u = 0;
              for (int jj = q; jj < Cr && u < 2; jj += 4, ++u) {
                w1a[u] = __ldg(o.ca_w1 + jj * 64 + lane);
                w1b[u] = __ldg(o.ca_w1 + jj * 64 + lane + 32);
                b1v[u] = __ldg(o.ca_b1 + jj);
              }
            }
            float w2r[4] = {0.f, 0.f, 0.f, 0.f}, b2v = 0.f;
            if (row < 64) {
              b2v = __ldg(o.ca_b2 + row);
              for (int jj = 0; jj < Cr && jj < 4; ++jj) w2r[jj] = __ldg(o.ca_w2 + row * Cr + jj);
            }
            if (store_thread) {
              wait_counter(cnt_part, p.tiles_per_sample);
              CH_TRACE(c, op, TR_POOL);
            }
            ptx::named_bar_sync(bar_id, 128);
            if (row < 64) ca_s[c][row] = __ldcg(o.colsum + (int64_t)n * 64 + row) * inv_hw;
            ptx::named_bar_sync(bar_id, 128);
            {
              int u = 0;
              for (int jj = q; jj < Cr; jj += 4, ++u) {
                float a = (u < 2 ? w1a[u] : __ldg(o.ca_w1 + jj * 64 + lane)) * ca_s[c][lane] +
                          (u < 2 ? w1b[u] : __ldg(o.ca_w1 + jj * 64 + lane + 32)) * ca_s[c][lane + 32];
                a = warp_sum(a);
                if (lane == 0) ca_z[c][jj] = fmaxf(a + (u < 2 ? b1v[u] : __ldg(o.ca_b1 + jj)), 0.f);
              }
            }
            ptx::named_bar_sync(bar_id, 128);
            if (row < 64) {
              float u = b2v;
              for (int jj = 0; jj < Cr; ++jj) u += (jj < 4 ? w2r[jj] : __ldg(o.ca_w2 + row * Cr + jj)) * ca_z[c][jj];
              const float yv = 1.f / (1.f + expf(-u));
              ca_y[c][row] = yv;
              if (r == 0) {
                o.ca_s[(int64_t)n * 64 + row] = ca_s[c][row];
                o.ca_y[(int64_t)n * 64 + row] = yv;
              }
            }
            ptx::named_bar_sync(bar_id, 128);
            // out = t * gate + skip, written over the skip tile (dead afterwards): the staging buffer may
            // still be read by the store of t
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t off = (((uint32_t)(c0 >> 3) + g) ^ sw) << 4;
                const uint4 tv = ptx::lds128(orow + off);
                const uint4 xv = ptx::lds128(erow + off);
                const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w}, xw[4] = {xv.x, xv.y, xv.z, xv.w};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 ft = unpack_bf16x2(tw[e]), fx = unpack_bf16x2(xw[e]);
                  const int ch = c0 + g * 8 + e * 2;
                  pk[e] = pack_bf16x2(fmaf(ft.x, ca_y[c][ch], fx.x), fmaf(ft.y, ca_y[c][ch + 1], fx.y));
                }
                ptx::sts128(erow + off, make_uint4(pk[0], pk[1], pk[2], pk[3]));
              }
            }
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(bar_id, 128);
            if (store_thread) {
              tma_store_5d(&maps.tile[ref_space(o.y2)], ebuf, 0, w0, h0, n, ref_slot(o.y2));
              ptx::bulk_commit_group();
            }
          }
          ++acc_k;
          if (has_e) ++e_k;
        } else {
          // ---- CALayer + skip backward (tile op, no MMA): x = t tile (window buffer), e = g tile ----
          ptx::mbar_wait(&a_full[c], a_k & 1u);
          ptx::mbar_wait(&e_full[c], e_k & 1u);
          ca_bwd_tile(erow, trow, orow);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (store_thread) {
            ptx::mbar_arrive(&a_empty[c]);
            ptx::mbar_arrive(&e_empty[c]);
            tma_store_5d(&maps.tile[ref_space(o.y)], stg, 0, w0, h0, n, ref_slot(o.y));
            ptx::bulk_commit_group();
          }
          if (o.colsum) late_colsum(stg, o.colsum, 0.f);
          ++e_k;
        }
        ++a_k;
        // ---- publish: outputs complete in global memory, then release the sample counter ----
        if (store_thread) {
          ptx::bulk_wait_group<0>();
          if (o.kind == SRB_CHAIN_CONV && (flags & SRB_CHAIN_CA)) ptx::mbar_arrive(&e_empty[c]);   // out was staged in it
          if (o.kind == SRB_CHAIN_CONV && (flags & SRB_CHAIN_CA_BWD_FUSED)) ptx::mbar_arrive(&e2_empty[c]);   // so was dt
          CH_TRACE(c, op, TR_STORED);
          fence_proxy_async_all();            // async-proxy (TMA) writes ordered before the generic-proxy release
          // release.gpu: no separate __threadfence (a MEMBAR.SC costs ~1 us)
          if (p.tile_flags != nullptr) red_release_gpu(p.tile_flags + (size_t)op * p.total_tiles + t, 1);
          else red_release_gpu(cnt_done, 1);
          CH_TRACE(c, op, TR_RELEASED);
        }
        ptx::named_bar_sync(bar_id, 128);          // staging buffer free before the next item writes it
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0) {
    long long* tail = p.trace + (size_t)gridDim.x * 2 * p.n_ops * 8 + (size_t)blockIdx.x * 4;
    tail[2] = (long long)ptx::globaltimer_ns();
    tail[3] = clock64();
  }
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_space(srb_ctx* ctx, CUtensorMap* map, void* ptr, int slots, int N, int H, int W, int box_w, int box_h) {
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)slots};
  cuuint64_t strides[4] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128, (cuuint64_t)N * H * W * 128};
  cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    srb_set_error("srb_conv_chain: cuTensorMapEncodeTiled failed with CUresult %d (slots=%d N=%d H=%d W=%d box %dx%d)", (int)r,
                  slots, N, H, W, box_w, box_h);
    return 4;
  }
  return 0;
}

int chain_grid(int num_sms, int N, int H, int W) {
  const int64_t tiles = (int64_t)N * srb_cdiv(W, kTW) * srb_cdiv(H, kTH);
  int64_t g = (tiles + 1) / 2;        // two tiles (one per chain) per CTA when the batch allows it
  if (g > num_sms) g = num_sms;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int srb_conv_chain_grid(const srb_ctx* ctx, int N, int H, int W) {
  if (!ctx || N < 1 || H < 1 || W < 1) return 0;
  return chain_grid(ctx->num_sms, N, H, W);
}

// conv_cluster.cu: one thread-block cluster per sample, halos through distributed shared memory
int srb_chain_cluster_eligible(const srb_chain_desc* d);
int srb_chain_cluster_launch(srb_ctx* ctx, const srb_chain_desc* d, void* stream);

extern "C" int srb_conv_chain_uses_cluster(const srb_chain_desc* d) { return d && d->ops ? srb_chain_cluster_eligible(d) : 0; }

extern "C" int srb_conv_chain(srb_ctx* ctx, const srb_chain_desc* d, void* stream) {
  SRB_REQUIRE(ctx && d && d->ops && d->counters, "srb_conv_chain: null argument");
  SRB_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "srb_conv_chain: empty shape N=%d H=%d W=%d", d->N, d->H, d->W);
  SRB_REQUIRE(d->n_ops >= 1 && d->n_ops <= SRB_CHAIN_MAX_OPS, "srb_conv_chain: n_ops %d outside [1,%d]", d->n_ops,
              SRB_CHAIN_MAX_OPS);
  SRB_REQUIRE(d->N < (1 << 14), "srb_conv_chain: batch too large");
  if (srb_chain_cluster_eligible(d)) return srb_chain_cluster_launch(ctx, d, stream);
  static_assert(sizeof(ChainParams) + sizeof(ChainMaps) < 32000, "kernel parameters exceed the 32 KB limit");
  ChainParams* pp = new ChainParams();
  ChainMaps* mm = new ChainMaps();
  struct Guard {
    ChainParams* a;
    ChainMaps* b;
    ~Guard() {
      delete a;
      delete b;
    }
  } guard{pp, mm};
  ChainParams& p = *pp;
  ChainMaps& maps = *mm;
  p.N = d->N;
  p.H = d->H;
  p.W = d->W;
  p.tiles_w = srb_cdiv(d->W, kTW);
  p.tiles_h = srb_cdiv(d->H, kTH);
  p.tiles_per_sample = p.tiles_w * p.tiles_h;
  const int64_t tiles = (int64_t)d->N * p.tiles_per_sample;
  SRB_REQUIRE(tiles < (1ll << 30), "srb_conv_chain: too many tiles");
  p.total_tiles = (int)tiles;
  p.n_ops = d->n_ops;
  p.counters = d->counters;
  p.tile_flags = d->tile_flags;
  p.trace = reinterpret_cast<long long*>(d->trace);

  // A CALayer op makes a tile's epilogue wait for every tile of its sample; two tiles of one sample
  // on the same (CTA, chain) would wait for each other.  Tile t runs on CTA t % grid, chain
  // (t / grid) & 1, so tiles_per_sample <= 2 * grid keeps a sample's tiles on distinct (CTA, chain)s.
  const int grid_for_check = chain_grid(ctx->num_sms, d->N, d->H, d->W);
  const bool sample_sync_ok = p.tiles_per_sample <= 2 * grid_for_check;
  bool used[4] = {false, false, false, false};
  auto check_ref = [&](uint16_t r, bool required, const char* what, int op) -> int {
    if (r == SRB_CHAIN_NONE) {
      SRB_REQUIRE(!required, "srb_conv_chain: op %d needs a %s buffer", op, what);
      return 0;
    }
    const int sp = r >> 14, slot = r & 0x3FFF;
    SRB_REQUIRE(d->space_base[sp] != nullptr && slot < d->space_slots[sp],
                "srb_conv_chain: op %d %s reference (space %d, slot %d) outside the declared spaces", op, what, sp, slot);
    used[sp] = true;
    return 0;
  };
  bool any_conv = false;
  for (int i = 0; i < d->n_ops; ++i) {
    const srb_chain_op& o = d->ops[i];
    p.ops[i] = o;
    int rc;
    if ((rc = check_ref(o.x, true, "x", i))) return rc;
    if ((rc = check_ref(o.y, true, "y", i))) return rc;
    if (o.kind == SRB_CHAIN_CONV) {
      any_conv = true;
      SRB_REQUIRE(o.w_layer >= 0 && o.w_layer < d->n_layers, "srb_conv_chain: op %d filter index %d outside [0,%d)", i,
                  o.w_layer, d->n_layers);
      const bool m = (o.flags & SRB_MASK) != 0, r = (o.flags & SRB_RESIDUAL) != 0;
      SRB_REQUIRE(!(m && r), "srb_conv_chain: op %d: MASK and RESIDUAL together are not supported", i);
      SRB_REQUIRE(!(o.flags & SRB_OUT2), "srb_conv_chain: op %d: OUT2 is not supported", i);
      if ((rc = check_ref(o.e, m || r, "mask/residual", i))) return rc;
      SRB_REQUIRE((m || r) || o.e == SRB_CHAIN_NONE, "srb_conv_chain: op %d has an operand tile but no MASK/RESIDUAL flag", i);
      SRB_REQUIRE(!(o.flags & SRB_COLSUM) || (o.colsum && (o.colsum_groups == 1 || o.colsum_groups == d->N)),
                  "srb_conv_chain: op %d: COLSUM needs a pointer and groups in {1, N}", i);
      SRB_REQUIRE((o.e2 != SRB_CHAIN_NONE) == ((o.flags & SRB_CHAIN_CA_BWD_FUSED) != 0),
                  "srb_conv_chain: op %d: a second operand tile goes with CA_BWD_FUSED and only with it", i);
      if (o.flags & SRB_CHAIN_CA_BWD_FUSED) {
        SRB_REQUIRE(sample_sync_ok, "srb_conv_chain: op %d: CA ops need tiles_per_sample (%d) <= 2 x grid (%d); use srb_ca_bwd", i,
                    p.tiles_per_sample, grid_for_check);
        SRB_REQUIRE(!(o.flags & SRB_CHAIN_CA), "srb_conv_chain: op %d: CA and CA_BWD_FUSED are exclusive", i);
        if ((rc = check_ref(o.e2, true, "saved t", i))) return rc;
        if ((rc = check_ref(o.y2, true, "dt output", i))) return rc;
        SRB_REQUIRE(o.ca_w1 && o.ca_b1 && o.ca_w2 && o.ca_b2 && o.ca_s && o.ca_y && o.ca_dw1 && o.ca_db1 && o.ca_dw2 &&
                        o.ca_db2 && o.ca_scratch && o.ca_cr >= 1 && o.ca_cr <= kMaxCr,
                    "srb_conv_chain: op %d: CA_BWD_FUSED pointers missing or Cr outside [1,%d]", i, kMaxCr);
      }
      if (o.flags & SRB_CHAIN_CA) {
        SRB_REQUIRE(sample_sync_ok, "srb_conv_chain: op %d: CA ops need tiles_per_sample (%d) <= 2 x grid (%d); use srb_ca_fwd",
                    i, p.tiles_per_sample, grid_for_check);
        SRB_REQUIRE((o.flags & SRB_COLSUM) && o.colsum_groups == d->N && r && !(o.flags & SRB_RELU) && o.scale == 1.f,
                    "srb_conv_chain: op %d: CA needs COLSUM per sample, RESIDUAL, no ReLU, scale 1", i);
        SRB_REQUIRE(o.ca_w1 && o.ca_b1 && o.ca_w2 && o.ca_b2 && o.ca_s && o.ca_y && o.ca_cr >= 1 && o.ca_cr <= kMaxCr,
                    "srb_conv_chain: op %d: CA parameters missing or Cr outside [1,%d]", i, kMaxCr);
        if ((rc = check_ref(o.y2, true, "CA output", i))) return rc;
      }
    } else if (o.kind == SRB_CHAIN_CA_BWD) {
      SRB_REQUIRE(o.e2 == SRB_CHAIN_NONE, "srb_conv_chain: op %d: CA_BWD takes no second operand tile", i);
      SRB_REQUIRE(sample_sync_ok, "srb_conv_chain: op %d: CA ops need tiles_per_sample (%d) <= 2 x grid (%d); use srb_ca_bwd", i,
                  p.tiles_per_sample, grid_for_check);
      if ((rc = check_ref(o.e, true, "gradient", i))) return rc;
      SRB_REQUIRE(o.ca_w1 && o.ca_b1 && o.ca_w2 && o.ca_b2 && o.ca_s && o.ca_y && o.ca_dw1 && o.ca_db1 && o.ca_dw2 &&
                      o.ca_db2 && o.ca_scratch && o.ca_cr >= 1 && o.ca_cr <= kMaxCr,
                  "srb_conv_chain: op %d: CA_BWD pointers missing or Cr outside [1,%d]", i, kMaxCr);
    } else {
      SRB_REQUIRE(false, "srb_conv_chain: op %d has unknown kind %d", i, o.kind);
    }
  }
  SRB_REQUIRE(!any_conv || (d->weights && d->n_layers > 0), "srb_conv_chain: conv ops need a filter bank");

  for (int s = 0; s < 4; ++s) {
    if (!used[s]) {
      // unused spaces still need a valid descriptor object; alias the first used one
      continue;
    }
    SRB_REQUIRE(((uintptr_t)d->space_base[s] & 127) == 0, "srb_conv_chain: space %d must be 128-byte aligned", s);
    int rc = encode_space(ctx, &maps.win[s], d->space_base[s], d->space_slots[s], d->N, d->H, d->W, kP, kRows);
    if (rc) return rc;
    rc = encode_space(ctx, &maps.tile[s], d->space_base[s], d->space_slots[s], d->N, d->H, d->W, kTW, kTH);
    if (rc) return rc;
  }
  int first = -1;
  for (int s = 0; s < 4; ++s)
    if (used[s]) {
      first = s;
      break;
    }
  for (int s = 0; s < 4; ++s)
    if (!used[s]) {
      maps.win[s] = maps.win[first];
      maps.tile[s] = maps.tile[first];
    }
  if (any_conv) {
    SRB_REQUIRE(((uintptr_t)d->weights & 127) == 0, "srb_conv_chain: filter bank must be 128-byte aligned");
    cuuint64_t dims[4] = {64, 64, 9, (cuuint64_t)d->n_layers};
    cuuint64_t strides[3] = {128, 64 * 128, 9 * 64 * 128};
    cuuint32_t box[4] = {64, 64, 3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
    CUresult r = fn(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->weights), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "srb_conv_chain: cuTensorMapEncodeTiled(filters) failed with CUresult %d", (int)r);
  } else {
    maps.w = maps.tile[first];
  }

  SRB_REQUIRE((int)kSmemBytes + 1024 <= ctx->smem_optin, "srb_conv_chain: needs %u bytes of shared memory, device offers %d",
              kSmemBytes, ctx->smem_optin);
  static bool attr_set = false;
  if (!attr_set) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  const int grid = chain_grid(ctx->num_sms, d->N, d->H, d->W);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  // The CTAs spin-wait on each other's tile flags, so ALL of them must be resident at once.  grid <= #SMs with one CTA per SM
  // makes that true on an idle GPU, but not beside another stream's kernel (weight gradients, NCCL) or under MPS / a profiler:
  // a cooperative launch makes the driver hold the kernel back until the whole grid fits instead of letting the waits run into
  // their 2 s bound.  SRB200_CHAIN_COOP=0 launches plainly (A/B timing).
  static const bool coop = [] {
    const char* e = getenv("SRB200_CHAIN_COOP");
    return !(e && e[0] == '0');
  }();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = coop ? 1 : 0;
  SRB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, maps, p));
  SRB_LAUNCH_CHECK();
  return 0;
}
