// Layer chains, per-sample thread-block clusters (round 2; the L2-flag form is conv_chain.cu).
//
// Why: profiles/r01_chain_bench_v9.txt — in conv_chain.cu a dependent 64->64 layer costs 5.0 us per op, of which
// 2.8 us is the L2 round trip between CTAs (TMA store complete -> red.release -> acquire poll -> TMA load) and only
// 2 x 1.15 us is tensor work; the two sample-wide reductions of an RCAB (CALayer pool, sum g*t) add 3.5 / 4.7 us.
// Everything that couples tiles on this path is INTRA-sample (the 1-pixel halo of a 3x3 conv, the CALayer pool), and a
// sample ([48,48,64] bf16 = 288 KB per tensor) fits the shared memory of a few SMs.  So here:
//   * one cluster per sample: (H/16) bands x (W/24) halves of CTAs (48x48: 3 x 2 = 6), a CTA owns a 16 x 24 pixel
//     region = three 16x8 UMMA tiles; clusters never synchronise with each other (no co-residency assumption);
//   * the activation lives in shared memory, as the next layer's A operand: two [18][26]-pixel buffers with halo
//     (128-byte swizzled rows, the layout tcgen05.mma and TMA use); op i reads X[i & 1], its epilogue writes
//     X[(i + 1) & 1]: the centre with st.shared, the border pixels ALSO into the neighbours' halos with
//     st.shared::cluster, then mbarrier.arrive.release.cluster on the neighbours' "ready" barriers (measured hop:
//     0.30 us, profiles/r02_hw_probes_p7.txt; the L2 path it replaces: 2.8 us);
//   * saved activations (what backward / the weight gradients need) are copied X[out] -> global AFTER the tile has been
//     published, by all 256 epilogue threads with 8 lanes per 128-byte pixel line (the first version stored them from
//     the epilogue registers, one 16-byte piece per lane in 32 different lines per instruction: 32 LSU wavefronts
//     instead of 4, and the epilogue — not the tensor pipe — set the pace: profiles/r02_cluster_trace_v1.txt);
//   * operands: a residual that is the block input is still in X[out] and is read in place; a residual produced two ops
//     earlier (the dL/dout chain of the backward pass) is parked in TMEM by its producer (tcgen05.st) and read back
//     with tcgen05.ld; saved forward tensors (ReLU mask, pre-attention t) arrive as TMA tiles in one 16 KB buffer;
//   * CALayer (rcan.py:10-29) and its backward: per-thread running sums over the CTA's three tiles, ONE warp
//     transpose-reduce, a 64-float DSMEM all-gather among the cluster, the 64->Cr->64 gate evaluated by every CTA in
//     the same (rank) order, then a second pass applies it.  Forward re-reads t from the accumulators still in TMEM.
//   * tiles are processed boundary-first (left half: 2,1,0; right half: 0,1,2) and readiness is tracked with two
//     barriers per buffer — A: "positions 0-1 of every contributor are in" (enables position 0 of the next op),
//     B: "everything is in" — so the next op's first tile runs under the epilogue of this op's last tile.
// Plain conv ops issue the same MMA sequence and epilogue arithmetic as conv_chain.cu / conv_c64.cu: results are
// bit-identical (tests/test_cluster_chain_gpu.py).
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 384;                       // 12 warps: operand tiles, MMA, filters, spare, 8 x epilogue
constexpr int kEpi = 256;
constexpr int kPub = kEpi + 32;                     // epilogue threads + the publisher warp
constexpr int kBarTile = 2, kBarPair = 6;           // named barriers: 2 + q per tile position (publisher); 6 + quarter (warp pairs)
constexpr int kXW = 26, kXH = 18;                   // activation buffer: 24 + 2 columns, 16 + 2 rows
constexpr uint32_t kXBytes = kXH * kXW * 128u;      // 59904
constexpr uint32_t kXStride = 60u * 1024u;
constexpr uint32_t kSlabBytes = 3u * 64u * 128u;    // one kw slab: [kh][cout][cin]
constexpr uint32_t kWBytes = 3u * kSlabBytes;       // 73728
constexpr uint32_t kEBytes = 128u * 128u;           // one operand tile
constexpr uint32_t kTmemCols = 512;                 // 0-255: four 64-column accumulators; 256-447: two parking areas
constexpr uint32_t kParkBase = 256, kParkArea = 96; //          (3 tiles x 32 columns of packed bf16 pairs each)
constexpr int kMaxCr = 16;
constexpr int kMaxCluster = 8;
constexpr uint32_t kScaled = 1u << 16;    // epilogue specialisation keys: scale != 1 / no specialisation
constexpr uint32_t kGeneric = 1u << 17;

enum { RES_NONE = 0, RES_INPLACE = 1, RES_TMEM = 2, RES_GLOBAL = 3 };

struct COp {
  uint32_t flags;
  int32_t w_layer;
  float scale, colsum_scale;
  int32_t colsum_groups, ca_cr;
  const float* bias;
  float* colsum;
  float* colsum2;
  const float *ca_w1, *ca_b1, *ca_w2, *ca_b2;
  float *ca_s, *ca_y, *ca_dw1, *ca_db1, *ca_dw2, *ca_db2;
  uint8_t* y;            // slot base addresses
  uint8_t* y2;
  const uint8_t* e;      // residual slot (RES_GLOBAL)
  uint16_t t_ref;        // operand tile by TMA: ReLU mask (MASK) or saved t (CA_BWD_FUSED); SRB_CHAIN_NONE if none
  uint8_t res_mode;      // RES_*
  uint8_t park;          // keep the packed result y in TMEM for the op after next (its residual)
  uint16_t y_ref, y2_ref;   // (space << 14) | slot of y / y2 for the TMA row stores
};

struct CParams {
  int N, H, W, bands, halves, n_ops, x0_slot;
  long long* trace;      // diagnostics: [grid][n_ops][32] globaltimer ns (CL_TRACE)
  COp ops[SRB_CHAIN_MAX_OPS];
};

// diagnostics: trace[(cta * n_ops + op) * 32 + event] = globaltimer (ns)
//  16-23: epilogue warp w arrives on the tile barrier of q1; 24: publisher released by it; 25: publisher's arrives issued
//  MMA thread: 0 position-0 window ready, 1 position-0 MMAs issued, 2 position-1 window ready, 3 all MMAs issued
//  epilogue thread 0: 4 acc(q0) full, 5 q0 loaded from TMEM, 6 q0 stores issued, 7 q0 proxy fence done, 8 q0 published,
//                     9 acc(q1) full, 10 q1 published, 11 acc(q2) full, 12 q2 published, 13 op complete, 14 sums gathered
#define CL_TRACE(op, ev)                                                                                      \
  do {                                                                                                        \
    if (p.trace) p.trace[((size_t)blockIdx.x * p.n_ops + (op)) * 32 + (ev)] = (long long)ptx::globaltimer_ns(); \
  } while (0)

struct CMaps {
  CUtensorMap x0;        // 5-D (c, w, h, n, slot) over the space of op 0's input, box 64 x 26 x 18
  CUtensorMap w;         // 4-D (cin, cout, tap, layer), box 64 x 64 x 3
  CUtensorMap tile[4];   // per space: box 64 x 8 x 16 (operand tiles)
  CUtensorMap row[4];    // per space: box 64 x 8 x 1 (one tile row: X[out] -> global)
};

struct Small {
  uint64_t w_full[3], w_empty[3], acc_full[4], acc_empty[4], x_full, e_full, e_empty;
  uint64_t ready[2][2];          // [buffer][A / B]
  uint64_t pool_full[2];
  uint32_t tmem_slot, pad;
  float bias[2][64];
  float wsum[8][32];
  float pool[2][kMaxCluster][64];
  float ca_s[64], ca_y[64], ca_du[64], ca_ds[64], ca_tot[64], ca_z[kMaxCr], ca_dv[kMaxCr];
};

constexpr uint32_t kSmemBytes = kWBytes + 2u * kXStride + kEBytes + (uint32_t)sizeof(Small) + 1024u;

// ---- cluster / DSMEM primitives ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
// Remote store that carries its own completion: the bytes are counted on an mbarrier of the DESTINATION CTA when they have
// landed (complete_tx, release at cluster scope).  The consumer's barrier expects the byte count of a phase, so no thread
// here has to run a cluster-scope release fence after its halo stores (that fence — MEMBAR + ERRBAR + CGAERRBAR — waited
// for the CTA's in-flight global / TMA stores too and cost ~0.85 us per tile, profiles/r02_cluster_trace_v5.txt).
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
// release at cluster scope: cumulative over everything ordered before it in this CTA (the arriving thread has been
// through a bar.sync with the writers) — the cluster-scope form of conv_chain.cu's red.release.gpu after a barrier
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a protocol bug traps (sticky launch error) instead of hanging the GPU.  Only waits INSIDE a cluster exist.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = ptx::smem_u32(bar);
  if (mbar_try_wait_cluster(a, parity)) return;
  const uint64_t t0 = ptx::globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(a, parity)) {
    if ((++spins & 0x3FFu) == 0 && ptx::globaltimer_ns() - t0 > 2000000000ull) {
      printf("srb200: cluster chain wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, a, parity);
      __trap();
    }
  }
}
// generic-proxy writes ordered before async-proxy reads (tcgen05.mma operands).  Writers fence their LOCAL stores
// (the standard st.shared -> fence -> barrier -> UMMA pattern); remote halo stores become visible through the
// release/acquire pair on the consumer's ready barrier, and the consumer's MMA thread runs its own proxy fence between
// that acquire and the MMAs.  SRB_CLUSTER_STRICT_FENCE (compile time) makes every writer fence at cluster scope too
// (+0.4 us per tile, profiles/r02_cluster_trace_v1.txt).
__device__ __forceinline__ void fence_writer() {
#ifdef SRB_CLUSTER_STRICT_FENCE
  asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
#else
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// producer side of a named barrier: does not wait (the publisher warp is the only bar.sync participant)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ldg128(const uint8_t* p) {      // plain (coherent) 16-byte global load
  uint4 v;
  asm volatile("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg128(uint8_t* p, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
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

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp owns lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 36 MMAs of one tile (same order as conv_chain.cu mma_tile: kw slab, kh, four 16-channel k-steps)
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_lo, uint32_t w_lo, uint64_t* w_full, uint64_t* w_empty,
                                         bool first, bool last, uint32_t w_parity) {
  constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 0, 0);
  constexpr uint32_t hi_a = ptx::smem_desc_hi_sw128((uint32_t)kXW * 128u);
  constexpr uint32_t hi_b = ptx::smem_desc_hi_sw128(1024u);
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
        ptx::umma_bf16_lohi(tmem_d, a_kw + (uint32_t)(kh * kXW * 8 + k * 2), hi_a, b_kw + (uint32_t)(kh * 512 + k * 2), hi_b, idesc,
                            (kh | k) != 0 ? 1u : (kw != 0 ? 1u : 0u));
    }
    if (last) ptx::umma_commit(&w_empty[kw]);
  }
}

// Sum of v[c] over the 32 lanes of a warp for 32 values at once: after five exchange rounds lane L holds the total of
// value index L (31 shuffles; bit k of the lane selects bit k of the index).
template <int N>
__device__ __forceinline__ void tr_round(float (&v)[32], int lane) {
  const bool up = (lane & (N / 2)) != 0;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = up ? v[i] : v[i + N / 2];
    const float keep = up ? v[i + N / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N / 2);
  }
}
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
  tr_round<32>(v, lane);
  tr_round<16>(v, lane);
  tr_round<8>(v, lane);
  tr_round<4>(v, lane);
  tr_round<2>(v, lane);
  return v[0];
}

__device__ __forceinline__ void unpack4(const uint4 (&ev)[4], float (&f)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t w[4] = {ev[g].x, ev[g].y, ev[g].z, ev[g].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16x2(w[e]);
      f[g * 8 + e * 2] = t.x;
      f[g * 8 + e * 2 + 1] = t.y;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
chain_cluster_kernel(const __grid_constant__ CMaps maps, const __grid_constant__ CParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t wbase = base;
  const uint32_t xbase = base + kWBytes;                       // X[0]; X[1] = + kXStride
  const uint32_t ebase = xbase + 2u * kXStride;                // operand tile / staging
  Small& S = *reinterpret_cast<Small*>(base_ptr + kWBytes + 2u * kXStride + kEBytes);

  const int halves = p.halves, bands = p.bands;
  const int csize = halves * bands;
  const int rank = (int)cluster_ctarank();
  const int n = (int)blockIdx.x / csize;
  const int band = rank / halves, f = rank - band * halves;
  const bool has_up = band > 0, has_dn = band < bands - 1, has_side = halves == 2;
  const int up_rank = rank - halves, dn_rank = rank + halves, side_rank = band * halves + (1 - f);
  const bool mirror = has_side && f == 0;                      // boundary-first order: tile index of position q

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      ptx::mbar_init(&S.w_full[i], 1);
      ptx::mbar_init(&S.w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&S.acc_full[i], 1);
      ptx::mbar_init(&S.acc_empty[i], 8);
    }
    ptx::mbar_init(&S.x_full, 1);
    ptx::mbar_init(&S.e_full, 1);
    ptx::mbar_init(&S.e_empty, 8);
    // ready[buffer][A]: this CTA's positions 0 and 1 (two local arrivals) + the neighbours' halo bytes of their positions
    // 0 and 1 (st.async complete_tx); ready[buffer][B]: position 2 (one local arrival) + the neighbours' position-2 halo bytes
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&S.ready[b][0], 2);
      ptx::mbar_init(&S.ready[b][1], 1);
      ptx::mbar_init(&S.pool_full[b], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&S.tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  // X[1] starts as zeros: its image-border halo is the conv's zero padding and is never written afterwards
  // (X[0] gets the same from the TMA load's out-of-bounds fill)
  for (uint32_t i = threadIdx.x; i < kXBytes / 16u; i += kThreads) ptx::sts128(xbase + kXStride + i * 16u, make_uint4(0, 0, 0, 0));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers initialised and buffers cleared in every CTA before any remote store / arrive
  ptx::tc_fence_after();
  const uint32_t tmem_acc = S.tmem_slot;
  const float inv_hw = 1.f / (float)(p.H * p.W);

  if (warp == 0) {
    // ===================== initial window + operand tiles (saved forward tensors) by TMA =====================
    if (lane == 0) {
      ptx::prefetch_tensormap(&maps.x0);
      ptx::mbar_arrive_expect_tx(&S.x_full, kXBytes);
      tma_load_5d(xbase, &maps.x0, &S.x_full, 0, f * 24 - 1, band * 16 - 1, n, p.x0_slot);
      uint32_t e_k = 0;
      for (int op = 0; op < p.n_ops; ++op) {
        const uint16_t r = p.ops[op].t_ref;
        if (r == SRB_CHAIN_NONE) continue;
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          ptx::mbar_wait(&S.e_empty, (e_k & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&S.e_full, kEBytes);
          tma_load_5d(ebase, &maps.tile[r >> 14], &S.e_full, 0, f * 24 + 8 * j, band * 16, n, r & 0x3FFF);
          ++e_k;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one_sync()) {
      const uint32_t w_lo = ptx::smem_desc_lo(wbase, 16u);
      for (int op = 0; op < p.n_ops; ++op) {
        const int ib = op & 1;
        const uint32_t rdy_par = (uint32_t)((op - 1) >> 1) & 1u;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          const uint32_t g = (uint32_t)(3 * op + q), slot = g & 3u;
          ptx::mbar_wait(&S.acc_empty[slot], ((g >> 2) & 1u) ^ 1u);
          if (op == 0) {
            if (q == 0) ptx::mbar_wait(&S.x_full, 0);
          } else if (q == 0) {
            mbar_wait_cluster(&S.ready[ib][0], rdy_par);
          } else if (q == 1) {
            mbar_wait_cluster(&S.ready[ib][1], rdy_par);
          }
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_after();
          if (q < 2) CL_TRACE(op, 2 * q);
          const uint32_t a_lo = ptx::smem_desc_lo(xbase + (uint32_t)ib * kXStride + (uint32_t)(8 * j) * 128u, 16u);
          mma_tile(tmem_acc + slot * 64u, a_lo, w_lo, S.w_full, S.w_empty, q == 0, q == 2, (uint32_t)op & 1u);
          ptx::umma_commit(&S.acc_full[slot]);
          if (q != 1) CL_TRACE(op, q == 0 ? 1 : 3);
        }
      }
    }
  } else if (warp == 2) {
    // ===================== filter producer =====================
    if (lane == 0) {
      ptx::prefetch_tensormap(&maps.w);
      for (int op = 0; op < p.n_ops; ++op) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          ptx::mbar_wait(&S.w_empty[kw], ((uint32_t)op & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&S.w_full[kw], kSlabBytes);
          ptx::tma_load_4d(wbase + (uint32_t)kw * kSlabBytes, &maps.w, &S.w_full[kw], 0, 0, kw * 3, p.ops[op].w_layer);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== publisher =====================
    // The epilogue warps only ARRIVE on a named barrier once a tile sits in X[out]; this warp waits for it and arrives on this
    // CTA's own ready barrier, so that no epilogue warp ever blocks on a tile barrier.
    // Local arrivals only (CTA scope): the neighbours' halo bytes are counted by the st.async stores themselves.
    const uint32_t nv = (has_up ? 1u : 0u) + (has_dn ? 1u : 0u);
    const uint32_t ra_bytes = nv * 2048u + (has_side ? 2048u + nv * 128u : 0u);      // rows of positions 0-1, side column, corners
    const uint32_t rb_bytes = nv * 1024u;                                            // rows of position 2
    for (int op = 0; op < p.n_ops; ++op) {
      const int ob = (op + 1) & 1;
      for (int q = 0; q < 3; ++q) {
        ptx::named_bar_sync(kBarTile + q, kPub);
        if (lane == 0) {
          if (q == 1) CL_TRACE(op, 24);
          if (q == 0) ptx::mbar_arrive_expect_tx(&S.ready[ob][0], ra_bytes);
          else if (q == 1) ptx::mbar_arrive(&S.ready[ob][0]);
          else ptx::mbar_arrive_expect_tx(&S.ready[ob][1], rb_bytes);
          CL_TRACE(op, q == 0 ? 8 : (q == 1 ? 10 : 12));
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps, thread = (pixel of the tile, 32-channel half) =====================
    const int et = (int)threadIdx.x - 128;          // 0..255
    const int w8 = warp - 4;                        // 0..7
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int h = w8 >> 2;                          // channel half
    const int row = quarter * 32 + lane;            // pixel within the tile
    const int ty = row >> 3, tx = row & 7;
    const uint32_t tm_lane = (uint32_t)(quarter * 32) << 16;
    // shared-memory addresses of this pixel for tile 0 of X[0] (tile j: + j * 1024, buffer 1: + kXStride)
    const uint32_t idx_l = (uint32_t)((ty + 1) * kXW + tx + 1);
    const uint32_t xl = xbase + idx_l * 128u, sw_l = idx_l & 7u;
    const uint32_t el = ebase + (uint32_t)row * 128u;            // this pixel's row of the operand tile (swizzle = tx)
    // remote halo targets (buffer 0) and the destination CTA's ready barrier A of buffer 0 (buffer 1: + 16, barrier B: + 8)
    uint32_t up_a = 0, dn_a = 0, side_a = 0, dgu_a = 0, dgd_a = 0, sw_up = 0, sw_dn = 0, sw_side = 0, sw_dg = 0;
    uint32_t up_b = 0, dn_b = 0, side_b = 0, dgu_b = 0, dgd_b = 0;
    const uint32_t rdy0 = ptx::smem_u32(&S.ready[0][0]);
    if (has_up && ty == 0) {
      const uint32_t idx = (uint32_t)(17 * kXW + tx + 1);
      up_a = mapa(xbase + idx * 128u, (uint32_t)up_rank);
      up_b = mapa(rdy0, (uint32_t)up_rank);
      sw_up = idx & 7u;
    }
    if (has_dn && ty == 15) {
      const uint32_t idx = (uint32_t)(tx + 1);
      dn_a = mapa(xbase + idx * 128u, (uint32_t)dn_rank);
      dn_b = mapa(rdy0, (uint32_t)dn_rank);
      sw_dn = idx & 7u;
    }
    if (has_side && tx == (f == 0 ? 7 : 0)) {      // only used at position 0 (the tile next to the other half)
      const uint32_t col = f == 0 ? 0u : 25u;
      const uint32_t idx = (uint32_t)((ty + 1) * kXW) + col;
      side_a = mapa(xbase + idx * 128u, (uint32_t)side_rank);
      side_b = mapa(rdy0, (uint32_t)side_rank);
      sw_side = idx & 7u;
      if (has_up && ty == 0) {
        const uint32_t di = (uint32_t)(17 * kXW) + col;
        dgu_a = mapa(xbase + di * 128u, (uint32_t)(up_rank + 1 - 2 * f));
        dgu_b = mapa(rdy0, (uint32_t)(up_rank + 1 - 2 * f));
        sw_dg = di & 7u;
      }
      if (has_dn && ty == 15) {
        const uint32_t di = col;
        dgd_a = mapa(xbase + di * 128u, (uint32_t)(dn_rank + 1 - 2 * f));
        dgd_b = mapa(rdy0, (uint32_t)(dn_rank + 1 - 2 * f));
        sw_dg = di & 7u;
      }
    }
    // global byte offset of this thread's 64 bytes for tile 0 (tile j: + j * 1024): RES_GLOBAL operands only
    const size_t goff0 = ((((size_t)n * p.H + band * 16 + ty) * p.W) + f * 24 + tx) * 128u + (size_t)h * 64u;
    // cooperative tile copy (shared -> global): 8 consecutive threads move one pixel's 128-byte line; thread = (16-byte
    // chunk cc, pixel column ctx, pixel rows cty + 4k for k = 0..3).  (4 * 26) % 8 == 0, so the swizzle is the same for all k.
    const int cc = et & 7, ctx = (et >> 3) & 7, cty = et >> 6;
    const uint32_t ce = ebase + (uint32_t)(et >> 3) * 128u + (((uint32_t)cc ^ (uint32_t)ctx) << 4);   // + k * 32 * 128
    const size_t cg0 = ((((size_t)n * p.H + band * 16 + cty) * p.W) + f * 24 + ctx) * 128u + (size_t)cc * 16u;   // + j * 1024 + k * 4 * W * 128
    const size_t cg_k = (size_t)4 * p.W * 128u;
    const uint32_t pool_bar0 = ptx::smem_u32(&S.pool_full[0]);
    uint32_t ca_count = 0;                          // two-phase (CALayer) ops so far
    uint32_t e_k = 0;                               // operand tiles consumed so far

    // write this pixel's 32 channels into X[ob] (centre + the neighbours' halos)
    auto store_x = [&](const uint32_t (&pk)[16], const int j, const int q, const int ob) {
      const uint32_t off = (uint32_t)j * 1024u + (uint32_t)ob * kXStride;
      const uint32_t boff = (uint32_t)ob * 16u + (q == 2 ? 8u : 0u);       // which ready barrier of the destination CTA
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 v = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        const uint32_t c = (uint32_t)(h * 4 + g);
        ptx::sts128(xl + off + ((c ^ sw_l) << 4), v);
        if (up_a) st_async_v4(up_a + off + ((c ^ sw_up) << 4), v, up_b + boff);
        if (dn_a) st_async_v4(dn_a + off + ((c ^ sw_dn) << 4), v, dn_b + boff);
        if (q == 0 && side_a) {
          const uint32_t xoff = (uint32_t)ob * kXStride;
          st_async_v4(side_a + xoff + ((c ^ sw_side) << 4), v, side_b + boff);
          if (dgu_a) st_async_v4(dgu_a + xoff + ((c ^ sw_dg) << 4), v, dgu_b + boff);
          if (dgd_a) st_async_v4(dgd_a + xoff + ((c ^ sw_dg) << 4), v, dgd_b + boff);
        }
      }
    };
    // centre only (X[ob] as scratch: dL/dout before the CALayer backward rewrites it in place)
    auto store_centre = [&](const uint32_t (&pk)[16], const int j, const int ob) {
      const uint32_t off = (uint32_t)j * 1024u + (uint32_t)ob * kXStride;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        ptx::sts128(xl + off + (((uint32_t)(h * 4 + g) ^ sw_l) << 4), make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]));
    };
    auto load_centre = [&](uint4 (&ev)[4], const int j, const int ob) {
      const uint32_t off = (uint32_t)j * 1024u + (uint32_t)ob * kXStride;
#pragma unroll
      for (int g = 0; g < 4; ++g) ev[g] = ptx::lds128(xl + off + (((uint32_t)(h * 4 + g) ^ sw_l) << 4));
    };
    // this thread's part of position q is in X[ob] (and the neighbours' halos): hand it to the publisher warp
    // (the publisher warp does the rest: release-arrive on the consumers' ready barriers)
    auto publish = [&](const int q, const int op) {
      fence_writer();
      named_bar_arrive(kBarTile + q, kPub);
      if (et == 0 && q == 0) CL_TRACE(op, 7);
      if (lane == 0 && q == 1) CL_TRACE(op, 16 + w8);
    };
    // Tile j of X[ob] -> global slot `ref`.  The two warps of a TMEM lane quarter hold the two channel halves of the same four
    // tile rows: once both have written them (64-thread barrier) lanes 0-3 of the first issue one 1 KB TMA row store each.
    // The rows start at 128-byte (not 1024-byte) aligned shared-memory addresses; the swizzle of a TMA store follows the
    // absolute address just as the UMMA operand fetch does (profiles/r02_hw_probes_p10.txt).  Callers have run fence_writer().
    auto store_rows = [&](const uint16_t ref, const int j, const int ob) {
      // these rows of X[ob] are rewritten two ops (six tiles) later, by the two warps that meet here: before the partner
      // may run on, the store issued five calls ago has finished reading shared memory
      if (h == 0 && lane < 4) ptx::bulk_wait_group_read<4>();
      ptx::named_bar_sync(kBarPair + quarter, 64);
      if (h == 0 && lane < 4) {
        const int trow = quarter * 4 + lane;
        const uint32_t src = xbase + (uint32_t)ob * kXStride + (uint32_t)((trow + 1) * kXW + 1 + 8 * j) * 128u;
        tma_store_5d(&maps.row[ref >> 14], src, 0, f * 24 + 8 * j, band * 16 + trow, n, ref & 0x3FFF);
        ptx::bulk_commit_group();
      }
    };
    // staging tile -> global slot (after a barrier that follows the writes): full 128-byte lines per 8 lanes
    auto copy_out_e = [&](uint8_t* slot, const int j) {
      uint8_t* g0 = slot + cg0 + (size_t)j * 1024u;
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ptx::lds128(ce + (uint32_t)k * 4096u);
#pragma unroll
      for (int k = 0; k < 4; ++k) stg128(g0 + (size_t)k * cg_k, v[k]);
    };
    // this pixel's 64 bytes of the operand tile the producer warp requested for (op, q)
    auto take_operand = [&](uint4 (&tv)[4]) {
      ptx::mbar_wait(&S.e_full, e_k & 1u);
#pragma unroll
      for (int g = 0; g < 4; ++g) tv[g] = ptx::lds128(el + (((uint32_t)(h * 4 + g) ^ (uint32_t)tx) << 4));
      // The buffer is handed back for the next TMA load only once the loaded VALUES have arrived in registers (an
      // instruction that consumes them cannot issue earlier): with the arrive issued right behind the loads, quarter-warp
      // passes of an LDS.128 delayed in a busy shared-memory pipe were overtaken by the next tile's TMA write (torn masks
      // at 8-lane granularity, profiles/r02_cluster_parity_v3_race.txt).
      uint32_t dep = tv[0].x ^ tv[1].y ^ tv[2].z ^ tv[3].w;
      asm volatile("mov.b32 %0, %0;" : "+r"(dep));
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&S.e_empty);
      ++e_k;
    };
    auto park_addr = [&](const int area, const int q) {
      return tmem_acc + tm_lane + kParkBase + (uint32_t)area * kParkArea + (uint32_t)q * 32u + (uint32_t)h * 16u;
    };
    // per-thread running sums -> per-channel totals of this CTA in S.wsum; returns after a barrier
    auto reduce_cta = [&](float (&rs)[32]) {
      const float tot = warp_transpose_reduce32(rs, lane);
      S.wsum[w8][lane] = tot;
      ptx::named_bar_sync(1, kEpi);
    };
    auto cta_total = [&](const int c) {      // c = et < 64
      const int hh = c >> 5, ch = c & 31;
      return (S.wsum[hh * 4][ch] + S.wsum[hh * 4 + 1][ch]) + (S.wsum[hh * 4 + 2][ch] + S.wsum[hh * 4 + 3][ch]);
    };
    // 64 per-CTA sums -> every CTA of the cluster gets all of them; the sample total lands in S.ca_tot
    auto all_gather = [&]() {
      const uint32_t par = ca_count & 1u, ph = (ca_count >> 1) & 1u;
      if (et < 64) {
        const float part = cta_total(et);
        const uint32_t slot = ptx::smem_u32(&S.pool[par][rank][et]);
        for (int r = 0; r < csize; ++r) st_async_b32(mapa(slot, (uint32_t)r), __float_as_uint(part), mapa(pool_bar0 + par * 8u, (uint32_t)r));
      }
      if (et == 64) ptx::mbar_arrive_expect_tx(&S.pool_full[par], (uint32_t)csize * 256u);     // 64 floats from every CTA of the cluster
      mbar_wait_cluster(&S.pool_full[par], ph);
      if (et < 64) {
        float t = 0.f;
        for (int r = 0; r < csize; ++r) t += S.pool[par][r][et];     // same order in every CTA: identical gates
        S.ca_tot[et] = t;
      }
      ++ca_count;
      ptx::named_bar_sync(1, kEpi);
    };

    for (int op = 0; op < p.n_ops; ++op) {
      const COp& o = p.ops[op];
      const uint32_t flags = o.flags;
      const int ob = (op + 1) & 1;
      const int bsel = op & 1;
      const int res_mode = o.res_mode;
      if (et < 64) S.bias[bsel][et] = o.bias ? __ldg(o.bias + et) : 0.f;
      ptx::named_bar_sync(1, kEpi);
      const float* bias_h = &S.bias[bsel][h * 32];
      const float scale = o.scale;
      float rs[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rs[i] = 0.f;

      // accumulator of position q -> v[32] = this pixel's 32 channels (raw fp32 sums)
      auto load_acc = [&](const int q, float (&v)[32], const bool release) {
        const uint32_t g = (uint32_t)(3 * op + q), slot = g & 3u;
        uint32_t acc[32];
        ptx::tmem_ld_32x32b_x32(tmem_acc + tm_lane + slot * 64u + (uint32_t)(h * 32), acc);
        ptx::tmem_ld_wait();
        if (release) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&S.acc_empty[slot]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
      };
      auto wait_acc = [&](const int q) {
        const uint32_t g = (uint32_t)(3 * op + q), slot = g & 3u;
        ptx::mbar_wait(&S.acc_full[slot], (g >> 2) & 1u);
        ptx::tc_fence_after();
      };
      auto add_bias = [&](float (&v)[32]) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bq = *reinterpret_cast<const float4*>(bias_h + i);
          v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
        }
      };
      // residual operand of position q (tile j) by the op's mode
      auto fetch_residual = [&](uint4 (&ev)[4], const int q, const int j) {
        if (res_mode == RES_INPLACE) {
          load_centre(ev, j, ob);
        } else if (res_mode == RES_TMEM) {
          uint32_t r16[16];
          tmem_ld_x16(park_addr(op & 1, q), r16);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) ev[g] = make_uint4(r16[g * 4], r16[g * 4 + 1], r16[g * 4 + 2], r16[g * 4 + 3]);
        } else {
          const uint8_t* src = o.e + goff0 + (size_t)j * 1024u;
#pragma unroll
          for (int g = 0; g < 4; ++g) ev[g] = ldg128(src + g * 16);
        }
      };
      auto park = [&](const uint32_t (&pk)[16], const int q) {
        tmem_st_x16(park_addr(op & 1, q), pk);
        tmem_st_wait();
      };

      if (!(flags & (SRB_CHAIN_CA | SRB_CHAIN_CA_BWD_FUSED))) {
        // ---------------- plain conv: bias, ReLU, scale, mask or residual (arithmetic = conv_chain.cu) ----------------
        auto tile = [&](auto FC, const int q) {
          constexpr uint32_t FK = decltype(FC)::value;
          const uint32_t F = (FK == kGeneric) ? ((flags & 15u) | kScaled) : FK;     // compile-time constant unless generic
          const int j = mirror ? 2 - q : q;
          uint4 ev[4];
          if (F & SRB_MASK) take_operand(ev);
          if (F & SRB_RESIDUAL) fetch_residual(ev, q, j);
          wait_acc(q);
          if (et == 0) CL_TRACE(op, q == 0 ? 4 : (q == 1 ? 9 : 11));
          float v[32];
          load_acc(q, v, true);
          if (et == 0 && q == 0) CL_TRACE(op, 5);
          add_bias(v);
          if (F & SRB_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if ((F & kScaled) && scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= scale;
          }
          if (F & SRB_MASK) {
            float fm[32];
            unpack4(ev, fm);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!(fm[i] > 0.f)) v[i] = 0.f;
          }
          if (F & SRB_RESIDUAL) {
            float fr[32];
            unpack4(ev, fr);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += fr[i];
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          if (o.park) park(pk, q);
          store_x(pk, j, q, ob);
          if (F & SRB_COLSUM) {          // sums of the STORED (bf16-rounded) values, as the other conv kernels
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 fr = unpack_bf16x2(pk[i]);
              rs[2 * i] += fr.x;
              rs[2 * i + 1] += fr.y;
            }
          }
          if (et == 0 && q == 0) CL_TRACE(op, 6);
          publish(q, op);
          store_rows(o.y_ref, j, ob);
        };
        auto run = [&](auto FC) {
#pragma unroll 1
          for (int q = 0; q < 3; ++q) tile(FC, q);
        };
        switch ((flags & 15u) | (scale != 1.f ? kScaled : 0u)) {
          case SRB_RELU: run(std::integral_constant<uint32_t, SRB_RELU>{}); break;
          case SRB_RESIDUAL: run(std::integral_constant<uint32_t, SRB_RESIDUAL>{}); break;
          case SRB_RESIDUAL | kScaled: run(std::integral_constant<uint32_t, SRB_RESIDUAL | kScaled>{}); break;
          case SRB_RESIDUAL | SRB_COLSUM: run(std::integral_constant<uint32_t, SRB_RESIDUAL | SRB_COLSUM>{}); break;
          case SRB_MASK | SRB_COLSUM: run(std::integral_constant<uint32_t, SRB_MASK | SRB_COLSUM>{}); break;
          case SRB_MASK | SRB_COLSUM | kScaled: run(std::integral_constant<uint32_t, SRB_MASK | SRB_COLSUM | kScaled>{}); break;
          case SRB_COLSUM: run(std::integral_constant<uint32_t, SRB_COLSUM>{}); break;
          case 0: run(std::integral_constant<uint32_t, 0>{}); break;
          default: run(std::integral_constant<uint32_t, kGeneric>{}); break;
        }
        if (flags & SRB_COLSUM) {        // bias gradients etc.: after the last publish, off the chain's critical path
          reduce_cta(rs);
          if (et < 64) {
            const float tot = cta_total(et);
            atomicAdd(o.colsum + (size_t)(o.colsum_groups > 1 ? n : 0) * 64 + et, o.colsum_scale != 0.f ? tot * o.colsum_scale : tot);
          }
        }
      } else if (flags & SRB_CHAIN_CA) {
        // ---------------- RCAB conv2 + CALayer + skip (rcan.py:10-29,54) ----------------
        const int Cr = o.ca_cr;
        const bool fast = Cr <= 4;
        // pass 1: t = conv + bias (bf16) -> staging tile -> global; pooled sums of the stored values; accumulators stay in TMEM
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          wait_acc(q);
          if (et == 0) CL_TRACE(op, q == 0 ? 4 : (q == 1 ? 9 : 11));
          float v[32];
          load_acc(q, v, false);
          add_bias(v);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              pk[e] = pack_bf16x2(v[g * 8 + e * 2], v[g * 8 + e * 2 + 1]);
              const float2 fr = unpack_bf16x2(pk[e]);
              rs[g * 8 + e * 2] += fr.x;
              rs[g * 8 + e * 2 + 1] += fr.y;
            }
            ptx::sts128(el + (((uint32_t)(h * 4 + g) ^ (uint32_t)tx) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          }
          ptx::named_bar_sync(1, kEpi);
          copy_out_e(o.y, j);
          ptx::named_bar_sync(1, kEpi);          // staging tile free for the next position
        }
        // gate operands do not depend on the pool: fetch them while the sums travel
        float w1a = 0.f, w1b = 0.f, b1v = 0.f;
        if (fast && w8 < Cr) {
          w1a = __ldg(o.ca_w1 + w8 * 64 + lane);
          w1b = __ldg(o.ca_w1 + w8 * 64 + lane + 32);
          b1v = __ldg(o.ca_b1 + w8);
        }
        float w2r[4] = {0.f, 0.f, 0.f, 0.f}, b2v = 0.f;
        if (et < 64) {
          b2v = __ldg(o.ca_b2 + et);
          if (fast)
            for (int jj = 0; jj < Cr; ++jj) w2r[jj] = __ldg(o.ca_w2 + et * Cr + jj);
        }
        reduce_cta(rs);
        all_gather();
        if (et == 0) CL_TRACE(op, 14);
        if (et < 64) {
          const float tot = S.ca_tot[et];
          S.ca_s[et] = tot * inv_hw;
          if (rank == 0) {
            o.colsum[(size_t)n * 64 + et] = tot;
            o.ca_s[(size_t)n * 64 + et] = tot * inv_hw;
          }
        }
        ptx::named_bar_sync(1, kEpi);
        if (fast) {
          if (w8 < Cr) {
            const float a = warp_sum(w1a * S.ca_s[lane] + w1b * S.ca_s[lane + 32]);
            if (lane == 0) S.ca_z[w8] = fmaxf(a + b1v, 0.f);
          }
        } else {
          for (int jj = w8; jj < Cr; jj += 8) {
            const float a = warp_sum(__ldg(o.ca_w1 + jj * 64 + lane) * S.ca_s[lane] + __ldg(o.ca_w1 + jj * 64 + lane + 32) * S.ca_s[lane + 32]);
            if (lane == 0) S.ca_z[jj] = fmaxf(a + __ldg(o.ca_b1 + jj), 0.f);
          }
        }
        ptx::named_bar_sync(1, kEpi);
        if (et < 64) {
          float u = b2v;
          for (int jj = 0; jj < Cr; ++jj) u += (fast ? w2r[jj & 3] : __ldg(o.ca_w2 + et * Cr + jj)) * S.ca_z[jj];
          const float yv = 1.f / (1.f + expf(-u));
          S.ca_y[et] = yv;
          if (rank == 0) o.ca_y[(size_t)n * 64 + et] = yv;
        }
        ptx::named_bar_sync(1, kEpi);
        // pass 2: out = t * gate + skip -> X[ob] (+ the neighbours' halos) -> global
        const float* yh = &S.ca_y[h * 32];
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          uint4 ev[4];
          fetch_residual(ev, q, j);
          float v[32];
          load_acc(q, v, true);
          add_bias(v);
          uint32_t pk[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t xw[4] = {ev[g].x, ev[g].y, ev[g].z, ev[g].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 ft = unpack_bf16x2(pack_bf16x2(v[g * 8 + e * 2], v[g * 8 + e * 2 + 1]));   // t as stored
              const float2 fx = unpack_bf16x2(xw[e]);
              const int ch = g * 8 + e * 2;
              pk[g * 4 + e] = pack_bf16x2(fmaf(ft.x, yh[ch], fx.x), fmaf(ft.y, yh[ch + 1], fx.y));
            }
          }
          if (o.park) park(pk, q);
          store_x(pk, j, q, ob);
          publish(q, op);
          store_rows(o.y2_ref, j, ob);
        }
      } else {
        // ---------------- dgrad conv (+ residual) fused with the CALayer backward of the block whose dL/dout it produces --------
        const int Cr = o.ca_cr;
        const bool fast = Cr <= 4;
        const bool has_res = (flags & SRB_RESIDUAL) != 0;
        // pass 1: g = conv (* scale) + residual (bf16) -> X[ob] centre (scratch) -> global; running sums of g * t
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          uint4 ev[4], tv[4];
          take_operand(tv);
          if (has_res) fetch_residual(ev, q, j);
          wait_acc(q);
          if (et == 0) CL_TRACE(op, q == 0 ? 4 : (q == 1 ? 9 : 11));
          float v[32];
          load_acc(q, v, true);
          add_bias(v);
          if (scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= scale;
          }
          if (has_res) {
            float fr[32];
            unpack4(ev, fr);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += fr[i];
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          {
            float ft[32];
            unpack4(tv, ft);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 fg = unpack_bf16x2(pk[i]);
              rs[2 * i] = fmaf(fg.x, ft[2 * i], rs[2 * i]);
              rs[2 * i + 1] = fmaf(fg.y, ft[2 * i + 1], rs[2 * i + 1]);
            }
          }
          if (o.park) park(pk, q);
          store_centre(pk, j, ob);
          fence_writer();
          store_rows(o.y_ref, j, ob);
        }
        // everything the gate backward needs that does not depend on the sample sums
        float w1a = 0.f, w1b = 0.f, b1v = 0.f, w2a = 0.f, w2b = 0.f;
        if (fast && w8 < Cr) {
          w1a = __ldg(o.ca_w1 + w8 * 64 + lane);
          w1b = __ldg(o.ca_w1 + w8 * 64 + lane + 32);
          b1v = __ldg(o.ca_b1 + w8);
          w2a = __ldg(o.ca_w2 + lane * Cr + w8);
          w2b = __ldg(o.ca_w2 + (lane + 32) * Cr + w8);
        }
        float w2r[4] = {0.f, 0.f, 0.f, 0.f}, w1c[4] = {0.f, 0.f, 0.f, 0.f}, b2v = 0.f;
        if (et < 64) {
          b2v = __ldg(o.ca_b2 + et);
          if (fast)
            for (int jj = 0; jj < Cr; ++jj) {
              w2r[jj] = __ldg(o.ca_w2 + et * Cr + jj);
              w1c[jj] = __ldg(o.ca_w1 + jj * 64 + et);
            }
          S.ca_s[et] = o.ca_s[(size_t)n * 64 + et];      // written by the forward launch
          S.ca_y[et] = o.ca_y[(size_t)n * 64 + et];
        }
        reduce_cta(rs);                                   // (barrier: ca_s / ca_y visible)
        if (fast) {
          if (w8 < Cr) {
            float a = warp_sum(w1a * S.ca_s[lane] + w1b * S.ca_s[lane + 32]);
            if (lane == 0) {
              a += b1v;
              S.ca_z[w8] = fmaxf(a, 0.f);
              S.ca_dv[w8] = a > 0.f ? 1.f : 0.f;
            }
          }
        } else {
          for (int jj = w8; jj < Cr; jj += 8) {
            float a = warp_sum(__ldg(o.ca_w1 + jj * 64 + lane) * S.ca_s[lane] + __ldg(o.ca_w1 + jj * 64 + lane + 32) * S.ca_s[lane + 32]);
            if (lane == 0) {
              a += __ldg(o.ca_b1 + jj);
              S.ca_z[jj] = fmaxf(a, 0.f);
              S.ca_dv[jj] = a > 0.f ? 1.f : 0.f;
            }
          }
        }
        all_gather();                                     // (barriers: ca_z / ca_dv visible, S.ca_tot = sum g*t over the sample)
        if (et == 0) CL_TRACE(op, 14);
        if (et < 64) {
          float u = b2v;
          for (int jj = 0; jj < Cr; ++jj) u += (fast ? w2r[jj & 3] : __ldg(o.ca_w2 + et * Cr + jj)) * S.ca_z[jj];
          const float sp = 1.f / (1.f + expf(-u)), sn = 1.f / (1.f + expf(u));
          S.ca_du[et] = S.ca_tot[et] * (sp * sn);         // sigmoid'(u) from u itself
        }
        ptx::named_bar_sync(1, kEpi);
        if (fast) {
          if (w8 < Cr) {
            const float dz = warp_sum(w2a * S.ca_du[lane] + w2b * S.ca_du[lane + 32]);
            if (lane == 0) S.ca_dv[w8] *= dz;
          }
        } else {
          for (int jj = w8; jj < Cr; jj += 8) {
            const float dz = warp_sum(__ldg(o.ca_w2 + lane * Cr + jj) * S.ca_du[lane] + __ldg(o.ca_w2 + (lane + 32) * Cr + jj) * S.ca_du[lane + 32]);
            if (lane == 0) S.ca_dv[jj] *= dz;
          }
        }
        ptx::named_bar_sync(1, kEpi);
        if (et < 64) {
          float d = 0.f;
          for (int jj = 0; jj < Cr; ++jj) d += (fast ? w1c[jj & 3] : __ldg(o.ca_w1 + jj * 64 + et)) * S.ca_dv[jj];
          S.ca_ds[et] = d * inv_hw;
        }
        if (rank == 0) {                                  // parameter gradients, once per sample
          for (int i = et; i < 64 * Cr; i += kEpi) {
            atomicAdd(o.ca_dw2 + i, S.ca_du[i / Cr] * S.ca_z[i % Cr]);     // w2 [64][Cr]
            atomicAdd(o.ca_dw1 + i, S.ca_dv[i / 64] * S.ca_s[i % 64]);     // w1 [Cr][64]
          }
          if (et < 64) atomicAdd(o.ca_db2 + et, S.ca_du[et]);
          if (et < Cr) atomicAdd(o.ca_db1 + et, S.ca_dv[et]);
        }
        if (h == 0 && lane < 4) ptx::bulk_wait_group_read<0>();      // dL/dout has left X[ob]: it may be rewritten in place
        ptx::named_bar_sync(1, kEpi);
        // pass 2: dt = g * gate + ds / HW, in place in X[ob] (+ halos) -> global; column sums of dt
        const float* yh = &S.ca_y[h * 32];
        const float* dsh = &S.ca_ds[h * 32];
#pragma unroll
        for (int i = 0; i < 32; ++i) rs[i] = 0.f;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          const int j = mirror ? 2 - q : q;
          uint4 gv[4];
          load_centre(gv, j, ob);
          uint32_t pk[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t gw[4] = {gv[g].x, gv[g].y, gv[g].z, gv[g].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fg = unpack_bf16x2(gw[e]);
              const int ch = g * 8 + e * 2;
              pk[g * 4 + e] = pack_bf16x2(fmaf(fg.x, yh[ch], dsh[ch]), fmaf(fg.y, yh[ch + 1], dsh[ch + 1]));
              const float2 fd = unpack_bf16x2(pk[g * 4 + e]);
              rs[ch] += fd.x;
              rs[ch + 1] += fd.y;
            }
          }
          store_x(pk, j, q, ob);
          publish(q, op);
          store_rows(o.y2_ref, j, ob);
        }
        if (o.colsum2) {
          reduce_cta(rs);
          if (et < 64) atomicAdd(o.colsum2 + et, cta_total(et));
        }
      }
      if (et == 0) CL_TRACE(op, 13);
    }
    if (h == 0 && lane < 4) ptx::bulk_wait_group<0>();      // all row stores complete before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA exits while a peer may still store into its shared memory or arrive on its barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool cluster_enabled() {
  const char* e = getenv("SRB200_CHAIN_CLUSTER");
  return !(e && e[0] == '0');
}

int encode_5d(srb_ctx* ctx, CUtensorMap* map, void* ptr, int slots, int N, int H, int W, int box_w, int box_h) {
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)slots};
  cuuint64_t strides[4] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128, (cuuint64_t)N * H * W * 128};
  cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    srb_set_error("srb_conv_chain: cuTensorMapEncodeTiled failed with CUresult %d (slots=%d N=%d H=%d W=%d box %dx%d)", (int)r, slots, N, H,
                  W, box_w, box_h);
    return 4;
  }
  return 0;
}

}  // namespace

// The cluster form takes a chain when: H % 16 == 0, W in {24, 48}, (H/16) * (W/24) <= 8 CTAs per sample; every op is a
// conv whose input is the previous op's result (y, or y2 for the two-phase CALayer ops); no standalone CA_BWD tile op;
// CALayer-forward ops (which use the operand-tile buffer as staging) and TMA operand tiles (MASK / CA_BWD_FUSED) do not
// occur in the same chain.  Returns 1 if it can run the chain, 0 if not (srb_conv_chain then uses the L2-flag kernel).
int srb_chain_cluster_eligible(const srb_chain_desc* d) {
  if (!cluster_enabled() || d->kernel_hint == 1) return 0;
  if (d->H % 16 != 0 || (d->W != 24 && d->W != 48) || d->H < 16) return 0;
  const int bands = d->H / 16, halves = d->W / 24;
  if (bands * halves > kMaxCluster) return 0;
  uint16_t prev = SRB_CHAIN_NONE;
  bool any_ca = false, any_tile = false;
  for (int i = 0; i < d->n_ops; ++i) {
    const srb_chain_op& o = d->ops[i];
    if (o.kind != SRB_CHAIN_CONV) return 0;
    if (i > 0 && o.x != prev) return 0;
    const uint32_t fl = o.flags;
    if (fl & ~(uint32_t)(SRB_RELU | SRB_RESIDUAL | SRB_MASK | SRB_COLSUM | SRB_CHAIN_CA | SRB_CHAIN_CA_BWD_FUSED | SRB_CHAIN_Y_SCRATCH)) return 0;
    if ((fl & SRB_MASK) && (fl & SRB_RESIDUAL)) return 0;
    if (fl & SRB_MASK) any_tile = true;
    if (fl & SRB_CHAIN_CA) {
      if ((fl & (SRB_RELU | SRB_MASK | SRB_CHAIN_CA_BWD_FUSED)) || !(fl & SRB_RESIDUAL) || o.scale != 1.f) return 0;
      if (o.ca_cr < 1 || o.ca_cr > kMaxCr) return 0;
      any_ca = true;
      prev = o.y2;
    } else if (fl & SRB_CHAIN_CA_BWD_FUSED) {
      if (fl & (SRB_RELU | SRB_MASK | SRB_COLSUM)) return 0;
      if (o.ca_cr < 1 || o.ca_cr > kMaxCr) return 0;
      any_tile = true;
      prev = o.y2;
    } else {
      prev = o.y;
    }
  }
  if (any_ca && any_tile) return 0;
  return 1;
}

int srb_chain_cluster_launch(srb_ctx* ctx, const srb_chain_desc* d, void* stream) {
  static_assert(sizeof(CParams) + sizeof(CMaps) < 32000, "kernel parameters exceed the 32 KB limit");
  static_assert(kSmemBytes <= 227u * 1024u, "shared memory budget");
  static_assert(kParkBase + 2 * kParkArea <= kTmemCols, "TMEM budget");
  CParams* pp = new CParams();
  CMaps* mm = new CMaps();
  struct Guard {
    CParams* a;
    CMaps* b;
    ~Guard() {
      delete a;
      delete b;
    }
  } guard{pp, mm};
  CParams& p = *pp;
  p.N = d->N;
  p.H = d->H;
  p.W = d->W;
  p.bands = d->H / 16;
  p.halves = d->W / 24;
  p.n_ops = d->n_ops;
  p.trace = reinterpret_cast<long long*>(d->trace);
  const size_t slot_bytes = (size_t)d->N * d->H * d->W * 128u;
  bool used[4] = {false, false, false, false}, used_row[4] = {false, false, false, false};
  auto slot_ptr = [&](uint16_t r, const char* what, int op, uint8_t** out) -> int {
    *out = nullptr;
    if (r == SRB_CHAIN_NONE) return 0;
    const int sp = r >> 14, slot = r & 0x3FFF;
    SRB_REQUIRE(d->space_base[sp] != nullptr && slot < d->space_slots[sp],
                "srb_conv_chain: op %d %s reference (space %d, slot %d) outside the declared spaces", op, what, sp, slot);
    *out = static_cast<uint8_t*>(d->space_base[sp]) + (size_t)slot * slot_bytes;
    return 0;
  };
  for (int i = 0; i < d->n_ops; ++i) {
    const srb_chain_op& o = d->ops[i];
    COp& c = p.ops[i];
    c.flags = o.flags;
    c.w_layer = o.w_layer;
    SRB_REQUIRE(o.w_layer >= 0 && o.w_layer < d->n_layers, "srb_conv_chain: op %d filter index %d outside [0,%d)", i, o.w_layer, d->n_layers);
    c.scale = o.scale;
    c.colsum_scale = o.colsum_scale;
    c.colsum_groups = o.colsum_groups;
    c.ca_cr = o.ca_cr;
    c.bias = o.bias;
    c.colsum = o.colsum;
    c.colsum2 = o.colsum2;
    c.ca_w1 = o.ca_w1; c.ca_b1 = o.ca_b1; c.ca_w2 = o.ca_w2; c.ca_b2 = o.ca_b2;
    c.ca_s = o.ca_s; c.ca_y = o.ca_y;
    c.ca_dw1 = o.ca_dw1; c.ca_db1 = o.ca_db1; c.ca_dw2 = o.ca_dw2; c.ca_db2 = o.ca_db2;
    int rc;
    uint8_t *y, *y2, *e, *e2;
    if ((rc = slot_ptr(o.y, "y", i, &y))) return rc;
    if ((rc = slot_ptr(o.y2, "y2", i, &y2))) return rc;
    if ((rc = slot_ptr(o.e, "mask/residual", i, &e))) return rc;
    if ((rc = slot_ptr(o.e2, "saved t", i, &e2))) return rc;
    c.y = y; c.y2 = y2; c.e = e;
    c.park = 0;
    c.y_ref = o.y;
    c.y2_ref = o.y2;
    used_row[o.y >> 14] = true;
    if (o.y2 != SRB_CHAIN_NONE) used_row[o.y2 >> 14] = true;
    SRB_REQUIRE(y != nullptr, "srb_conv_chain: op %d needs a y buffer", i);
    const bool m = (o.flags & SRB_MASK) != 0, r = (o.flags & SRB_RESIDUAL) != 0;
    SRB_REQUIRE(!(m || r) || e != nullptr, "srb_conv_chain: op %d needs a mask/residual buffer", i);
    SRB_REQUIRE((m || r) || o.e == SRB_CHAIN_NONE, "srb_conv_chain: op %d has an operand tile but no MASK/RESIDUAL flag", i);
    SRB_REQUIRE(!(o.flags & SRB_COLSUM) || (o.colsum && (o.colsum_groups == 1 || o.colsum_groups == d->N)),
                "srb_conv_chain: op %d: COLSUM needs a pointer and groups in {1, N}", i);
    SRB_REQUIRE((o.e2 != SRB_CHAIN_NONE) == ((o.flags & SRB_CHAIN_CA_BWD_FUSED) != 0),
                "srb_conv_chain: op %d: a second operand tile goes with CA_BWD_FUSED and only with it", i);
    if (o.flags & SRB_CHAIN_CA) {
      SRB_REQUIRE((o.flags & SRB_COLSUM) && o.colsum_groups == d->N, "srb_conv_chain: op %d: CA needs COLSUM per sample", i);
      SRB_REQUIRE(o.ca_w1 && o.ca_b1 && o.ca_w2 && o.ca_b2 && o.ca_s && o.ca_y && y2, "srb_conv_chain: op %d: CA parameters missing", i);
    }
    if (o.flags & SRB_CHAIN_CA_BWD_FUSED) {
      SRB_REQUIRE(o.ca_w1 && o.ca_b1 && o.ca_w2 && o.ca_b2 && o.ca_s && o.ca_y && o.ca_dw1 && o.ca_db1 && o.ca_dw2 && o.ca_db2 && y2 && e2,
                  "srb_conv_chain: op %d: CA_BWD_FUSED pointers missing", i);
    }
    // operand tile by TMA
    c.t_ref = m ? o.e : ((o.flags & SRB_CHAIN_CA_BWD_FUSED) ? o.e2 : (uint16_t)SRB_CHAIN_NONE);
    if (c.t_ref != SRB_CHAIN_NONE) used[c.t_ref >> 14] = true;
    // where the residual comes from: still in X[out] (it was the previous op's input), parked in TMEM by the op before
    // the previous one, or global memory (read by the thread that owns the pixel)
    c.res_mode = RES_NONE;
    if (r) {
      if (i >= 1 && o.e == d->ops[i - 1].x) {
        c.res_mode = RES_INPLACE;
      } else if (i >= 2 && o.e == d->ops[i - 2].y && !(d->ops[i - 2].flags & SRB_CHAIN_CA)) {
        c.res_mode = RES_TMEM;
        p.ops[i - 2].park = 1;
      } else {
        c.res_mode = RES_GLOBAL;
      }
    }
  }
  SRB_REQUIRE(d->weights && d->n_layers > 0, "srb_conv_chain: conv ops need a filter bank");
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  {
    const uint16_t r = d->ops[0].x;
    SRB_REQUIRE(r != SRB_CHAIN_NONE, "srb_conv_chain: op 0 needs an x buffer");
    const int sp = r >> 14, slot = r & 0x3FFF;
    SRB_REQUIRE(d->space_base[sp] != nullptr && slot < d->space_slots[sp], "srb_conv_chain: op 0 x reference outside the declared spaces");
    SRB_REQUIRE(((uintptr_t)d->space_base[sp] & 127) == 0, "srb_conv_chain: space %d must be 128-byte aligned", sp);
    p.x0_slot = slot;
    int rc = encode_5d(ctx, &mm->x0, d->space_base[sp], d->space_slots[sp], d->N, d->H, d->W, kXW, kXH);
    if (rc) return rc;
  }
  for (int s = 0; s < 4; ++s) {
    mm->tile[s] = mm->x0;        // unused entries are never dereferenced; this keeps the parameter block initialised
    mm->row[s] = mm->x0;
    if (!used[s] && !used_row[s]) continue;
    SRB_REQUIRE(((uintptr_t)d->space_base[s] & 127) == 0, "srb_conv_chain: space %d must be 128-byte aligned", s);
    int rc = 0;
    if (used[s]) rc = encode_5d(ctx, &mm->tile[s], d->space_base[s], d->space_slots[s], d->N, d->H, d->W, 8, 16);
    if (rc) return rc;
    if (used_row[s]) rc = encode_5d(ctx, &mm->row[s], d->space_base[s], d->space_slots[s], d->N, d->H, d->W, 8, 1);
    if (rc) return rc;
  }
  {
    SRB_REQUIRE(((uintptr_t)d->weights & 127) == 0, "srb_conv_chain: filter bank must be 128-byte aligned");
    cuuint64_t dims[4] = {64, 64, 9, (cuuint64_t)d->n_layers};
    cuuint64_t strides[3] = {128, 64 * 128, 9 * 64 * 128};
    cuuint32_t box[4] = {64, 64, 3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = fn(&mm->w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->weights), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(cr == CUDA_SUCCESS, "srb_conv_chain: cuTensorMapEncodeTiled(filters) failed with CUresult %d", (int)cr);
  }
  SRB_REQUIRE((int)kSmemBytes <= ctx->smem_optin, "srb_conv_chain: needs %u bytes of shared memory, device offers %d", kSmemBytes,
              ctx->smem_optin);
  static bool attr_set = false;
  if (!attr_set) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(chain_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  const int csize = p.bands * p.halves;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(d->N * csize));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SRB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, chain_cluster_kernel, *mm, p));
  SRB_LAUNCH_CHECK();
  return 0;
}
