// 3x3 convolution, <=64 -> 64 channels, bf16 NHWC: the layer shape that carries 95 % of RCAN's and
// 61 % of EDSR-baseline's FLOPs.  Persistent tcgen05 kernel, one CTA per SM.
//
// What the hardware probes (profiles/r01_hw_probes.txt) say about this shape and how the kernel
// answers:
//  * A 128 x 64 x 16 SS-mode MMA is shared-memory-bound (6 KB of operands per 32-cycle MMA): 48 cycles,
//    66.6 % of the tensor peak.  Nothing to win on the MMA itself without cta_group::2.
//  * A TMA box takes ~2300 cycles to land from L2 whatever its size and an SM pulls ~65 B/cycle, so
//    the bytes an SM must import per layer are the budget.  UMMA applies the 128-byte swizzle to
//    ABSOLUTE shared-memory address bits, so a K-major operand may start at any 128-byte row of a
//    TMA-written buffer: ONE (16+2) x (8+2)-pixel input window (23 KB) serves all nine filter taps
//    (tap (kh,kw) = descriptor start + (kh*10+kw) rows, stride-byte-offset = one window row).
//    Per tile that is 23 KB instead of three kw-shifted 18 KB windows.
//  * The filter bank (72 KB) is loaded once per CTA and stays resident.  It does not depend on the
//    previous kernel, so it is requested BEFORE griddepcontrol.wait: with programmatic dependent
//    launch the request overlaps the previous layer's tail.
//  * Thread-per-pixel global accesses in the epilogue (32 different 128-byte lines per warp
//    instruction) cost microseconds on a 5 us kernel.  Residual / ReLU-mask tiles are therefore
//    prefetched by TMA into shared memory while the MMAs run, results are staged in shared memory
//    (128-byte swizzle, conflict-free 16-byte accesses) and written by TMA stores, which also clip
//    partial tiles.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 =
// epilogue.  Two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kTW = 8, kTH = 16;                 // output tile: 8 wide x 16 tall = 128 pixels = UMMA M
constexpr int kP = kTW + 2, kRows = kTH + 2;     // input window incl. the 1-pixel halo
constexpr uint32_t kWinBytes = kRows * kP * 128; // 23040
constexpr uint32_t kWinStride = 23u * 1024u;     // ring stride (1024-byte aligned)
constexpr uint32_t kWBytes = 9u * 64u * 128u;    // 73728: [kw][kh][cout][cin] bf16
constexpr uint32_t kTileBytes = 128u * 128u;     // 128 pixels x 64 channels bf16
constexpr int kMaxAStages = 4;
constexpr uint32_t kTmemCols = 128;              // two 64-column fp32 accumulators

struct C64Params {
  srb_conv_desc d;
  const float* bias;
  float* colsum;
  int tiles_w, tiles_h, total_tiles;
  int a_stages;                 // input-window ring depth
  int n_e;                      // epilogue operand tiles per output tile: [mask][residual]
  uint32_t off_a, off_e, off_o; // offsets from the 1024-aligned shared-memory base
  long long* trace;             // diagnostics: per-CTA event clocks (srb_debug_set_trace), else NULL
};

// event slots of the per-CTA trace record (16 x int64 per CTA)
enum { EV_T0_NS = 0, EV_PROLOGUE, EV_DEPWAIT, EV_WFULL, EV_A0FULL, EV_MMA0_ISSUED, EV_ACC0, EV_EPI0, EV_ACC1, EV_EPI1,
       EV_STORED, EV_END, EV_T1_NS, EV_SMID };
#define TRACE(ev)                                                                   \
  do {                                                                              \
    if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (ev)] = clock64() - t_start;     \
  } while (0)

__device__ __forceinline__ void tile_coords(const C64Params& p, int t, int& n, int& h0, int& w0) {
  const int tw_i = t % p.tiles_w;
  const int th_i = (t / p.tiles_w) % p.tiles_h;
  n = t / (p.tiles_w * p.tiles_h);
  h0 = th_i * kTH;
  w0 = tw_i * kTW;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_c64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY2,
                const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmR, const C64Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full;
  __shared__ uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ uint64_t e_full[2], e_empty[2];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  // the next kernel of the stream may start its own prologue now (it cannot touch our outputs
  // before its griddepcontrol.wait, which waits for this whole grid)
  ptx::griddep_launch_dependents();
  const long long t_start = clock64();
  if (p.trace && threadIdx.x == 0) {
    p.trace[(size_t)blockIdx.x * 16 + EV_T0_NS] = (long long)ptx::globaltimer_ns();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[(size_t)blockIdx.x * 16 + EV_SMID] = smid;
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const srb_conv_desc& d = p.d;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = base;
  const uint32_t abase = base + p.off_a;
  const uint32_t ebase = base + p.off_e;
  const uint32_t obase = base + p.off_o;
  const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool has_mask = (d.flags & SRB_MASK) != 0, has_res = (d.flags & SRB_RESIDUAL) != 0;

  if (threadIdx.x == 0) {
    ptx::mbar_init(&w_full, 1);
    for (int s = 0; s < kMaxAStages; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&e_full[b], 1);
      ptx::mbar_init(&e_empty[b], 128);
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;
  if (threadIdx.x == 0) TRACE(EV_PROLOGUE);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && my_tiles > 0) {
      ptx::prefetch_tensormap(&tmW);
      ptx::prefetch_tensormap(&tmX);
      ptx::mbar_arrive_expect_tx(&w_full, kWBytes);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) ptx::tma_load_3d(wbase + kw * 3u * 64u * 128u, &tmW, &w_full, 0, 0, kw * 3);
      if (p.n_e) {
        if (has_mask) ptx::prefetch_tensormap(&tmM);
        if (has_res) ptx::prefetch_tensormap(&tmR);
      }
      ptx::griddep_wait();   // activations below are written by the previous kernel(s)
      TRACE(EV_DEPWAIT);
      for (int ti = 0; ti < my_tiles; ++ti) {
        int n, h0, w0;
        tile_coords(p, (int)blockIdx.x + ti * (int)gridDim.x, n, h0, w0);
        const int s = ti % p.a_stages;
        ptx::mbar_wait(&a_empty[s], (((uint32_t)(ti / p.a_stages)) & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&a_full[s], kWinBytes);
        ptx::tma_load_4d(abase + (uint32_t)s * kWinStride, &tmX, &a_full[s], d.x_co, w0 - 1, h0 - 1, n);
        if (p.n_e) {
          const int es = ti & 1;
          ptx::mbar_wait(&e_empty[es], (((uint32_t)(ti >> 1)) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&e_full[es], (uint32_t)p.n_e * kTileBytes);
          uint32_t dst = ebase + (uint32_t)es * 2u * kTileBytes;
          if (has_mask) {
            ptx::tma_load_4d(dst, &tmM, &e_full[es], d.m_co, w0, h0, n);
            dst += kTileBytes;
          }
          if (has_res) ptx::tma_load_4d(dst, &tmR, &e_full[es], d.r_co, w0, h0, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (my_tiles > 0 && ptx::elect_one_sync()) {   // elect.sync: operands stay in uniform registers
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 0, 0);
      constexpr uint32_t hi_a = ptx::smem_desc_hi_sw128((uint32_t)kP * 128u);   // 8-pixel row groups, one window row apart
      constexpr uint32_t hi_b = ptx::smem_desc_hi_sw128(1024u);
      ptx::mbar_wait(&w_full, 0);
      TRACE(EV_WFULL);
      const uint32_t w_lo = ptx::smem_desc_lo(wbase, 16u);
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = ti & 1;
        ptx::mbar_wait(&acc_empty[buf], (((uint32_t)(ti >> 1)) & 1u) ^ 1u);
        const int s = ti % p.a_stages;
        ptx::mbar_wait(&a_full[s], ((uint32_t)(ti / p.a_stages)) & 1u);
        ptx::tc_fence_after();
        if (ti == 0) TRACE(EV_A0FULL);
        const uint32_t tmem_d = tmem_acc + (uint32_t)(buf * 64);
        const uint32_t a_lo = ptx::smem_desc_lo(abase + (uint32_t)s * kWinStride, 16u);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(tmem_d, a_lo + (uint32_t)((kh * kP + kw) * 8 + k * 2), hi_a,
                                  w_lo + (uint32_t)((kw * 3 + kh) * 512 + k * 2), hi_b, idesc, (kw | kh | k) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&a_empty[s]);
        ptx::umma_commit(&acc_full[buf]);
        if (ti == 0) TRACE(EV_MMA0_ISSUED);
      }
    }
  } else {
    // ===================== epilogue (128 threads, thread = one pixel of the tile) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;          // pixel index inside the tile: (row / 8, row % 8)
    const uint32_t sw = (uint32_t)(row & 7);
    const bool store_thread = (threadIdx.x == 64);
    ptx::griddep_wait();
    if (store_thread) {
      ptx::prefetch_tensormap(&tmY);
      if (d.flags & SRB_OUT2) ptx::prefetch_tensormap(&tmY2);
    }
    for (int ti = 0; ti < my_tiles; ++ti) {
      int n, h0, w0;
      tile_coords(p, (int)blockIdx.x + ti * (int)gridDim.x, n, h0, w0);
      const int buf = ti & 1;
      const uint32_t par = ((uint32_t)(ti >> 1)) & 1u;
      const bool valid = (h0 + (row >> 3) < d.H) && (w0 + (row & 7) < d.W);
      const uint32_t orow = obase + (uint32_t)buf * kTileBytes + (uint32_t)row * 128u;
      const uint32_t erow = ebase + (uint32_t)buf * 2u * kTileBytes + (uint32_t)row * 128u;
      if (ti >= 2) {
        // the TMA store of tile ti-2 must have finished reading this staging buffer
        if (store_thread) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(1, 128);
      }
      ptx::mbar_wait(&acc_full[buf], par);
      ptx::tc_fence_after();
      if (store_thread && ti < 2) TRACE(ti == 0 ? EV_ACC0 : EV_ACC1);
      if (p.n_e) ptx::mbar_wait(&e_full[buf], par);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t acc[32];
        ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + c0), acc);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (d.flags & SRB_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (d.scale != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= d.scale;
        }
        if (has_mask) {
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
        if (has_res) {
          const uint32_t rrow = erow + (has_mask ? kTileBytes : 0u);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 m = ptx::lds128(rrow + ((((uint32_t)(c0 >> 3) + g) ^ sw) << 4));
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
        for (int j = 0; j < 16; ++j) packed[j] = valid ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : 0u;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          ptx::sts128(orow + ((((uint32_t)(c0 >> 3) + g) ^ sw) << 4),
                      make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]));
        if (d.flags & SRB_COLSUM) {
          // sums of the STORED (bf16-rounded) values over the 32 pixels of this warp: butterfly
          // transpose-reduce, lane l ends with the total of column c0 + l
          float s[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 f = unpack_bf16x2(packed[j]);
            s[2 * j] = f.x;
            s[2 * j + 1] = f.y;
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = upper ? s[i] : s[i + off];
              const float keep = upper ? s[i + off] : s[i];
              s[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          const int g = d.colsum_groups > 1 ? n : 0;
          atomicAdd(p.colsum + (int64_t)g * 64 + c0 + lane, s[0]);
        }
      }
      ptx::tc_fence_before();                 // our tcgen05.ld's are complete (wait::ld)
      ptx::mbar_arrive(&acc_empty[buf]);
      if (p.n_e) ptx::mbar_arrive(&e_empty[buf]);
      ptx::fence_proxy_async_smem();          // staging writes -> visible to the TMA store
      ptx::named_bar_sync(1, 128);
      if (store_thread) {
        ptx::tma_store_4d(&tmY, obase + (uint32_t)buf * kTileBytes, d.y_co, w0, h0, n);
        if (d.flags & SRB_OUT2) ptx::tma_store_4d(&tmY2, obase + (uint32_t)buf * kTileBytes, d.y2_co, w0, h0, n);
        ptx::bulk_commit_group();
        if (ti < 2) TRACE(ti == 0 ? EV_EPI0 : EV_EPI1);
      }
    }
    if (store_thread) {
      ptx::bulk_wait_group<0>();   // all output bytes written before the grid completes
      TRACE(EV_STORED);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0) {
    TRACE(EV_END);
    p.trace[(size_t)blockIdx.x * 16 + EV_T1_NS] = (long long)ptx::globaltimer_ns();
  }
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_nhwc(srb_ctx* ctx, CUtensorMap* map, const void* ptr, int N, int H, int W, int cs, int c_extent, int box_w,
                int box_h, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    srb_set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (N=%d H=%d W=%d cs=%d c_extent=%d box %dx%d)", what,
                  (int)r, N, H, W, cs, c_extent, box_w, box_h);
    return 4;
  }
  return 0;
}

}  // namespace

// 1 if the conv is the resident-filter shape this kernel handles
int srb_conv_c64_ok(const srb_conv_desc* d) {
  if (d->dtype != SRB_BF16 || d->ksize != 3 || d->Cout != 64 || d->Cin < 1 || d->Cin > 64 || d->shuffle > 1) return 0;
  if (d->x_cs % 8 || d->x_co % 8 || d->y_cs % 8 || d->y_co % 8) return 0;
  if ((d->flags & SRB_RESIDUAL) && (d->r_cs % 8 || d->r_co % 8)) return 0;
  if ((d->flags & SRB_MASK) && (d->m_cs % 8 || d->m_co % 8)) return 0;
  if ((d->flags & SRB_OUT2) && (d->y2_cs % 8 || d->y2_co % 8)) return 0;
  if (d->W < 1 || d->H < 1) return 0;
  return 1;
}

int srb_conv_c64(srb_ctx* ctx, const srb_conv_desc* d, const void* x, const void* w, const float* bias, const void* res,
                 const void* mask, void* y, void* y2, float* colsum, cudaStream_t st) {
  SRB_REQUIRE(srb_conv_c64_ok(d), "srb_conv(c64): conv not eligible");
  SRB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                  ((uintptr_t)res & 15) == 0 && ((uintptr_t)mask & 15) == 0 && ((uintptr_t)y2 & 15) == 0,
              "srb_conv(c64): tensors must be 16-byte aligned");
  C64Params p;
  p.d = *d;
  p.bias = bias;
  p.colsum = colsum;
  p.trace = ctx->trace;
  p.tiles_w = srb_cdiv(d->W, kTW);
  p.tiles_h = srb_cdiv(d->H, kTH);
  const int64_t tiles = (int64_t)d->N * p.tiles_w * p.tiles_h;
  SRB_REQUIRE(tiles < (1ll << 31), "srb_conv(c64): too many tiles");
  p.total_tiles = (int)tiles;
  const bool has_mask = (d->flags & SRB_MASK) != 0, has_res = (d->flags & SRB_RESIDUAL) != 0;
  p.n_e = (has_mask ? 1 : 0) + (has_res ? 1 : 0);
  const unsigned grid = (unsigned)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const int tiles_per_cta = srb_cdiv(tiles, grid);
  const uint32_t fixed = kWBytes + (p.n_e ? 4u * kTileBytes : 0u) + 2u * kTileBytes + 1024u;
  int a_stages = (int)(((uint32_t)ctx->smem_optin - 512u /* static */ - fixed) / kWinStride);
  if (a_stages > kMaxAStages) a_stages = kMaxAStages;
  if (a_stages > tiles_per_cta) a_stages = tiles_per_cta;
  SRB_REQUIRE(a_stages >= 1, "srb_conv(c64): not enough shared memory (%d bytes opt-in)", ctx->smem_optin);
  p.a_stages = a_stages;
  p.off_a = kWBytes;
  p.off_e = p.off_a + (uint32_t)a_stages * kWinStride;
  p.off_o = p.off_e + (p.n_e ? 4u * kTileBytes : 0u);
  const size_t smem = (size_t)p.off_o + 2u * kTileBytes + 1024u;

  CUtensorMap tmX, tmW, tmY, tmY2, tmM, tmR;
  int rc = encode_nhwc(ctx, &tmX, x, d->N, d->H, d->W, d->x_cs, d->x_co + d->Cin, kP, kRows, "x window");
  if (rc) return rc;
  rc = encode_nhwc(ctx, &tmY, y, d->N, d->H, d->W, d->y_cs, d->y_co + 64, kTW, kTH, "y tile");
  if (rc) return rc;
  tmY2 = tmY;
  tmM = tmY;
  tmR = tmY;
  if (d->flags & SRB_OUT2) {
    rc = encode_nhwc(ctx, &tmY2, y2, d->N, d->H, d->W, d->y2_cs, d->y2_co + 64, kTW, kTH, "y2 tile");
    if (rc) return rc;
  }
  if (has_mask) {
    rc = encode_nhwc(ctx, &tmM, mask, d->N, d->H, d->W, d->m_cs, d->m_co + 64, kTW, kTH, "mask tile");
    if (rc) return rc;
  }
  if (has_res) {
    rc = encode_nhwc(ctx, &tmR, res, d->N, d->H, d->W, d->r_cs, d->r_co + 64, kTW, kTH, "residual tile");
    if (rc) return rc;
  }
  {
    // packed weights [kw][kh][64 cout][64 cin] bf16 (one 64-channel chunk): 3-D view (cin, cout, tap)
    cuuint64_t dims[3] = {64, 64, 9};
    cuuint64_t strides[2] = {128, 64 * 128};
    cuuint32_t box[3] = {64, 64, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
    CUresult r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
  }

  static int attr_smem = 0;
  if ((int)smem > attr_smem) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = (int)smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  // The kernel requests its filter bank before griddepcontrol.wait.  That is only safe when no
  // kernel that WRITES packed weights can still be running: the first launch after a (re)pack is
  // therefore an ordinary, fully serialised launch; every later one is ordered behind it.
  const bool pdl = !ctx->weights_dirty && !ctx->no_pdl;
  cfg.numAttrs = pdl ? 1 : 0;
  ctx->weights_dirty = 0;
  SRB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_c64_kernel, tmX, tmW, tmY, tmY2, tmM, tmR, p));
  SRB_LAUNCH_CHECK();
  return 0;
}
