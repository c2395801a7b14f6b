// 3x3 convolution with wide layers (Cin a multiple of 64, Cout a multiple of 128): EDSR-large's
// 256 -> 256 / 256 -> 1024 convs (edsr.py:21-33 with n_feats = 256: 95 % of the 52 TFLOP of a
// 960x540 -> 4K frame) and the 64 -> 256 up-sampling convs of every model (common.py:112-139).
//
// Why a second kernel: conv_umma_kernel<128> gives each CTA one 128-pixel x 128-cout tile and reloads a
// 48 KB filter stage (3 kh slices x 128 couts x 64 cin) for every 18 KB activation window: 66 KB of TMA
// traffic per 12 MMAs (768 tensor cycles) against ~65 B/cycle/SM of L2 -> SM bandwidth
// (profiles/r01_hw_probes.txt) = 1015 cycles, with the epilogue and the launch/prologue exposed on top
// (1 CTA/SM): 737 TFLOP/s on the 4K frame, 44 % of the measured peak.  Here
//   * a CTA works on TWO horizontally adjacent 8x16 pixel tiles per filter stage: one 16 x 18-pixel window
//     (36 KB; tile t = window columns 8t..8t+7, i.e. the same buffer at +t*1024 B with a 2048-B
//     stride between 8-pixel row groups) + the 48 KB filter stage feed 24 MMAs (1536 tensor cycles) for
//     84 KB of traffic (1300 cycles): tensor-bound instead of TMA-bound;
//   * CTAs are persistent (grid = #SMs) and TMEM holds two sets of two 128-column accumulators, so the
//     epilogue of one item (2 tiles x 128 couts) runs under the MMAs of the next;
//   * MMAs are issued from an elect.sync region (3 SASS instructions per MMA).
// Epilogue semantics = srb_conv (bias, ReLU, scale, ReLU mask, residual, PixelShuffle store addressing).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 192;                        // TMA producer, MMA issuer, 4 epilogue warps
constexpr int kTW = 8, kTH = 16;                     // one pixel tile = UMMA M = 128
constexpr int kPairW = 2 * kTW;                      // two tiles side by side
constexpr int kRows = kTH + 2;
constexpr uint32_t kABytes = kRows * kPairW * 128;   // 36864: one kw-shifted window, 64 channels
constexpr int kMaxStages = 3;
// BN = 128: Cout a multiple of 128 (EDSR-large); BN = 64: the other multiples of 64 (RDN dense layers
// 128..576 -> 64, the 64 -> 256 up-sampling convs whose PixelShuffle groups are 64 wide, their dgrads)
template <int BN> struct WideCfg {
  static constexpr uint32_t kBBytes = 3u * BN * 128u;            // three kh slices: 49152 / 24576
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;     // 86016 / 61440
  static constexpr int kStages = BN == 128 ? 2 : 3;
  static constexpr uint32_t kTmemCols = 4u * BN;                 // [2 buffers][2 tiles][BN columns]
};

struct WideParams {
  srb_conv_desc d;
  const float* bias;
  const __nv_bfloat16* res;
  const __nv_bfloat16* mask;
  __nv_bfloat16* y;
  int pairs_w, tiles_h, n_tiles_n, total_items, nchunks;
};

template <int BN>
__device__ __forceinline__ void item_coords(const WideParams& p, int item, int& n, int& h0, int& w0, int& n0) {
  const int nt = item % p.n_tiles_n;
  int pp = item / p.n_tiles_n;
  const int pw = pp % p.pairs_w;
  pp /= p.pairs_w;
  const int th = pp % p.tiles_h;
  n = pp / p.tiles_h;
  h0 = th * kTH;
  w0 = pw * kPairW;
  n0 = nt * BN;
}

// one 128-pixel x BN-cout accumulator -> global memory
template <int BN>
__device__ __forceinline__ void wide_epilogue(const WideParams& p, uint32_t tmem_acc, int q, int lane, int n, int h0, int w0,
                                              int n0) {
  const srb_conv_desc& d = p.d;
  const int row = q * 32 + lane;
  const int h = h0 + row / kTW, w = w0 + row % kTW;
  const bool valid = (h < d.H) && (w < d.W);
  int oc0 = n0, oh = h, ow = w, OH = d.H, OW = d.W;
  if (d.shuffle > 1) {          // conv channels are packed in (ij, c') order and 128 divides C'
    const int r = d.shuffle, Cp = d.Cout / (r * r);
    const int ij = n0 / Cp;
    oc0 = n0 % Cp;
    oh = h * r + ij / r;
    ow = w * r + ij % r;
    OH = d.H * r;
    OW = d.W * r;
  }
  const int64_t opix = ((int64_t)n * OH + oh) * OW + ow;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t acc[32];
    ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
    ptx::tmem_ld_wait();
    if (!valid) continue;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
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
    if (d.flags & SRB_MASK) {
      const uint4* mp = reinterpret_cast<const uint4*>(p.mask + opix * d.m_cs + d.m_co + oc0 + c0);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 m = __ldg(mp + g);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(mw[e]);
          if (!(f.x > 0.f)) v[g * 8 + e * 2] = 0.f;
          if (!(f.y > 0.f)) v[g * 8 + e * 2 + 1] = 0.f;
        }
      }
    }
    if (d.flags & SRB_RESIDUAL) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + opix * d.r_cs + d.r_co + oc0 + c0);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 m = __ldg(rp + g);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(mw[e]);
          v[g * 8 + e * 2] += f.x;
          v[g * 8 + e * 2 + 1] += f.y;
        }
      }
    }
    uint4* yp = reinterpret_cast<uint4*>(p.y + opix * d.y_cs + d.y_co + oc0 + c0);
#pragma unroll
    for (int g = 0; g < 4; ++g)
      yp[g] = make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                         pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WideParams p) {
  constexpr uint32_t kBBytes = WideCfg<BN>::kBBytes, kStageBytes = WideCfg<BN>::kStageBytes, kTmemCols = WideCfg<BN>::kTmemCols;
  constexpr int kStages = WideCfg<BN>::kStages;
  (void)kBBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const srb_conv_desc& d = p.d;
  const uint32_t ring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int my_items = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int iters = p.nchunks * 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
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
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmA);
      ptx::prefetch_tensormap(&tmB);
      uint32_t it = 0;
      for (int i = 0; i < my_items; ++i) {
        int n, h0, w0, n0;
        item_coords<BN>(p, (int)blockIdx.x + i * (int)gridDim.x, n, h0, w0, n0);
        for (int chunk = 0; chunk < p.nchunks; ++chunk) {
          for (int kw = 0; kw < 3; ++kw, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            const uint32_t a_dst = ring + s * kStageBytes;
            ptx::tma_load_4d(a_dst, &tmA, &full_bar[s], d.x_co + chunk * 64, w0 + kw - 1, h0 - 1, n);
            ptx::tma_load_3d(a_dst + kABytes, &tmB, &full_bar[s], 0, n0, (chunk * 3 + kw) * 3);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one_sync()) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, BN, 0, 0);
      constexpr uint32_t hi_a = ptx::smem_desc_hi_sw128((uint32_t)kPairW * 128u);   // 8-pixel groups one window row apart
      constexpr uint32_t hi_b = ptx::smem_desc_hi_sw128(1024u);
      uint32_t it = 0;
      for (int i = 0; i < my_items; ++i) {
        const uint32_t buf = (uint32_t)i & 1u;
        ptx::mbar_wait(&acc_empty[buf], (((uint32_t)i >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t acc0 = tmem_base + buf * (2u * BN);
        for (int ci = 0; ci < iters; ++ci, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = ptx::smem_desc_lo(ring + s * kStageBytes, 16u);
          const uint32_t b_lo = ptx::smem_desc_lo(ring + s * kStageBytes + kABytes, 16u);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_lohi(acc0 + (uint32_t)t * (uint32_t)BN, a_lo + (uint32_t)(kh * kPairW * 8 + t * 64 + k * 2), hi_a,
                                    b_lo + (uint32_t)(kh * BN * 8 + k * 2), hi_b, idesc, (ci != 0 || kh != 0 || k != 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    for (int i = 0; i < my_items; ++i) {
      int n, h0, w0, n0;
      item_coords<BN>(p, (int)blockIdx.x + i * (int)gridDim.x, n, h0, w0, n0);
      const uint32_t buf = (uint32_t)i & 1u;
      ptx::mbar_wait(&acc_full[buf], ((uint32_t)i >> 1) & 1u);
      ptx::tc_fence_after();
      wide_epilogue<BN>(p, tmem_base + buf * (2u * BN), q, lane, n, h0, w0, n0);
      wide_epilogue<BN>(p, tmem_base + buf * (2u * BN) + (uint32_t)BN, q, lane, n, h0, w0 + kTW, n0);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN>
int launch_wide(srb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const WideParams& p, cudaStream_t st) {
  const size_t smem = (size_t)WideCfg<BN>::kStages * WideCfg<BN>::kStageBytes + 1024;
  SRB_REQUIRE((int)smem <= ctx->smem_optin, "srb_conv(wide): needs %zu bytes of shared memory", smem);
  static bool attr = false;
  if (!attr) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_wide_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const unsigned grid = (unsigned)(p.total_items < ctx->num_sms ? p.total_items : ctx->num_sms);
  conv_wide_kernel<BN><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  SRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// N-tile width this kernel would use for the conv (128 or 64), 0 if it does not handle it (the caller falls back)
int srb_conv_wide_ok(const srb_conv_desc* d) {
  if (d->dtype != SRB_BF16 || d->ksize != 3) return 0;
  if (d->Cin < 64 || d->Cin % 64) return 0;
  const int rr = d->shuffle > 1 ? d->shuffle * d->shuffle : 1;
  if (d->Cout % rr || (d->Cout / rr) % 64) return 0;
  if (d->flags & ~(SRB_RELU | SRB_RESIDUAL | SRB_MASK)) return 0;
  if (d->x_cs % 8 || d->x_co % 8 || d->y_cs % 8 || d->y_co % 8) return 0;
  if ((d->flags & SRB_RESIDUAL) && (d->r_cs % 8 || d->r_co % 8)) return 0;
  if ((d->flags & SRB_MASK) && (d->m_cs % 8 || d->m_co % 8)) return 0;
  if (d->W < kPairW || d->H < 1) return 0;
  return (d->Cout / rr) % 128 == 0 ? 128 : 64;
}

int srb_conv_wide(srb_ctx* ctx, const srb_conv_desc* d, const void* x, const void* w, const float* bias, const void* res,
                  const void* mask, void* y, cudaStream_t st) {
  const int bn = srb_conv_wide_ok(d);
  SRB_REQUIRE(bn != 0, "srb_conv(wide): conv not eligible");
  SRB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0 &&
                  ((uintptr_t)mask & 15) == 0,
              "srb_conv(wide): tensors must be 16-byte aligned");
  WideParams p;
  p.d = *d;
  p.bias = bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.pairs_w = srb_cdiv(d->W, kPairW);
  p.tiles_h = srb_cdiv(d->H, kTH);
  p.n_tiles_n = d->Cout / bn;
  p.nchunks = d->Cin / 64;
  const int64_t items = (int64_t)d->N * p.tiles_h * p.pairs_w * p.n_tiles_n;
  SRB_REQUIRE(items < (1ll << 31), "srb_conv(wide): too many work items");
  p.total_items = (int)items;

  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)(d->x_co + d->Cin), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_cs * 2, (cuuint64_t)d->W * d->x_cs * 2, (cuuint64_t)d->H * d->W * d->x_cs * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kPairW, (cuuint32_t)kRows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "srb_conv(wide): cuTensorMapEncodeTiled(activations) failed with CUresult %d", (int)r);
  }
  {
    // packed weights [Cin/64][kw][kh][Cout][64]: 3-D view (cin, cout, chunk*9 + kw*3 + kh)
    cuuint64_t dims[3] = {64, (cuuint64_t)d->Cout, (cuuint64_t)p.nchunks * 9};
    cuuint64_t strides[2] = {128, (cuuint64_t)d->Cout * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)bn, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, "srb_conv(wide): cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
  }
  return bn == 128 ? launch_wide<128>(ctx, tmA, tmB, p, st) : launch_wide<64>(ctx, tmA, tmB, p, st);
}
