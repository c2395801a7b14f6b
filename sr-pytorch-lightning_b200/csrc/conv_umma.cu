// Implicit-GEMM 3x3 / 1x1 convolution on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA from NHWC bf16 activations.  sm_100a only.
//
// GEMM view:  D[128 pixels, BN couts] = sum over (cin chunk, kw, kh, 4 x k16)  A * B^T
//   A  = input pixels, K-major (64 channels = one 128-byte line per pixel), 128-B swizzle
//   B  = packed weights [Cin/64][kw][kh][Cout][64], K-major, 128-B swizzle
//   D  = fp32 accumulator in TMEM (BN columns x 128 lanes)
//
// Tile = TW x TH output pixels (TW*TH = 128, TW a multiple of 8).  One pipeline stage holds the
// input window for ONE kw shift, TW wide and TH+2 tall (zero padding = TMA out-of-bounds fill),
// plus the 3 kh weight slices of that kw.  Because a TW-pixel row is a whole number of 1024-byte
// swizzle atoms, the kh = 0,1,2 taps are the same buffer at +0, +TW, +2TW lines: three aligned
// UMMA descriptors over one TMA load, so the activations are fetched 3x (kw) instead of 9x.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected
// lane), warps 2-5 = epilogue (TMEM lane quarter = warp % 4).
// Epilogue (include/srb200.h): +bias, ReLU, *res_scale, ReLU-mask, +residual, pixel-shuffle store
// addressing, second output, per-channel sums (CALayer pooling / bias gradients).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 4;
constexpr int kTileM = 128;

struct UmmaConvParams {
  srb_conv_desc d;
  const float* bias;
  const __nv_bfloat16* res;
  const __nv_bfloat16* mask;
  __nv_bfloat16* y;
  __nv_bfloat16* y2;
  float* colsum;
  int tiles_w, tiles_h;
  int TW, TH;
  int KS;           // filter size (1 or 3)
  int nchunks;      // ceil(Cin / 64)
  int a_bytes;      // bytes of one A stage
  int stage_bytes;  // A + B
  int num_stages;
};

// Epilogue of one 128-pixel x BN tile: TMEM -> registers -> (+bias, ReLU, *scale, mask, +residual)
// -> bf16 -> global (pixel-shuffle addressing, optional second output) -> optional column sums.
// Called by the 4 epilogue warps after the accumulator-full barrier; `tmem_acc` is the column base
// of this tile's accumulator.
template <int BN>
__device__ __forceinline__ void conv_epilogue(const UmmaConvParams& p, uint32_t tmem_acc, int warp, int lane, int n,
                                              int h0, int w0, int n0) {
  const srb_conv_desc& d = p.d;
  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int row = q * 32 + lane;     // accumulator row = pixel index inside the tile
  const int h = h0 + row / p.TW, w = w0 + row % p.TW;
  const bool valid = (h < d.H) && (w < d.W);

  // output coordinates; PixelShuffle(r) is a change of address: conv channels are packed in
  // (ij, c') order, and BN divides C' so the whole N-tile maps to one (i, j)
  int oc0 = n0, oh = h, ow = w, OH = d.H, OW = d.W;
  if (d.shuffle > 1) {
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
    constexpr int NV = BN < 32 ? BN : 32;  // live columns in this chunk
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = __uint_as_float(acc[j]);
    if constexpr (BN == 16) {
      if (d.Cout < BN) {
        // narrow output (e.g. the F -> 3 tail conv, edsr.py:32-33): weight rows >= Cout were
        // zero-filled by TMA; only Cout channels exist in y, so store them one by one
        if (valid) {
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            if (j < d.Cout) {
              float o = v[j] + (p.bias ? __ldg(p.bias + j) : 0.f);
              if (d.flags & SRB_RELU) o = fmaxf(o, 0.f);
              p.y[opix * d.y_cs + d.y_co + j] = __float2bfloat16_rn(o * d.scale);
            }
          }
        }
        continue;
      }
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (d.flags & SRB_RELU) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (d.scale != 1.f) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] *= d.scale;
    }
    uint32_t packed[NV / 2];
    if (valid) {
      if (d.flags & SRB_MASK) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + opix * d.m_cs + d.m_co + oc0 + c0);
#pragma unroll
        for (int g = 0; g < NV / 8; ++g) {
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
        for (int g = 0; g < NV / 8; ++g) {
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
#pragma unroll
      for (int j = 0; j < NV / 2; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      uint4* yp = reinterpret_cast<uint4*>(p.y + opix * d.y_cs + d.y_co + oc0 + c0);
#pragma unroll
      for (int g = 0; g < NV / 8; ++g)
        yp[g] = make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
      if (d.flags & SRB_OUT2) {
        uint4* y2p = reinterpret_cast<uint4*>(p.y2 + opix * d.y2_cs + d.y2_co + oc0 + c0);
#pragma unroll
        for (int g = 0; g < NV / 8; ++g)
          y2p[g] = make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV / 2; ++j) packed[j] = 0u;
    }
    if (d.flags & SRB_COLSUM) {
      if constexpr (NV == 32) {
        // sums of the STORED (bf16-rounded) values over the 32 pixels of this warp:
        // butterfly transpose-reduce, lane l ends with the total of column c0 + l
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
        atomicAdd(p.colsum + (int64_t)g * d.Cout + n0 + c0 + lane, s[0]);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads) conv_umma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const srb_conv_desc& d = p.d;

  // 1024-byte aligned operand ring (128-B swizzle atoms must be 1024-B aligned)
  const uint32_t ring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;

  const int tw_i = blockIdx.x % p.tiles_w;
  const int th_i = (blockIdx.x / p.tiles_w) % p.tiles_h;
  const int n = blockIdx.x / (p.tiles_w * p.tiles_h);
  const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
  const int n0 = blockIdx.y * BN;
  const int pad = p.KS >> 1;
  const int iters = p.nchunks * p.KS;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;  // power of two >= 32 (BN in {16,32,64,128})

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(&tmem_full_bar, 1);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmA);
      ptx::prefetch_tensormap(&tmB);
      int it = 0;
      for (int chunk = 0; chunk < p.nchunks; ++chunk) {
        for (int kw = 0; kw < p.KS; ++kw, ++it) {
          const int s = it % p.num_stages;
          const uint32_t ph = (it / p.num_stages) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
          const uint32_t a_dst = ring + (uint32_t)s * p.stage_bytes;
          const uint32_t b_dst = a_dst + p.a_bytes;
          ptx::tma_load_4d(a_dst, &tmA, &full_bar[s], d.x_co + chunk * 64, w0 + kw - pad, h0 - pad, n);
          ptx::tma_load_3d(b_dst, &tmB, &full_bar[s], 0, n0, (chunk * p.KS + kw) * p.KS);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one_sync()) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(kTileM, BN, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.num_stages;
        const uint32_t ph = (it / p.num_stages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_lo = ptx::smem_desc_lo(ring + (uint32_t)s * p.stage_bytes, 16u);
        const uint32_t b_lo = ptx::smem_desc_lo(ring + (uint32_t)s * p.stage_bytes + p.a_bytes, 16u);
        constexpr uint32_t hi = ptx::smem_desc_hi_sw128(1024u);
        const uint32_t a_kh = (uint32_t)p.TW * 8u;        // one tile row of pixels, in 16-byte units
        for (int kh = 0; kh < p.KS; ++kh) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_lohi(tmem_acc, a_lo + kh * a_kh + k * 2u, hi, b_lo + (uint32_t)kh * (BN * 8u) + k * 2u, hi, idesc,
                                (uint32_t)((it | kh | k) != 0));
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
      }
      ptx::umma_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    ptx::mbar_wait(&tmem_full_bar, 0);
    ptx::tc_fence_after();
    conv_epilogue<BN>(p, tmem_acc, warp, lane, n, h0, w0, n0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for the hot layers (3x3, Cin <= 64, Cout == BN): the whole filter bank
// (9 taps x BN rows x 128 B = 72 KB for BN = 64) is loaded ONCE per CTA and stays resident; the CTA
// then walks its share of the pixel tiles (grid = #SMs), streaming only the 18 KB kw-windows through
// an 8-deep TMA ring, with two TMEM accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Versus the streaming kernel this removes the 24 KB/stage weight re-fetch (the kernel
// is bound by L2->SM traffic, not by the tensor pipe) and the per-tile launch/prologue cost.
// ------------------------------------------------------------------------------------------------
constexpr int kResStages = 8;

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_resident_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full;
  __shared__ uint64_t a_full[kResStages];
  __shared__ uint64_t a_empty[kResStages];
  __shared__ uint64_t acc_full[2];
  __shared__ uint64_t acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const srb_conv_desc& d = p.d;
  const uint32_t wbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // [kw][kh][BN][64] bf16
  constexpr uint32_t kWBytes = 9u * BN * 128u;
  const uint32_t ring = wbase + kWBytes;
  const int total_tiles = d.N * p.tiles_w * p.tiles_h;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr uint32_t kTmemCols = 2 * BN;   // two accumulators (BN = 64 -> 128 columns)

  if (threadIdx.x == 0) {
    ptx::mbar_init(&w_full, 1);
    for (int s = 0; s < kResStages; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], 128);   // every epilogue thread arrives
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

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      ptx::prefetch_tensormap(&tmA);
      ptx::prefetch_tensormap(&tmB);
      ptx::mbar_arrive_expect_tx(&w_full, kWBytes);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) ptx::tma_load_3d(wbase + kw * 3u * BN * 128u, &tmB, &w_full, 0, 0, kw * 3);
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int t = (int)blockIdx.x + ti * (int)gridDim.x;
        const int tw_i = t % p.tiles_w, th_i = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
        const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
        for (int kw = 0; kw < 3; ++kw, ++it) {
          const int s = it % kResStages;
          const uint32_t ph = (it / kResStages) & 1;
          ptx::mbar_wait(&a_empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&a_full[s], (uint32_t)p.a_bytes);
          ptx::tma_load_4d(ring + (uint32_t)s * p.a_bytes, &tmA, &a_full[s], d.x_co, w0 + kw - 1, h0 - 1, n);
        }
      }
    }
  } else if (warp == 1) {
    if (my_tiles > 0 && ptx::elect_one_sync()) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(kTileM, BN, 0, 0);
      ptx::mbar_wait(&w_full, 0);
      const uint32_t w_lo = ptx::smem_desc_lo(wbase, 16u);
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = ti & 1;
        ptx::mbar_wait(&acc_empty[buf], ((uint32_t)(ti >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_acc + (uint32_t)(buf * BN);
        for (int kw = 0; kw < 3; ++kw, ++it) {
          const int s = it % kResStages;
          const uint32_t ph = (it / kResStages) & 1;
          ptx::mbar_wait(&a_full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = ptx::smem_desc_lo(ring + (uint32_t)s * p.a_bytes, 16u);
          const uint32_t b_lo = w_lo + (uint32_t)kw * (3u * BN * 8u);
          constexpr uint32_t hi = ptx::smem_desc_hi_sw128(1024u);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(tmem_d, a_lo + kh * 64u + k * 2u, hi, b_lo + kh * (BN * 8u) + k * 2u, hi, idesc,
                                  (kh | k) != 0 ? 1u : (uint32_t)(kw != 0));
          }
          ptx::umma_commit(&a_empty[s]);
        }
        ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int t = (int)blockIdx.x + ti * (int)gridDim.x;
      const int tw_i = t % p.tiles_w, th_i = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
      const int buf = ti & 1;
      ptx::mbar_wait(&acc_full[buf], (uint32_t)(ti >> 1) & 1u);
      ptx::tc_fence_after();
      conv_epilogue<BN>(p, tmem_acc + (uint32_t)(buf * BN), warp, lane, n, th_i * p.TH, tw_i * p.TW, 0);
      ptx::tc_fence_before();            // our tcgen05.ld's are complete (wait::ld) and ordered
      ptx::mbar_arrive(&acc_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(srb_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    srb_set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)", what,
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                  box[0], box[1], box[2]);
    return 4;
  }
  return 0;
}

template <int BN>
int launch(srb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const UmmaConvParams& p, dim3 grid,
           size_t smem, cudaStream_t st) {
  static int attr_smem = 0;
  if ((int)smem > attr_smem) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = (int)smem;
  }
  conv_umma_kernel<BN><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  SRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Which N tile the tcgen05 path would use for this conv, or 0 if the conv is not eligible.
int srb_conv_umma_bn(const srb_conv_desc* d) {
  if (d->dtype != SRB_BF16) return 0;
  if (d->ksize != 1 && d->ksize != 3) return 0;
  if (d->Cin < 1) return 0;
  // any Cin: a partial last 64-channel chunk is zero-filled by TMA (tensor extent = x_co + Cin)
  // and by the weight packing; only the channel STRIDE must keep pixels 16-byte aligned
  if (d->x_cs % 8 || d->x_co % 8) return 0;
  if (d->W < 8 || d->H < 1) return 0;
  if (d->Cout < 16 && d->shuffle <= 1) {
    // narrow output: N tile 16, scalar stores, plain epilogue only
    if (d->flags & ~SRB_RELU) return 0;
    return 16;
  }
  if (d->y_cs % 8 || d->y_co % 8) return 0;
  if ((d->flags & SRB_RESIDUAL) && (d->r_cs % 8 || d->r_co % 8)) return 0;
  if ((d->flags & SRB_MASK) && (d->m_cs % 8 || d->m_co % 8)) return 0;
  if ((d->flags & SRB_OUT2) && (d->y2_cs % 8 || d->y2_co % 8)) return 0;
  int cgroup = d->Cout;  // channels that must stay together in one N tile
  if (d->shuffle > 1) {
    if (d->Cout % (d->shuffle * d->shuffle)) return 0;
    cgroup = d->Cout / (d->shuffle * d->shuffle);
  }
  int bn = 0;
  if (cgroup % 128 == 0) bn = 128;
  else if (cgroup % 64 == 0) bn = 64;
  else if (cgroup % 32 == 0) bn = 32;
  else if (cgroup % 16 == 0) bn = 16;
  else return 0;
  if ((d->flags & SRB_COLSUM) && (bn < 32 || d->shuffle > 1)) return 0;
  return bn;
}

int srb_conv_c64_ok(const srb_conv_desc* d);
int srb_conv_wide_ok(const srb_conv_desc* d);
int srb_conv_wide(srb_ctx*, const srb_conv_desc*, const void*, const void*, const float*, const void*, const void*, void*,
                  cudaStream_t);
int srb_conv_c64(srb_ctx*, const srb_conv_desc*, const void*, const void*, const float*, const void*, const void*, void*,
                 void*, float*, cudaStream_t);

int srb_conv_umma(srb_ctx* ctx, const srb_conv_desc* d, const void* x, const void* w, const float* bias,
                  const void* res, const void* mask, void* y, void* y2, float* colsum, cudaStream_t st) {
  // hot shape (3x3, <=64 -> 64 channels): persistent resident-filter kernel (conv_c64.cu)
  if (srb_conv_c64_ok(d) && !getenv("SRB200_NO_C64")) return srb_conv_c64(ctx, d, x, w, bias, res, mask, y, y2, colsum, st);
  // wide layers (Cin % 64 == 0, Cout % 128 == 0): persistent kernel, two pixel tiles per filter stage (conv_wide.cu)
  if (srb_conv_wide_ok(d) && !getenv("SRB200_NO_WIDE")) return srb_conv_wide(ctx, d, x, w, bias, res, mask, y, st);
  const int BN = srb_conv_umma_bn(d);
  SRB_REQUIRE(BN != 0, "srb_conv(umma): conv not eligible for the tcgen05 path (bf16, k in {1,3}, channel "
              "strides/offsets %% 8 == 0, Cout %% 16 == 0 or < 16, W >= 8): Cin=%d Cout=%d k=%d dtype=%d", d->Cin, d->Cout, d->ksize, d->dtype);
  SRB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0,
              "srb_conv(umma): x, w, y must be 16-byte aligned");
  UmmaConvParams p;
  p.d = *d;
  p.bias = bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.y2 = reinterpret_cast<__nv_bfloat16*>(y2);
  p.colsum = colsum;
  p.KS = d->ksize;
  const int pad = d->ksize / 2;
  // tile shape: 8 wide x 16 tall keeps the halo overhead at 18/16 and divides 48x48 patches
  p.TW = 8;
  p.TH = kTileM / p.TW;
  if (d->H <= 8 && d->W >= 16) {  // short, wide images: 16 x 8
    p.TW = 16;
    p.TH = 8;
  }
  p.tiles_w = srb_cdiv(d->W, p.TW);
  p.tiles_h = srb_cdiv(d->H, p.TH);
  p.nchunks = srb_cdiv(d->Cin, 64);
  p.a_bytes = (p.TH + 2 * pad) * p.TW * 128;
  const int b_bytes = d->ksize * BN * 128;
  p.stage_bytes = p.a_bytes + b_bytes;
  SRB_REQUIRE(p.stage_bytes % 1024 == 0, "srb_conv(umma): internal: stage not 1024-byte aligned");
  const int iters = p.nchunks * d->ksize;
  // 2 stages let two CTAs share an SM for the 64-channel layers; deeper rings for long K loops
  int stages = (2 * p.stage_bytes <= 100 * 1024 && iters <= 3) ? 2 : (int)((200 * 1024) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > iters) stages = iters;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  const size_t smem = (size_t)stages * p.stage_bytes + 1024;
  SRB_REQUIRE((int)smem <= ctx->smem_optin, "srb_conv(umma): needs %zu bytes of shared memory", smem);

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)(d->x_co + d->Cin), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_cs * 2, (cuuint64_t)d->W * d->x_cs * 2, (cuuint64_t)d->H * d->W * d->x_cs * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)(p.TH + 2 * pad), 1};
    int rc = encode_map(ctx, &tmA, x, 4, dims, strides, box, "activations");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)d->Cout, (cuuint64_t)p.nchunks * d->ksize * d->ksize};
    cuuint64_t strides[2] = {128, (cuuint64_t)d->Cout * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)BN, (cuuint32_t)d->ksize};
    int rc = encode_map(ctx, &tmB, w, 3, dims, strides, box, "weights");
    if (rc) return rc;
  }
  const int64_t tiles = (int64_t)d->N * p.tiles_w * p.tiles_h;
  SRB_REQUIRE(tiles < (1ll << 31), "srb_conv(umma): too many tiles");
  if (d->ksize == 3 && p.nchunks == 1 && BN == 64 && d->Cout == 64 && d->shuffle <= 1 && p.TW == 8 &&
      !getenv("SRB200_NO_RESIDENT")) {
    // hot layers: persistent CTAs, filter bank resident in shared memory
    const size_t rsmem = 9u * 64 * 128 + (size_t)kResStages * p.a_bytes + 1024;
    if ((int)rsmem <= ctx->smem_optin) {
      static bool attr = false;
      if (!attr) {
        SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_umma_resident_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        attr = true;
      }
      const unsigned g = (unsigned)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
      conv_umma_resident_kernel<64><<<g, kThreads, rsmem, st>>>(tmA, tmB, p);
      SRB_LAUNCH_CHECK();
      return 0;
    }
  }
  dim3 grid((unsigned)tiles, (unsigned)srb_cdiv(d->Cout, BN));
  switch (BN) {
    case 128: return launch<128>(ctx, tmA, tmB, p, grid, smem, st);
    case 64: return launch<64>(ctx, tmA, tmB, p, grid, smem, st);
    case 32: return launch<32>(ctx, tmA, tmB, p, grid, smem, st);
    default: return launch<16>(ctx, tmA, tmB, p, grid, smem, st);
  }
}
