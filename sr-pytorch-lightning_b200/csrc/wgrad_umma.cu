// Weight gradient of 3x3 convolutions on tcgen05 tensor cores (sm_100a).
//
//   dW[co][ci][kh][kw] = alpha * sum_{n,h,w} gy[n,h,w,co] * x[n,h+kh-1,w+kw-1,ci]
//
// GEMM view per filter tap: D_tap[ci, co] = sum_pixels X_tap[pixel, ci] * G[pixel, co].  The
// reduction index is the PIXEL, and NHWC keeps channels contiguous, so both operands are
// "MN-major": one 128-byte line = 64 channels of one pixel, 8 pixels = one 1024-byte swizzle atom
// along K.  The same TMA-written tiles the forward kernel uses serve as both operands.
//
// Work unit ("block") = 64 input channels x 64 output channels x 9 taps of one layer.  The 9 taps
// are stacked in pairs along M (UMMA M = 128 = 2 taps x 64 ci; the second tap is just a different
// start address, i.e. the descriptor's leading-dim byte offset), giving 5 accumulators of 64 fp32
// TMEM columns.  A CTA owns a block (or 1/S of its pixels), walks its 128-pixel tiles with a
// 3-stage TMA ring and keeps accumulating in TMEM; only at the very end does it touch global
// memory: plain stores when it owns the whole reduction (S = 1), red.global.add otherwise.
//
// Many layers are batched into ONE launch (their tensor maps travel as kernel parameters): with a
// whole reduction per CTA there is no split-K traffic at all, and 148 CTAs stay busy even though
// one layer's dW is only 64x64x9.  The backward pass defers its weight gradients into such batches
// (srb200/functional.py WgradQueue): they are off the critical path of the dgrad chain.
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 192;       // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr int kStages = 3;
constexpr int kTW = 8, kTH = 16;    // 128-pixel tile, 8 wide (one swizzle atom per tile row)
constexpr int kABytes = (kTH + 2) * kTW * 128;   // one kw-shifted input window (18 rows)
constexpr int kGBytes = kTH * kTW * 128;         // gradient tile
constexpr int kStageBytes = 3 * kABytes + kGBytes;
// Single-window variant (default for batches of 3x3 layers; SRB200_WGRAD_ONEWIN=0 selects the three-window
// form): ONE (8+2)-pixel-wide window per tile serves all three kw shifts (tap (kh,kw) starts (kh*10+kw)*128
// bytes into it; 8-pixel K groups are one 1280-byte window row apart), 39 KB per stage instead of 71 KB ->
// five stages and 45 % less TMA traffic.  The swizzle is a function of the absolute shared-memory address
// for MN-major operands too (K-major: profiles/r01_hw_probes.txt P1; MN-major: the parity tests pass in
// this mode).  Measured: RCAN step 8.58 -> 8.46 ms; a batch with 1x1 layers (RDN) keeps three windows —
// a 1x1 block loads only the centre window there (18 KB instead of 23 KB) and was 2 % faster that way.
constexpr int kPW = kTW + 2;
constexpr int kWinBytes1 = 23 * 1024;            // 18 x 10 x 128 = 23040, padded so that the G tile stays 1024-aligned
constexpr int kWinTx1 = (kTH + 2) * kPW * 128;
constexpr int kStageBytes1 = kWinBytes1 + kGBytes;
constexpr int kStages1 = 5;
// N = 128 variant of the single-window form (default; SRB200_WGRAD_N128=0 selects the N = 64 form above): the gy tile is
// loaded with ONE extra row on top (17 x 8 pixels) and used as two N blocks, [gy shifted up one row | gy], the second simply
// 1024 bytes (one tile row) further (the descriptor's leading-dimension offset, as for the two stacked taps of A).  With
// A = taps (0,kw) | (1,kw) one 128 x 128 accumulator per kw then holds
//     lanes 0-63,  columns 64-127: tap (0,kw)        lanes 64-127, columns 64-127: tap (1,kw)
//     lanes 64-127, columns 0-63 : tap (2,kw)        lanes 0-63,  columns 0-63  : tap (1,kw) again, discarded
// (x[p + 1 row] * gy[p - 1 row + 1 row]: shifting BOTH operands of a tap by one row re-tiles the same pixel sum; the row it
// drops at the bottom edge multiplies the zero padding row of x, the row it adds at the top is TMA's zero fill of gy).
// 24 MMAs of 128 x 128 x 16 per tile instead of 40 of 128 x 64 x 16: an N = 64 MMA is bound by the shared-memory port
// (48 cycles for 32 of math, profiles/r01_hw_probes.txt), N = 128 by the tensor pipe (64 for 64).
constexpr int kGBytes2 = (kTH + 1) * kTW * 128;  // 17 rows
constexpr int kStageBytes2 = kWinBytes1 + kGBytes2;
constexpr int kMaxBlocks = 74;      // blocks per launch (kernel-parameter space: 74 x 320 B < 32 KB)
constexpr uint32_t kTmemCols = 512; // 5 accumulators x 64 columns -> next power of two

struct alignas(64) WgradBlock {
  CUtensorMap tmX;      // x  [N,H,W,Cs] bf16, box (64, 8, 18, 1)
  CUtensorMap tmG;      // gy [N,H,W,Cs] bf16, box (64, 8, 16, 1)
  float* dw;            // OIHW fp32 base of the layer
  int xc0, gc0;         // channel coordinate (offset + chunk * 64) in x / gy
  int ci0, co0;         // first input / output channel of this block in dW (co0 in gy order)
  int Cin, Cout;        // layer sizes (dW strides, shuffle mapping)
  int tiles_w, tiles_h, ntiles;
  int shuffle;
  int16_t rmw;          // 1: dw += (non-atomic, S == 1 and accumulate)
  int16_t k1;           // 1: 1x1 layer (RDN LFF / GFF, rdn.py:37,70): only the centre tap exists
  float alpha;
  int cta0, S;          // this block is processed by CTAs [cta0, cta0 + S) of the launch, each
                        // taking every S-th pixel tile (S proportional to the block's tile count)
};

struct WgradParams {
  WgradBlock blk[kMaxBlocks];
  int nblocks;
  int scatter;      // N = 128 form only: 1 = write dW straight from the TMEM lanes (SRB200_WGRAD_SCATTER=1, for A/B runs)
};

template <int V>      // 0: three kw-shifted windows (also 1x1 layers), 1: single window, 2: single window, N = 128
__global__ void __launch_bounds__(kThreads, 1) wgrad_umma_kernel(const __grid_constant__ WgradParams P) {
  constexpr bool OW = V >= 1, N128 = V == 2;
  constexpr int kStages = OW ? kStages1 : ::kStages;
  constexpr int kStageBytes = N128 ? kStageBytes2 : (OW ? kStageBytes1 : ::kStageBytes);
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bi = 0;
  while (bi + 1 < P.nblocks && (int)blockIdx.x >= P.blk[bi + 1].cta0) ++bi;
  const WgradBlock& B = P.blk[bi];
  const int S = B.S, split = (int)blockIdx.x - B.cta0;
  const uint32_t ring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  // tiles of this CTA: split, split + S, ...
  const int my_tiles = (B.ntiles - split + S - 1) / S;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
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
    if (lane == 0 && my_tiles > 0) {
      ptx::prefetch_tensormap(&B.tmX);
      ptx::prefetch_tensormap(&B.tmG);
      for (int it = 0; it < my_tiles; ++it) {
        const int t = split + it * S;
        const int tw_i = t % B.tiles_w, th_i = (t / B.tiles_w) % B.tiles_h, n = t / (B.tiles_w * B.tiles_h);
        const int h0 = th_i * kTH, w0 = tw_i * kTW;
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        const uint32_t base = ring + (uint32_t)s * kStageBytes;
        if constexpr (OW) {
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(kWinTx1 + (N128 ? kGBytes2 : kGBytes)));
          ptx::tma_load_4d(base, &B.tmX, &full_bar[s], B.xc0, w0 - 1, h0 - 1, n);
          ptx::tma_load_4d(base + kWinBytes1, &B.tmG, &full_bar[s], B.gc0, w0, N128 ? h0 - 1 : h0, n);
        } else {
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(B.k1 ? kABytes + kGBytes : kStageBytes));
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            if (!B.k1 || kw == 1)
              ptx::tma_load_4d(base + kw * kABytes, &B.tmX, &full_bar[s], B.xc0, w0 + kw - 1, h0 - 1, n);
          ptx::tma_load_4d(base + 3 * kABytes, &B.tmG, &full_bar[s], B.gc0, w0, h0, n);
        }
      }
    }
  } else if (warp == 1) {
    if (my_tiles > 0 && ptx::elect_one_sync()) {
      // A and B both MN-major (bits 15, 16), M = 128, N = 64
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 1, 1);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t base = ring + (uint32_t)s * kStageBytes;
        const uint32_t gbase = base + (OW ? kWinBytes1 : 3 * kABytes);
        // accumulator a: two taps stacked along M.  a<3: (kh0,kw=a)+(kh1,kw=a), second atom one
        // tile row (1024 B) further; a=3: (kh2,kw0)+(kh2,kw1), second atom in the next window;
        // a=4: (kh2,kw2) + don't-care rows (upper 64 lanes are discarded by the epilogue)
        constexpr uint32_t hi = ptx::smem_desc_hi_sw128(1024u);
        const uint32_t g_lo = ptx::smem_desc_lo(gbase, 1024u);
        const uint32_t acc_flag = (uint32_t)(it != 0);
        if constexpr (N128) {
          constexpr uint32_t idesc2 = ptx::idesc_bf16_f32(128, 128, 1, 1);
          constexpr uint32_t hi1 = ptx::smem_desc_hi_sw128((uint32_t)kPW * 128u);
          const uint32_t g2_lo = ptx::smem_desc_lo(gbase, 1024u);      // second N block: the unshifted tile, one row further
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const uint32_t a_lo = ptx::smem_desc_lo(base + (uint32_t)(a * 128), (uint32_t)kPW * 128u);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::umma_bf16_lohi(tmem_acc + (uint32_t)a * 128u, a_lo + j * (2u * kPW * 128u / 16u), hi1, g2_lo + j * 128u, hi,
                                  idesc2, j != 0 ? 1u : acc_flag);
          }
        } else
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          if (B.k1 && a != 1) continue;   // 1x1: the centre tap is the upper half of accumulator 1
          if constexpr (OW) {
            // tap (kh, kw) starts at window pixel (kh, kw); second tap of the pair: next row (a < 3, a = 4's
            // discarded half) or next pixel (a = 3); K groups of 8 pixels are one window row (1280 B) apart
            constexpr uint32_t hi1 = ptx::smem_desc_hi_sw128((uint32_t)kPW * 128u);
            const uint32_t a0 = base + (uint32_t)((a < 3 ? a : (a == 3 ? 2 * kPW : 2 * kPW + 2)) * 128);
            const uint32_t a_lo = ptx::smem_desc_lo(a0, a == 3 ? 128u : (uint32_t)kPW * 128u);
#pragma unroll
            for (int j = 0; j < 8; ++j)   // 8 x 16 pixels = 8 x two window rows
              ptx::umma_bf16_lohi(tmem_acc + (uint32_t)a * 64u, a_lo + j * (2u * kPW * 128u / 16u), hi1, g_lo + j * 128u, hi,
                                  idesc, j != 0 ? 1u : acc_flag);
          } else {
            const uint32_t a0 = a < 3 ? base + a * kABytes : (a == 3 ? base + 2 * kTW * 128 : base + 2 * kABytes + 2 * kTW * 128);
            const uint32_t a_lo = ptx::smem_desc_lo(a0, a == 3 ? (uint32_t)kABytes : 1024u);
#pragma unroll
            for (int j = 0; j < 8; ++j)   // 8 x 16 pixels
              ptx::umma_bf16_lohi(tmem_acc + (uint32_t)a * 64u, a_lo + j * 128u, hi, g_lo + j * 128u, hi, idesc,
                                  j != 0 ? 1u : acc_flag);
          }
        }
        ptx::umma_commit(&empty_bar[s]);
      }
      ptx::umma_commit(&tmem_full_bar);
    }
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // TMEM lane = (tap-in-pair, ci)
    const int half = row >> 6, ci = B.ci0 + (row & 63);
    const int rr = B.shuffle > 1 ? B.shuffle * B.shuffle : 1;
    const int Cp = B.Cout / rr;
    ptx::mbar_wait(&tmem_full_bar, 0);
    ptx::tc_fence_after();
    if constexpr (N128) {
      if (!P.scatter) {
        // A TMEM lane holds one (tap, ci) against 64 co: written straight to OIHW that is one float per 32-byte sector (and
        // one red.global per float when the block's pixels are split over CTAs: ~40 us per launch).  Instead the block is
        // staged as [co][ci][9 taps] fp32 in the (now idle) operand ring — for one co that is the layer's contiguous run of
        // 64 x 9 floats — and written out / added with consecutive threads on consecutive floats.
        float* stg = reinterpret_cast<float*>(smem_raw + (ring - ptx::smem_u32(smem_raw)));
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
#pragma unroll 1
          for (int c0 = half ? 0 : 64; c0 < 128; c0 += 32) {
            const int tap = (c0 < 64 ? 2 : half) * 3 + a;
            uint32_t acc[32];
            ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 128 + c0), acc);
            ptx::tmem_ld_wait();
            float* s0 = stg + ((c0 & 63) * 64 + (row & 63)) * 9 + tap;      // lanes: stride 9 floats, conflict-free
#pragma unroll
            for (int j = 0; j < 32; ++j) s0[j * 576] = __uint_as_float(acc[j]) * B.alpha;
          }
        }
        ptx::named_bar_sync(1, 128);
        const int t = (int)threadIdx.x - 64;
        const int run = min(64, B.Cin - B.ci0) * 9;
#pragma unroll 1
        for (int c = 0; c < 64; ++c) {
          const int cop = B.co0 + c;
          if (cop >= B.Cout) break;
          const int co = rr > 1 ? (cop % Cp) * rr + cop / Cp : cop;
          float* dst = B.dw + ((int64_t)co * B.Cin + B.ci0) * 9;
          const float* src = stg + c * 576;
          if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)(run * 4)) & 15) == 0) {      // 16-byte runs: vector red / stores
            float4* dst4 = reinterpret_cast<float4*>(dst);
            const float4* src4 = reinterpret_cast<const float4*>(src);
            if (S > 1) {
              for (int e = t; e < run / 4; e += 128) atomicAdd(dst4 + e, src4[e]);
            } else if (B.rmw) {
              for (int e = t; e < run / 4; e += 128) {
                float4 d = dst4[e];
                const float4 v = src4[e];
                d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
                dst4[e] = d;
              }
            } else {
              for (int e = t; e < run / 4; e += 128) dst4[e] = src4[e];
            }
          } else if (S > 1) {
            for (int e = t; e < run; e += 128) atomicAdd(dst + e, src[e]);
          } else if (B.rmw) {
            for (int e = t; e < run; e += 128) dst[e] += src[e];
          } else {
            for (int e = t; e < run; e += 128) dst[e] = src[e];
          }
        }
      } else
#pragma unroll 1
      for (int a = 0; a < 3; ++a) {          // a = kw
#pragma unroll 1
        for (int c0 = half ? 0 : 64; c0 < 128; c0 += 32) {      // lanes 0-63 own nothing in the shifted columns
          const int kh = c0 < 64 ? 2 : half;
          uint32_t acc[32];
          ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 128 + c0), acc);
          ptx::tmem_ld_wait();
          if (ci < B.Cin) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cop = B.co0 + (c0 & 63) + j;
              if (cop >= B.Cout) continue;
              const int co = rr > 1 ? (cop % Cp) * rr + cop / Cp : cop;
              float* dst = B.dw + (((int64_t)co * B.Cin + ci) * 3 + kh) * 3 + a;
              const float v = __uint_as_float(acc[j]) * B.alpha;
              if (S > 1) atomicAdd(dst, v);
              else if (B.rmw) *dst += v;
              else *dst = v;
            }
          }
        }
      }
    } else
#pragma unroll 1
    for (int a = 0; a < 5; ++a) {
      if (B.k1 && a != 1) continue;
      int kh, kw;
      if (a < 3) { kh = half; kw = a; }
      else if (a == 3) { kh = 2; kw = half; }
      else { kh = 2; kw = 2; }
      const bool live = !(a == 4 && half == 1) && ci < B.Cin && !(B.k1 && half == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t acc[32];
        ptx::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c0), acc);
        ptx::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cop = B.co0 + c0 + j;                      // channel in gy's order
            if (cop >= B.Cout) continue;
            const int co = rr > 1 ? (cop % Cp) * rr + cop / Cp : cop;
            float* dst = B.k1 ? B.dw + (int64_t)co * B.Cin + ci : B.dw + (((int64_t)co * B.Cin + ci) * 3 + kh) * 3 + kw;
            const float v = __uint_as_float(acc[j]) * B.alpha;
            if (S > 1) atomicAdd(dst, v);       // split reduction (dW pre-zeroed or accumulated)
            else if (B.rmw) *dst += v;            // whole reduction here, gradient accumulation
            else *dst = v;
          }
        }
      }
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

int encode_act_map(srb_ctx* ctx, CUtensorMap* map, const void* base, int N, int H, int W, int cs, int cextent,
                   int box_h, int box_w = kTW) {
  cuuint64_t dims[4] = {(cuuint64_t)cextent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    srb_set_error("cuTensorMapEncodeTiled(wgrad) failed with CUresult %d", (int)r);
    return 4;
  }
  return 0;
}

// Launch plan of a batch (host only; also reachable without a device through srb_wgrad_plan for the CPU tests).
void plan_launches(std::vector<WgradBlock>& blocks, int num_sms, std::vector<int>& launch_start, bool& any_split) {
  const int total = (int)blocks.size();
  if (total == 0) return;
  // Group the blocks into launches of <= kMaxBlocks and split every block over S CTAs (each takes every S-th
  // pixel tile) so that all CTAs of a launch carry about the same work and a launch fills the SMs once:
  //   * blocks are sorted by cost (pixel tiles; a 1x1 block issues 8 of the 40 MMAs per tile), so the few big
  //     layers (192x192 tail: 16x the tiles of a 48x48 body layer) meet in one launch;
  //   * launches are packed with S_b = ceil(cost_b / T), T = half the cost of the most common block, until the
  //     CTAs would exceed the SM count — 74 body layers x 2 CTAs, or the big layers + as many body layers as fit;
  //   * within a launch T is then lowered as far as sum(S_b) <= #SMs allows (few blocks: many CTAs each).
  // (Proportional rounding inside fixed 74-block launches left the first RCAN launch — tail, up-sampling and 65
  // body layers — at 116 CTAs with 1.44x the work on the body layers' CTAs: 538 us, 60 % of the SM cycles active.)
  auto cost = [](const WgradBlock& B) -> long long { return (long long)B.ntiles * (B.k1 ? 3 : 10); };
  std::stable_sort(blocks.begin(), blocks.end(), [&](const WgradBlock& x, const WgradBlock& y) { return cost(x) > cost(y); });
  auto ctas_for = [&](const WgradBlock& B, long long T) -> int {
    long long sb = (cost(B) + T - 1) / T;
    if (sb < 1) sb = 1;
    if (sb > B.ntiles) sb = B.ntiles;
    if (sb > 96) sb = 96;
    return (int)sb;
  };
  long long mode_cost = cost(blocks[0]);
  {
    int best = 0, run = 0;
    for (int b = 0; b < total; ++b) {          // sorted: equal costs are adjacent
      run = (b > 0 && cost(blocks[b]) == cost(blocks[b - 1])) ? run + 1 : 1;
      if (run > best) {
        best = run;
        mode_cost = cost(blocks[b]);
      }
    }
  }
  const long long T_pack = mode_cost / 2 > 0 ? (mode_cost + 1) / 2 : 1;
  for (int b0 = 0; b0 < total;) {
    launch_start.push_back(b0);
    int nb = 0, ctas = 0;
    while (b0 + nb < total && nb < kMaxBlocks) {
      const int sb = ctas_for(blocks[b0 + nb], T_pack);
      if (nb > 0 && ctas + sb > num_sms) break;
      ctas += sb;
      ++nb;
    }
    // smallest T whose CTA count still fits the SMs
    long long lo = 1, hi = cost(blocks[b0]);
    while (lo < hi) {
      const long long mid = (lo + hi) / 2;
      long long sum = 0;
      for (int b = 0; b < nb; ++b) sum += ctas_for(blocks[b0 + b], mid);
      if (sum <= num_sms) hi = mid;
      else lo = mid + 1;
    }
    int cta = 0;
    for (int b = 0; b < nb; ++b) {
      WgradBlock& B = blocks[b0 + b];
      B.S = ctas_for(B, lo);
      B.cta0 = cta;
      cta += B.S;
      if (B.S > 1) any_split = true;
    }
    b0 += nb;
  }
}

}  // namespace

int srb_wgrad_umma_ok(const srb_wgrad_desc* d) {
  if (d->dtype != SRB_BF16 || (d->ksize != 3 && d->ksize != 1)) return 0;
  if (d->Cin < 1 || d->Cout < 1) return 0;   // partial 64-channel blocks: TMA zero-fills, epilogue clips
  if (d->x_cs % 8 || d->x_co % 8 || d->g_cs % 8 || d->g_co % 8) return 0;
  if (d->W < 8) return 0;
  if (d->shuffle > 1 && (d->Cout % (d->shuffle * d->shuffle))) return 0;
  return 1;
}

// Launches the batched kernel for a list of eligible items (all checked by the caller).
int srb_wgrad_umma_batched(srb_ctx* ctx, const srb_wgrad_desc* descs, const void* const* xs, const void* const* gys,
                           float* const* dws, int n_items, cudaStream_t st) {
  static bool attr_set = false;
  static const bool one_win_enabled = [] {
    const char* e = getenv("SRB200_WGRAD_ONEWIN");
    return !(e && e[0] == '0');
  }();
  static const bool n128_enabled = [] {
    const char* e = getenv("SRB200_WGRAD_N128");
    return !(e && e[0] == '0');
  }();
  bool one_win = one_win_enabled;
  int n_k1 = 0;
  for (int i = 0; i < n_items; ++i)
    if (descs[i].ksize == 1) ++n_k1;
  if (n_k1 > 0 && n_k1 < n_items && one_win_enabled) {
    // a mixed batch (RDN: dense 3x3 layers + 1x1 LFF / GFF): the 3x3 layers take the single-window N = 128 form in their own
    // launches, the 1x1 layers the three-window form (which loads only their centre window)
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<srb_wgrad_desc> d2;
      std::vector<const void*> x2, g2;
      std::vector<float*> w2;
      for (int i = 0; i < n_items; ++i)
        if ((descs[i].ksize == 1) == (pass == 1)) {
          d2.push_back(descs[i]);
          x2.push_back(xs[i]);
          g2.push_back(gys[i]);
          w2.push_back(dws[i]);
        }
      const int rc = srb_wgrad_umma_batched(ctx, d2.data(), x2.data(), g2.data(), w2.data(), (int)d2.size(), st);
      if (rc) return rc;
    }
    return 0;
  }
  if (n_k1 > 0) one_win = false;
  const bool n128 = one_win && n128_enabled;
  const size_t smem = (n128 ? (size_t)kStages1 * kStageBytes2 : one_win ? (size_t)kStages1 * kStageBytes1 : (size_t)kStages * kStageBytes) + 1024;
  if (!attr_set) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kStages * kStageBytes + 1024)));
    SRB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kStages1 * kStageBytes1 + 1024)));
    SRB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)kStages1 * kStageBytes2 + 1024)));
    attr_set = true;
  }
  std::vector<WgradBlock> blocks;
  for (int i = 0; i < n_items; ++i) {
    const srb_wgrad_desc& d = descs[i];
    CUtensorMap tmX, tmG;  // one pair per layer; its 64x64 blocks differ by channel coordinates only
    int rc = encode_act_map(ctx, &tmX, xs[i], d.N, d.H, d.W, d.x_cs, d.x_co + d.Cin, kTH + 2, one_win ? kPW : kTW);
    if (rc) return rc;
    rc = encode_act_map(ctx, &tmG, gys[i], d.N, d.H, d.W, d.g_cs, d.g_co + d.Cout, n128 ? kTH + 1 : kTH);
    if (rc) return rc;
    const int tiles_w = srb_cdiv(d.W, kTW), tiles_h = srb_cdiv(d.H, kTH);
    for (int cb = 0; cb < srb_cdiv(d.Cin, 64); ++cb) {
      for (int ob = 0; ob < srb_cdiv(d.Cout, 64); ++ob) {
        WgradBlock B;
        B.tmX = tmX;
        B.tmG = tmG;
        B.dw = dws[i];
        B.xc0 = d.x_co + cb * 64;
        B.gc0 = d.g_co + ob * 64;
        B.ci0 = cb * 64;
        B.co0 = ob * 64;
        B.Cin = d.Cin;
        B.Cout = d.Cout;
        B.tiles_w = tiles_w;
        B.tiles_h = tiles_h;
        B.ntiles = d.N * tiles_w * tiles_h;
        B.shuffle = d.shuffle;
        B.rmw = d.accumulate ? 1 : 0;
        B.k1 = d.ksize == 1 ? 1 : 0;
        B.alpha = d.alpha;
        blocks.push_back(B);
      }
    }
  }
  const int total = (int)blocks.size();
  if (total == 0) return 0;
  bool any_split = false;
  std::vector<int> launch_start;
  // a caller that runs these launches beside another kernel (the backward layer chain of the next ResidualGroup, which
  // leaves SMs free) caps the grid so that both are resident at once
  const int sm_cap = ctx->wgrad_sm_budget > 0 && ctx->wgrad_sm_budget < ctx->num_sms ? ctx->wgrad_sm_budget : ctx->num_sms;
  plan_launches(blocks, sm_cap, launch_start, any_split);
  if (any_split) {
    // split reductions add into dW with red.global.add, so overwritten layers start from zero
    for (int i = 0; i < n_items; ++i)
      if (!descs[i].accumulate)
        SRB_CHECK_CUDA(cudaMemsetAsync(dws[i], 0, sizeof(float) * (size_t)descs[i].Cout * descs[i].Cin * descs[i].ksize * descs[i].ksize, st));
    for (auto& B : blocks) B.rmw = 1;
  }
  WgradParams* P = new WgradParams();
  static const int scatter = [] {
    const char* e = getenv("SRB200_WGRAD_SCATTER");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  P->scatter = scatter;
  int rc = 0;
  for (size_t li = 0; li < launch_start.size() && rc == 0; ++li) {
    const int b0 = launch_start[li];
    const int nb = (li + 1 < launch_start.size() ? launch_start[li + 1] : total) - b0;
    for (int b = 0; b < nb; ++b) P->blk[b] = blocks[b0 + b];
    P->nblocks = nb;
    const int ctas = P->blk[nb - 1].cta0 + P->blk[nb - 1].S;
    if (n128) wgrad_umma_kernel<2><<<ctas, kThreads, smem, st>>>(*P);
    else if (one_win) wgrad_umma_kernel<1><<<ctas, kThreads, smem, st>>>(*P);
    else wgrad_umma_kernel<0><<<ctas, kThreads, smem, st>>>(*P);
    __atomic_fetch_add(&g_srb_launches, 1ull, __ATOMIC_RELAXED);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      srb_set_error("wgrad_umma_kernel launch failed: %s", cudaGetErrorString(e));
      rc = 1;
    }
  }
  delete P;
  return rc;
}

int srb_colsum_launch(srb_ctx* ctx, const void* x, int cs, int co, int C, int64_t npix, int dtype, float* out,
                      int accumulate, float alpha, int shuffle, cudaStream_t st);

int srb_wgrad_umma(srb_ctx* ctx, const srb_wgrad_desc* d, const void* x, const void* gy, float* dw, float* dbias,
                   cudaStream_t st) {
  SRB_REQUIRE(srb_wgrad_umma_ok(d), "srb_conv_wgrad(umma): not eligible (bf16, k=3, channel strides %% 8 == 0, W >= 8)");
  int rc = srb_wgrad_umma_batched(ctx, d, &x, &gy, &dw, 1, st);
  if (rc) return rc;
  if (dbias)  // bias gradient = per-channel sums of gy (gy's channel order -> parameter order)
    return srb_colsum_launch(ctx, gy, d->g_cs, d->g_co, d->Cout, (int64_t)d->N * d->H * d->W, d->dtype, dbias,
                             d->accumulate, d->alpha, d->shuffle, st);
  return 0;
}

// Diagnostics / CPU tests: the launch plan for blocks of ntiles[i] pixel tiles (k1[i] != 0: 1x1 layer) on a device of
// num_sms SMs.  Outputs are in plan order (blocks sorted by cost): tiles_out / launch_out / ctas_out [n].
extern "C" int srb_wgrad_plan(int num_sms, int n, const int* ntiles, const int* k1, int* tiles_out, int* launch_out,
                              int* ctas_out) {
  SRB_REQUIRE(num_sms >= 1 && n >= 0 && (n == 0 || (ntiles && k1 && tiles_out && launch_out && ctas_out)),
              "srb_wgrad_plan: bad argument");
  std::vector<WgradBlock> blocks((size_t)n);
  for (int i = 0; i < n; ++i) {
    SRB_REQUIRE(ntiles[i] >= 1, "srb_wgrad_plan: block %d has %d tiles", i, ntiles[i]);
    blocks[i].ntiles = ntiles[i];
    blocks[i].k1 = k1[i] ? 1 : 0;
  }
  std::vector<int> launch_start;
  bool any_split = false;
  plan_launches(blocks, num_sms, launch_start, any_split);
  for (size_t li = 0; li < launch_start.size(); ++li) {
    const int b1 = li + 1 < launch_start.size() ? launch_start[li + 1] : n;
    for (int b = launch_start[li]; b < b1; ++b) {
      tiles_out[b] = blocks[b].ntiles * (blocks[b].k1 ? -1 : 1);     // negative: 1x1 block
      launch_out[b] = (int)li;
      ctas_out[b] = blocks[b].S;
    }
  }
  return 0;
}
