// CUDA-core (FFMA, fp32 accumulate) convolution kernels.
//
// Role in the design: (1) the fp32 parity mode — north_star asks for fp32
// outputs within 1e-4 of the reference, which TF32/bf16 tensor-core math cannot give (SURVEY
// §7.2); (2) the layers the tcgen05 path does not take: Cin = 3 head convs, Cout = 3 tail convs,
// channel counts that are not multiples of 64 (RDN-A), 5x5 / 9x9 filters (SRCNN).
// Same epilogue contract as the tcgen05 kernel (include/srb200.h, srb_conv_desc).
#include "common.cuh"

struct ConvArgs {
  srb_conv_desc d;
  const void* x;
  const void* w;
  const float* bias;
  const void* res;
  const void* mask;
  void* y;
  void* y2;
  float* colsum;
};

constexpr int TS = 8;    // spatial tile edge
constexpr int CK = 16;   // input-channel chunk
constexpr int CO_T = 64; // output channels per block

template <typename T, int K>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  constexpr int HALO = TS + K - 1;
  constexpr int IN_STRIDE = (HALO * HALO) | 1;
  extern __shared__ float smem[];
  float* in_s = smem;                     // [CK][IN_STRIDE]
  float* w_s = smem + CK * IN_STRIDE;     // [K][CK][CO_T]
  __shared__ float colred[CO_T];

  const srb_conv_desc& d = a.d;
  const int tiles_w = (d.W + TS - 1) / TS;
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
  const int n = blockIdx.y;
  const int co0 = blockIdx.z * CO_T;
  const int h0 = th * TS, w0 = tw * TS;
  const int pad = K / 2;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int prow = ty >> 1, pcol0 = (ty & 1) * 4;

  const T* x = reinterpret_cast<const T*>(a.x);
  const float* w = reinterpret_cast<const float*>(a.w);

  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

  if (tid < CO_T) colred[tid] = 0.f;

  for (int ci0 = 0; ci0 < d.Cin; ci0 += CK) {
    __syncthreads();
    // stage the input halo tile for CK channels (zero padded)
    for (int idx = tid; idx < HALO * HALO * CK; idx += 256) {
      int c = idx % CK, pix = idx / CK;
      int r = pix / HALO, q = pix % HALO;
      int h = h0 + r - pad, ww = w0 + q - pad;
      float v = 0.f;
      if (h >= 0 && h < d.H && ww >= 0 && ww < d.W && ci0 + c < d.Cin)
        v = ld_elem(x + (((int64_t)n * d.H + h) * d.W + ww) * d.x_cs + d.x_co + ci0 + c);
      in_s[c * IN_STRIDE + pix] = v;
    }
    for (int kh = 0; kh < K; ++kh) {
      __syncthreads();
      for (int idx = tid; idx < K * CK * CO_T; idx += 256) {
        int co = idx % CO_T, c = (idx / CO_T) % CK, kw = idx / (CO_T * CK);
        float v = 0.f;
        if (ci0 + c < d.Cin && co0 + co < d.Cout)
          v = w[(((int64_t)kh * K + kw) * d.Cin + ci0 + c) * d.Cout + co0 + co];
        w_s[idx] = v;
      }
      __syncthreads();
      // two-level summation: a short inner partial sum per (channel chunk, kh), folded into the
      // running accumulator, keeps the fp32 rounding error ~sqrt(K/48) instead of ~sqrt(K)
      float part[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) part[p][c] = 0.f;
#pragma unroll 4
      for (int c = 0; c < CK; ++c) {
        float xin[4 + K - 1];
#pragma unroll
        for (int j = 0; j < 4 + K - 1; ++j) xin[j] = in_s[c * IN_STRIDE + (prow + kh) * HALO + pcol0 + j];
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          float4 wv = *reinterpret_cast<const float4*>(&w_s[(kw * CK + c) * CO_T + tx * 4]);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            part[p][0] = fmaf(xin[p + kw], wv.x, part[p][0]);
            part[p][1] = fmaf(xin[p + kw], wv.y, part[p][1]);
            part[p][2] = fmaf(xin[p + kw], wv.z, part[p][2]);
            part[p][3] = fmaf(xin[p + kw], wv.w, part[p][3]);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] += part[p][c];
    }
  }

  // ---- epilogue ------------------------------------------------------------------------------
  const int r = d.shuffle > 1 ? d.shuffle : 1;
  const int rr = r * r;
  const int Cp = d.Cout / rr;
  T* y = reinterpret_cast<T*>(a.y);
  T* y2 = reinterpret_cast<T*>(a.y2);
  const T* res = reinterpret_cast<const T*>(a.res);
  const T* mask = reinterpret_cast<const T*>(a.mask);
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  const int h = h0 + prow;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int ww = w0 + pcol0 + p;
    if (h >= d.H || ww >= d.W) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int co = co0 + tx * 4 + c;
      if (co >= d.Cout) continue;
      float v = acc[p][c];
      if (a.bias) v += a.bias[co];
      if (d.flags & SRB_RELU) v = fmaxf(v, 0.f);
      v *= d.scale;
      // output coordinates (pixel shuffle folds into the address)
      int oc = co, oh = h, ow = ww, OH = d.H, OW = d.W;
      if (r > 1) {
        int ij = co / Cp;
        oc = co % Cp;
        oh = h * r + ij / r;
        ow = ww * r + ij % r;
        OH = d.H * r;
        OW = d.W * r;
      }
      const int64_t opix = ((int64_t)n * OH + oh) * OW + ow;
      if (d.flags & SRB_MASK) {
        if (!(ld_elem(mask + opix * d.m_cs + d.m_co + oc) > 0.f)) v = 0.f;
      }
      if (d.flags & SRB_RESIDUAL) v += ld_elem(res + opix * d.r_cs + d.r_co + oc);
      T* dst = y + opix * d.y_cs + d.y_co + oc;
      st_elem(dst, v);
      if (d.flags & SRB_OUT2) st_elem(y2 + opix * d.y2_cs + d.y2_co + oc, v);
      if (d.flags & SRB_COLSUM) csum[c] += ld_elem(dst);  // sum what was stored (rounded)
    }
  }
  if (d.flags & SRB_COLSUM) {
#pragma unroll
    for (int c = 0; c < 4; ++c) atomicAdd(&colred[tx * 4 + c], csum[c]);
    __syncthreads();
    if (tid < CO_T && co0 + tid < d.Cout) {
      int g = d.colsum_groups > 1 ? n : 0;
      atomicAdd(a.colsum + (int64_t)g * d.Cout + co0 + tid, colred[tid]);
    }
  }
}

template <typename T, int K>
static int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  constexpr int HALO = TS + K - 1;
  constexpr int IN_STRIDE = (HALO * HALO) | 1;
  size_t smem = sizeof(float) * (CK * IN_STRIDE + K * CK * CO_T);
  static bool attr_set = false;
  if (!attr_set) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(conv_simt_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const srb_conv_desc& d = a.d;
  dim3 grid(srb_cdiv(d.W, TS) * srb_cdiv(d.H, TS), d.N, srb_cdiv(d.Cout, CO_T));
  conv_simt_kernel<T, K><<<grid, 256, smem, st>>>(a);
  SRB_LAUNCH_CHECK();
  return 0;
}

int srb_conv_simt(srb_ctx* ctx, const srb_conv_desc* d, const void* x, const void* w, const float* bias,
                  const void* res, const void* mask, void* y, void* y2, float* colsum, cudaStream_t st) {
  ConvArgs a{*d, x, w, bias, res, mask, y, y2, colsum};
  SRB_REQUIRE(d->ksize == 1 || d->ksize == 3 || d->ksize == 5 || d->ksize == 9,
              "srb_conv(simt): unsupported kernel size %d (1,3,5,9)", d->ksize);
  SRB_REQUIRE(d->N <= 65535, "srb_conv(simt): batch %d too large", d->N);
#define SRB_DISPATCH_K(T)                                       \
  switch (d->ksize) {                                           \
    case 1: return launch_conv_simt<T, 1>(a, st);               \
    case 3: return launch_conv_simt<T, 3>(a, st);               \
    case 5: return launch_conv_simt<T, 5>(a, st);               \
    default: return launch_conv_simt<T, 9>(a, st);              \
  }
  if (d->dtype == SRB_F32) {
    SRB_DISPATCH_K(float)
  } else {
    SRB_DISPATCH_K(__nv_bfloat16)
  }
#undef SRB_DISPATCH_K
}

// =============================================================================================
// weight / bias gradient
// =============================================================================================
struct WgradArgs {
  srb_wgrad_desc d;
  const void* x;
  const void* gy;
  float* dw;
  float* dbias;
  int nsplit;
  int total_tiles;
};

template <typename T, int K>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a) {
  constexpr int XW = TS + K - 1;
  __shared__ float g_s[TS * TS][CO_T];       // 16 KB
  __shared__ float x_s[CK][TS][XW | 1];
  const srb_wgrad_desc& d = a.d;
  const int co_chunks = (d.Cout + CO_T - 1) / CO_T;
  const int co0 = (blockIdx.x % co_chunks) * CO_T;
  const int ci0 = (blockIdx.x / co_chunks) * CK;
  const int kh = blockIdx.y;
  const int pad = K / 2;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_w = (d.W + TS - 1) / TS, tiles_h = (d.H + TS - 1) / TS;
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* gy = reinterpret_cast<const T*>(a.gy);

  float acc[K][4];
#pragma unroll
  for (int kw = 0; kw < K; ++kw)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[kw][c] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = a.dbias != nullptr && ci0 == 0 && kh == 0 && ty == 0;

  for (int t = blockIdx.z; t < a.total_tiles; t += a.nsplit) {
    const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, n = t / (tiles_w * tiles_h);
    const int h0 = th * TS, w0 = tw * TS;
    __syncthreads();
    for (int idx = tid; idx < TS * TS * CO_T; idx += 256) {
      int co = idx % CO_T, p = idx / CO_T;
      int h = h0 + p / TS, ww = w0 + p % TS;
      float v = 0.f;
      if (h < d.H && ww < d.W && co0 + co < d.Cout)
        v = ld_elem(gy + (((int64_t)n * d.H + h) * d.W + ww) * d.g_cs + d.g_co + co0 + co);
      g_s[p][co] = v;
    }
    for (int idx = tid; idx < CK * TS * XW; idx += 256) {
      int c = idx % CK, q = (idx / CK) % XW, r = idx / (CK * XW);
      int h = h0 + r + kh - pad, ww = w0 + q - pad;
      float v = 0.f;
      if (h >= 0 && h < d.H && ww >= 0 && ww < d.W && ci0 + c < d.Cin)
        v = ld_elem(x + (((int64_t)n * d.H + h) * d.W + ww) * d.x_cs + d.x_co + ci0 + c);
      x_s[c][r][q] = v;
    }
    __syncthreads();
    float part[K][4];  // per-tile partial sums (two-level summation, see conv kernel)
    float bpart[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
#pragma unroll
      for (int c = 0; c < 4; ++c) part[kw][c] = 0.f;
#pragma unroll 4
    for (int p = 0; p < TS * TS; ++p) {
      const int r = p / TS, c = p % TS;
      float4 g4 = *reinterpret_cast<const float4*>(&g_s[p][tx * 4]);
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        float xv = x_s[ty][r][c + kw];
        part[kw][0] = fmaf(g4.x, xv, part[kw][0]);
        part[kw][1] = fmaf(g4.y, xv, part[kw][1]);
        part[kw][2] = fmaf(g4.z, xv, part[kw][2]);
        part[kw][3] = fmaf(g4.w, xv, part[kw][3]);
      }
      if (do_bias) {
        bpart[0] += g4.x; bpart[1] += g4.y; bpart[2] += g4.z; bpart[3] += g4.w;
      }
    }
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[kw][c] += part[kw][c];
#pragma unroll
    for (int c = 0; c < 4; ++c) bsum[c] += bpart[c];
  }

  const int rr = d.shuffle > 1 ? d.shuffle * d.shuffle : 1;
  const int Cp = d.Cout / rr;
  const int ci = ci0 + ty;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cop = co0 + tx * 4 + c;   // channel index in gy's (possibly permuted) order
    if (cop >= d.Cout) continue;
    const int co = rr > 1 ? (cop % Cp) * rr + cop / Cp : cop;
    if (ci < d.Cin) {
#pragma unroll
      for (int kw = 0; kw < K; ++kw)
        atomicAdd(a.dw + (((int64_t)co * d.Cin + ci) * K + kh) * K + kw, acc[kw][c] * d.alpha);
    }
    if (do_bias) atomicAdd(a.dbias + co, bsum[c] * d.alpha);
  }
}

template <typename T, int K>
static int launch_wgrad_simt(WgradArgs& a, srb_ctx* ctx, cudaStream_t st) {
  const srb_wgrad_desc& d = a.d;
  int blocks_base = srb_cdiv(d.Cout, CO_T) * srb_cdiv(d.Cin, CK) * K;
  a.total_tiles = d.N * srb_cdiv(d.H, TS) * srb_cdiv(d.W, TS);
  int nsplit = (ctx->num_sms * 4 + blocks_base - 1) / blocks_base;
  if (nsplit > a.total_tiles) nsplit = a.total_tiles;
  if (nsplit > 65535) nsplit = 65535;
  if (nsplit < 1) nsplit = 1;
  a.nsplit = nsplit;
  dim3 grid(srb_cdiv(d.Cout, CO_T) * srb_cdiv(d.Cin, CK), K, nsplit);
  wgrad_simt_kernel<T, K><<<grid, 256, 0, st>>>(a);
  SRB_LAUNCH_CHECK();
  return 0;
}

int srb_wgrad_simt(srb_ctx* ctx, const srb_wgrad_desc* d, const void* x, const void* gy, float* dw, float* dbias,
                   cudaStream_t st) {
  SRB_REQUIRE(d->ksize == 1 || d->ksize == 3 || d->ksize == 5 || d->ksize == 9,
              "srb_conv_wgrad(simt): unsupported kernel size %d (1,3,5,9)", d->ksize);
  if (!d->accumulate) {
    SRB_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * d->ksize * d->ksize, st));
    if (dbias) SRB_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * d->Cout, st));
  }
  WgradArgs a{*d, x, gy, dw, dbias, 1, 0};
#define SRB_DISPATCH_K(T)                                          \
  switch (d->ksize) {                                              \
    case 1: return launch_wgrad_simt<T, 1>(a, ctx, st);            \
    case 3: return launch_wgrad_simt<T, 3>(a, ctx, st);            \
    case 5: return launch_wgrad_simt<T, 5>(a, ctx, st);            \
    default: return launch_wgrad_simt<T, 9>(a, ctx, st);           \
  }
  if (d->dtype == SRB_F32) {
    SRB_DISPATCH_K(float)
  } else {
    SRB_DISPATCH_K(__nv_bfloat16)
  }
#undef SRB_DISPATCH_K
}
