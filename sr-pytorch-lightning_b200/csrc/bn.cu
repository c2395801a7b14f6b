// BatchNorm2d (training statistics or given statistics) and PReLU for the SRResNet family (SURVEY.md section 8 f3;
// reference: models/srresnet.py:9-36 builds its blocks from common.py:33-55,74-109 with norm=nn.BatchNorm2d(n_feats),
// act=nn.PReLU()).  NHWC tensors, 16-byte channel vectors (C, strides and offsets multiples of 8 bf16 / 4 fp32 elements).
// All of it is HBM-bound element-wise work with per-channel parameters, fused where the reference chains modules:
//   statistics      two passes over x: sum, then sum of squared deviations from the mean (no E[x^2]-E[x]^2 cancellation)
//   forward         y = PReLU_a( (x - mean) * rstd * gamma + beta ) + residual      (each of BN / PReLU / residual optional)
//   backward        one reduction pass (sum gz, sum gz*xhat, sum g*z*[z<=0]) and one pass
//                   dx = gamma * rstd * (gz - mean(gz) - xhat * mean(gz*xhat)),  gz = g * PReLU'(z), z recomputed from x
// Algorithmic bytes per element: statistics 2 reads; forward 1 read (+1 residual) + 1 write; backward 4 reads + 1 write.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

struct BnArgs {
  const void* x;
  const void* g;
  const void* res;
  void* y;
  int x_cs, x_co, g_cs, g_co, r_cs, r_co, y_cs, y_co;
  int C;
  long long npix;
  const float *mean, *rstd, *gamma, *beta, *prelu_a;
  const float* s1;      // statistics pass 2: per-channel sums of pass 1 (mean = s1 / npix)
  float* sums;          // reduction output
  float inv_n;
};

// out[c] += sum over pixels of (x - mean)^P with mean = s1[c] / npix (P = 2) or of x (P = 1)
template <typename T, int P>
__global__ void __launch_bounds__(kBlock) bn_sum_kernel(const BnArgs a) {
  using V = VecT<T>;
  constexpr int W = V::W;
  const int vpp = a.C / W;
  extern __shared__ float acc_s[];        // [C]
  for (int i = threadIdx.x; i < a.C; i += kBlock) acc_s[i] = 0.f;
  __syncthreads();
  const int vc = threadIdx.x % vpp, prow = threadIdx.x / vpp, rows = kBlock / vpp;
  float m[W], acc[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    m[j] = P == 2 ? a.s1[vc * W + j] * a.inv_n : 0.f;
    acc[j] = 0.f;
  }
  const T* x = static_cast<const T*>(a.x);
  if (prow < rows) {
    for (long long p = (long long)blockIdx.x * rows + prow; p < a.npix; p += (long long)gridDim.x * rows) {
      float f[W];
      V::unpack(*reinterpret_cast<const typename V::raw*>(x + p * a.x_cs + a.x_co + vc * W), f);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float d = f[j] - m[j];
        acc[j] += P == 2 ? d * d : d;
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) atomicAdd(&acc_s[vc * W + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.C; i += kBlock) atomicAdd(a.sums + i, acc_s[i]);
}

__global__ void bn_finalize_kernel(const float* s1, const float* s2, int C, float inv_n, float unbias, float eps, float momentum,
                                   float* mean, float* rstd, float* running_mean, float* running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = s1[c] * inv_n, var = s2[c] * inv_n;
  mean[c] = m;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;     // unbiased, as torch
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_apply_kernel(const BnArgs a) {
  using V = VecT<T>;
  constexpr int W = V::W;
  const int vpp = a.C / W;
  const long long total = a.npix * vpp;
  const bool norm = a.mean != nullptr;
  const float slope = a.prelu_a ? *a.prelu_a : 1.f;
  const T* x = static_cast<const T*>(a.x);
  const T* res = static_cast<const T*>(a.res);
  T* y = static_cast<T*>(a.y);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
    const long long p = i / vpp;
    const int c0 = (int)(i - p * vpp) * W;
    float f[W];
    V::unpack(*reinterpret_cast<const typename V::raw*>(x + p * a.x_cs + a.x_co + c0), f);
    float r[W];
    if (res) V::unpack(*reinterpret_cast<const typename V::raw*>(res + p * a.r_cs + a.r_co + c0), r);
#pragma unroll
    for (int j = 0; j < W; ++j) {
      float z = f[j];
      if (norm) z = (z - a.mean[c0 + j]) * a.rstd[c0 + j] * a.gamma[c0 + j] + a.beta[c0 + j];
      if (a.prelu_a) z = z > 0.f ? z : slope * z;
      if (res) z += r[j];
      f[j] = z;
    }
    *reinterpret_cast<typename V::raw*>(y + p * a.y_cs + a.y_co + c0) = V::pack(f);
  }
}

// sums[0][c] = sum gz, sums[1][c] = sum gz * xhat, sums[2][c] = sum g * z * [z <= 0]   (gz = g * PReLU'(z))
template <typename T>
__global__ void __launch_bounds__(kBlock) bn_bwd_reduce_kernel(const BnArgs a) {
  using V = VecT<T>;
  constexpr int W = V::W;
  const int vpp = a.C / W;
  extern __shared__ float acc_s[];        // [3][C]
  for (int i = threadIdx.x; i < 3 * a.C; i += kBlock) acc_s[i] = 0.f;
  __syncthreads();
  const int vc = threadIdx.x % vpp, prow = threadIdx.x / vpp, rows = kBlock / vpp;
  const bool norm = a.mean != nullptr;
  const float slope = a.prelu_a ? *a.prelu_a : 1.f;
  float mu[W], rs[W], ga[W], be[W], s0[W], s1[W], s2[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int c = vc * W + j;
    mu[j] = norm ? a.mean[c] : 0.f;
    rs[j] = norm ? a.rstd[c] : 1.f;
    ga[j] = norm ? a.gamma[c] : 1.f;
    be[j] = norm ? a.beta[c] : 0.f;
    s0[j] = s1[j] = s2[j] = 0.f;
  }
  const T* x = static_cast<const T*>(a.x);
  const T* g = static_cast<const T*>(a.g);
  if (prow < rows) {
    for (long long p = (long long)blockIdx.x * rows + prow; p < a.npix; p += (long long)gridDim.x * rows) {
      float f[W], gg[W];
      V::unpack(*reinterpret_cast<const typename V::raw*>(x + p * a.x_cs + a.x_co + vc * W), f);
      V::unpack(*reinterpret_cast<const typename V::raw*>(g + p * a.g_cs + a.g_co + vc * W), gg);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float xh = (f[j] - mu[j]) * rs[j];
        const float z = xh * ga[j] + be[j];
        float gz = gg[j];
        if (a.prelu_a && !(z > 0.f)) {
          s2[j] += gg[j] * z;
          gz *= slope;
        }
        s0[j] += gz;
        s1[j] += gz * xh;
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      atomicAdd(&acc_s[vc * W + j], s0[j]);
      atomicAdd(&acc_s[a.C + vc * W + j], s1[j]);
      atomicAdd(&acc_s[2 * a.C + vc * W + j], s2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * a.C; i += kBlock) atomicAdd(a.sums + i, acc_s[i]);
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_bwd_apply_kernel(const BnArgs a) {
  using V = VecT<T>;
  constexpr int W = V::W;
  const int vpp = a.C / W;
  const long long total = a.npix * vpp;
  const bool norm = a.mean != nullptr;
  const float slope = a.prelu_a ? *a.prelu_a : 1.f;
  const T* x = static_cast<const T*>(a.x);
  const T* g = static_cast<const T*>(a.g);
  T* dx = static_cast<T*>(a.y);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
    const long long p = i / vpp;
    const int c0 = (int)(i - p * vpp) * W;
    float f[W], gg[W];
    V::unpack(*reinterpret_cast<const typename V::raw*>(x + p * a.x_cs + a.x_co + c0), f);
    V::unpack(*reinterpret_cast<const typename V::raw*>(g + p * a.g_cs + a.g_co + c0), gg);
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int c = c0 + j;
      const float xh = norm ? (f[j] - a.mean[c]) * a.rstd[c] : f[j];
      const float z = norm ? xh * a.gamma[c] + a.beta[c] : xh;
      float gz = gg[j];
      if (a.prelu_a && !(z > 0.f)) gz *= slope;
      f[j] = norm ? a.gamma[c] * a.rstd[c] * (gz - a.sums[c] * a.inv_n - xh * a.sums[a.C + c] * a.inv_n) : gz;
    }
    *reinterpret_cast<typename V::raw*>(dx + p * a.y_cs + a.y_co + c0) = V::pack(f);
  }
}

__global__ void bn_bwd_finalize_kernel(const float* sums, int C, int norm, int accumulate, float* dgamma, float* dbeta, float* da) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C && norm) {
    dbeta[c] = (accumulate ? dbeta[c] : 0.f) + sums[c];
    dgamma[c] = (accumulate ? dgamma[c] : 0.f) + sums[C + c];
  }
  if (c == 0 && da) {
    float t = 0.f;
    for (int i = 0; i < C; ++i) t += sums[2 * C + i];
    da[0] = (accumulate ? da[0] : 0.f) + t;
  }
}

int check_layout(const char* what, int C, int dtype, std::initializer_list<int> strides) {
  const int W = dtype == SRB_BF16 ? 8 : 4;
  SRB_REQUIRE(dtype == SRB_BF16 || dtype == SRB_F32, "%s: dtype must be SRB_F32 or SRB_BF16", what);
  SRB_REQUIRE(C > 0 && C % W == 0 && C <= 1024 && kBlock % (C / W) == 0, "%s: C = %d must be a multiple of %d that divides %d vectors",
              what, C, W, kBlock);
  for (int s : strides) SRB_REQUIRE(s % W == 0, "%s: channel strides / offsets must be multiples of %d", what, W);
  return 0;
}

int grid_for(const srb_ctx* ctx, long long work_items) {
  long long g = (work_items + kBlock - 1) / kBlock;
  const long long cap = (long long)ctx->num_sms * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int srb_bn_stats(srb_ctx* ctx, const void* x, int x_cs, int x_co, int C, int64_t npix, int dtype, float eps, float momentum,
                            float* ws, float* mean, float* rstd, float* running_mean, float* running_var, void* stream) {
  SRB_REQUIRE(ctx && x && ws && mean && rstd && npix > 1, "srb_bn_stats: bad argument");
  if (int rc = check_layout("srb_bn_stats", C, dtype, {x_cs, x_co})) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SRB_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, st));
  BnArgs a = {};
  a.x = x; a.x_cs = x_cs; a.x_co = x_co; a.C = C; a.npix = npix; a.inv_n = 1.f / (float)npix;
  const int W = dtype == SRB_BF16 ? 8 : 4;
  const int rows = kBlock / (C / W);
  const int grid = grid_for(ctx, (npix + rows - 1) / rows * kBlock / 4);
  a.sums = ws;
  if (dtype == SRB_BF16) bn_sum_kernel<__nv_bfloat16, 1><<<grid, kBlock, C * sizeof(float), st>>>(a);
  else bn_sum_kernel<float, 1><<<grid, kBlock, C * sizeof(float), st>>>(a);
  SRB_LAUNCH_CHECK();
  a.s1 = ws;
  a.sums = ws + C;
  if (dtype == SRB_BF16) bn_sum_kernel<__nv_bfloat16, 2><<<grid, kBlock, C * sizeof(float), st>>>(a);
  else bn_sum_kernel<float, 2><<<grid, kBlock, C * sizeof(float), st>>>(a);
  SRB_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, ws + C, C, a.inv_n, (float)npix / (float)(npix - 1), eps, momentum, mean, rstd,
                                                      running_mean, running_var);
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_bn_act_fwd(srb_ctx* ctx, const void* x, int x_cs, int x_co, int C, int64_t npix, int dtype, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, const float* prelu_a, const void* res, int r_cs,
                              int r_co, void* y, int y_cs, int y_co, void* stream) {
  SRB_REQUIRE(ctx && x && y && npix > 0, "srb_bn_act_fwd: bad argument");
  SRB_REQUIRE((mean != nullptr) == (rstd != nullptr) && (mean != nullptr) == (gamma != nullptr) && (mean != nullptr) == (beta != nullptr),
              "srb_bn_act_fwd: mean / rstd / gamma / beta go together");
  if (int rc = check_layout("srb_bn_act_fwd", C, dtype, {x_cs, x_co, y_cs, y_co, res ? r_cs : 0, res ? r_co : 0})) return rc;
  BnArgs a = {};
  a.x = x; a.x_cs = x_cs; a.x_co = x_co; a.res = res; a.r_cs = r_cs; a.r_co = r_co; a.y = y; a.y_cs = y_cs; a.y_co = y_co;
  a.C = C; a.npix = npix; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.prelu_a = prelu_a;
  const int W = dtype == SRB_BF16 ? 8 : 4;
  const int grid = grid_for(ctx, npix * (C / W));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SRB_BF16) bn_apply_kernel<__nv_bfloat16><<<grid, kBlock, 0, st>>>(a);
  else bn_apply_kernel<float><<<grid, kBlock, 0, st>>>(a);
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_bn_act_bwd(srb_ctx* ctx, const void* g, int g_cs, int g_co, const void* x, int x_cs, int x_co, int C, int64_t npix,
                              int dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                              const float* prelu_a, float* ws, void* dx, int dx_cs, int dx_co, float* dgamma, float* dbeta, float* da,
                              int accumulate, void* stream) {
  SRB_REQUIRE(ctx && g && x && dx && ws && npix > 0, "srb_bn_act_bwd: bad argument");
  const bool norm = mean != nullptr;
  SRB_REQUIRE(norm == (rstd != nullptr) && norm == (gamma != nullptr) && norm == (beta != nullptr) && norm == (dgamma != nullptr) &&
                  norm == (dbeta != nullptr),
              "srb_bn_act_bwd: mean / rstd / gamma / beta / dgamma / dbeta go together");
  SRB_REQUIRE((prelu_a != nullptr) == (da != nullptr), "srb_bn_act_bwd: prelu_a and da go together");
  if (int rc = check_layout("srb_bn_act_bwd", C, dtype, {g_cs, g_co, x_cs, x_co, dx_cs, dx_co})) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SRB_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 3 * C, st));
  BnArgs a = {};
  a.g = g; a.g_cs = g_cs; a.g_co = g_co; a.x = x; a.x_cs = x_cs; a.x_co = x_co; a.y = dx; a.y_cs = dx_cs; a.y_co = dx_co;
  a.C = C; a.npix = npix; a.inv_n = 1.f / (float)npix; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.prelu_a = prelu_a;
  a.sums = ws;
  const int W = dtype == SRB_BF16 ? 8 : 4;
  const int rows = kBlock / (C / W);
  const int rgrid = grid_for(ctx, (npix + rows - 1) / rows * kBlock / 4);
  if (dtype == SRB_BF16) bn_bwd_reduce_kernel<__nv_bfloat16><<<rgrid, kBlock, 3 * C * sizeof(float), st>>>(a);
  else bn_bwd_reduce_kernel<float><<<rgrid, kBlock, 3 * C * sizeof(float), st>>>(a);
  SRB_LAUNCH_CHECK();
  const int grid = grid_for(ctx, npix * (C / W));
  if (dtype == SRB_BF16) bn_bwd_apply_kernel<__nv_bfloat16><<<grid, kBlock, 0, st>>>(a);
  else bn_bwd_apply_kernel<float><<<grid, kBlock, 0, st>>>(a);
  SRB_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, C, norm ? 1 : 0, accumulate, dgamma, dbeta, da);
  SRB_LAUNCH_CHECK();
  return 0;
}
