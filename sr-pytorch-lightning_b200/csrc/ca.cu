// RCAN channel attention (reference models/rcan.py:10-29 CALayer, fused with the RCAB skip add
// rcan.py:54).  HBM/L2-bound streaming kernels; the two tiny FC layers (C -> C/r -> C) are
// recomputed by every block from the pooled sums (512 MACs for C=64, r=16) so that the gate never
// round-trips through global memory between a "gate" and a "scale" launch.
//
// forward :  s = mean_hw(t); z = relu(W1 s + b1); y = sigmoid(W2 z + b2); out = t*y + skip
// backward:  dy = sum_hw(g*t); du = dy*sigmoid'(u); dz = W2^T du; dv = dz*[v>0]; ds = W1^T dv
//            dt = g*y + ds/HW ;  dW2 += du z^T, db2 += du, dW1 += dv s^T, db1 += dv
//
// Streaming layout: a thread owns one 16-byte vector of channels (8 bf16 / 4 fp32) of a pixel;
// every thread first ISSUES the loads of its pixels (up to kPix per thread stay in registers) and
// only then runs the gate MLP, so the dependent-latency chain of the gate (pooled sums -> FC ->
// ReLU -> FC -> sigmoid, ~2 us) overlaps the memory latency instead of preceding it.
#include "common.cuh"

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kPix = 4;  // pixels kept in flight per thread

// sums over H*W per (n, c):  out[n][c] += sum  (out pre-zeroed).  MUL: sum of a*b instead of a.
template <typename T, bool MUL>
__global__ void __launch_bounds__(256) ca_reduce_kernel(const T* __restrict__ a, const T* __restrict__ b, int HW, int C,
                                                        float* __restrict__ out) {
  using V = VecT<T>;
  extern __shared__ float red_s[];  // [lanes][C]
  const int cg = C / V::W;
  const int n = blockIdx.y;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  float acc[V::W];
#pragma unroll
  for (int j = 0; j < V::W; ++j) acc[j] = 0.f;
  if (lane < lanes) {
    const int stride = gridDim.x * lanes;
    for (int p0 = blockIdx.x * lanes + lane; p0 < HW; p0 += stride * kPix) {
      typename V::raw ra[kPix], rb[kPix];
#pragma unroll
      for (int u = 0; u < kPix; ++u) {
        const int p = p0 + u * stride;
        if (p < HW) {
          const int64_t off = ((int64_t)n * HW + p) * C + g * V::W;
          ra[u] = *reinterpret_cast<const typename V::raw*>(a + off);
          if (MUL) rb[u] = *reinterpret_cast<const typename V::raw*>(b + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kPix; ++u) {
        const int p = p0 + u * stride;
        if (p < HW) {
          float fa[V::W], fb[V::W];
          V::unpack(ra[u], fa);
          if (MUL) {
            V::unpack(rb[u], fb);
#pragma unroll
            for (int j = 0; j < V::W; ++j) acc[j] = fmaf(fa[j], fb[j], acc[j]);
          } else {
#pragma unroll
            for (int j = 0; j < V::W; ++j) acc[j] += fa[j];
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < V::W; ++j) red_s[lane * C + g * V::W + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;  // cross-thread sum in fp64 (the per-thread partials are short)
    for (int l = 0; l < lanes; ++l) s += (double)red_s[l * C + c];
    atomicAdd(out + (int64_t)n * C + c, (float)s);
  }
}

// gate MLP from pooled sums; every thread of the block cooperates. smem: s[C] z[Cr] y[C]
__device__ __forceinline__ void ca_gate(const float* __restrict__ sums, float inv_hw, int C, int Cr,
                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                        const float* __restrict__ w2, const float* __restrict__ b2, float* s_s,
                                        float* z_s, float* y_s) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_s[c] = sums[c] * inv_hw;
  __syncthreads();
  // z[j] = relu(b1[j] + sum_c w1[j][c] s[c]) : one warp per j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Cr; j += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += w1[j * C + c] * s_s[c];
    acc = warp_sum(acc);
    if (lane == 0) z_s[j] = fmaxf(acc + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float u = b2[c];
    for (int j = 0; j < Cr; ++j) u += w2[c * Cr + j] * z_s[j];
    y_s[c] = 1.f / (1.f + expf(-u));
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) ca_scale_kernel(const T* __restrict__ t, const T* __restrict__ skip,
                                                       const float* __restrict__ sums, int HW, int C, int Cr,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       T* __restrict__ out, float* __restrict__ s_out,
                                                       float* __restrict__ y_out) {
  using V = VecT<T>;
  extern __shared__ float sm[];
  float* s_s = sm;
  float* y_s = sm + C;
  float* z_s = sm + 2 * C;
  const int n = blockIdx.y;
  const int cg = C / V::W;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  const int stride = gridDim.x * lanes;
  const int pfirst = blockIdx.x * lanes + lane;
  const bool active = lane < lanes;
  // 1) put the first batch of loads in flight
  typename V::raw rt[kPix], rs[kPix];
  if (active) {
#pragma unroll
    for (int u = 0; u < kPix; ++u) {
      const int p = pfirst + u * stride;
      if (p < HW) {
        const int64_t off = ((int64_t)n * HW + p) * C + g * V::W;
        rt[u] = *reinterpret_cast<const typename V::raw*>(t + off);
        if (skip) rs[u] = *reinterpret_cast<const typename V::raw*>(skip + off);
      }
    }
  }
  // 2) gate (dependent chain) while they land
  ca_gate(sums + (int64_t)n * C, 1.f / (float)HW, C, Cr, w1, b1, w2, b2, s_s, z_s, y_s);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      s_out[(int64_t)n * C + c] = s_s[c];
      y_out[(int64_t)n * C + c] = y_s[c];
    }
  }
  if (!active) return;
  float yv[V::W];
#pragma unroll
  for (int j = 0; j < V::W; ++j) yv[j] = y_s[g * V::W + j];
  for (int p0 = pfirst; p0 < HW; p0 += stride * kPix) {
    if (p0 != pfirst) {
#pragma unroll
      for (int u = 0; u < kPix; ++u) {
        const int p = p0 + u * stride;
        if (p < HW) {
          const int64_t off = ((int64_t)n * HW + p) * C + g * V::W;
          rt[u] = *reinterpret_cast<const typename V::raw*>(t + off);
          if (skip) rs[u] = *reinterpret_cast<const typename V::raw*>(skip + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPix; ++u) {
      const int p = p0 + u * stride;
      if (p < HW) {
        float ft[V::W], fs[V::W];
        V::unpack(rt[u], ft);
        if (skip) {
          V::unpack(rs[u], fs);
#pragma unroll
          for (int j = 0; j < V::W; ++j) ft[j] = fmaf(ft[j], yv[j], fs[j]);
        } else {
#pragma unroll
          for (int j = 0; j < V::W; ++j) ft[j] *= yv[j];
        }
        const int64_t off = ((int64_t)n * HW + p) * C + g * V::W;
        *reinterpret_cast<typename V::raw*>(out + off) = V::pack(ft);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ca_bwd_apply_kernel(
    const T* __restrict__ g, const float* __restrict__ s, const float* __restrict__ y, const float* __restrict__ dysum,
    int HW, int C, int Cr, const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, T* __restrict__ dt, float* __restrict__ dw1, float* __restrict__ db1,
    float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ colsum_dt) {
  using V = VecT<T>;
  extern __shared__ float sm[];
  float* s_s = sm;             // [C]
  float* y_s = sm + C;         // [C]
  float* du_s = sm + 2 * C;    // [C]
  float* ds_s = sm + 3 * C;    // [C]  (already divided by HW)
  float* z_s = sm + 4 * C;     // [Cr]
  float* dv_s = z_s + Cr;      // [Cr]
  float* cs_s = dv_s + Cr;     // [lanes][C] column sums of dt (only if colsum_dt)
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane_w = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int cg = C / V::W;
  const int gi = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  const int stride = gridDim.x * lanes;
  const int pfirst = blockIdx.x * lanes + lane;
  const bool active = lane < lanes;
  typename V::raw rg[kPix];
  if (active) {
#pragma unroll
    for (int u = 0; u < kPix; ++u) {
      const int p = pfirst + u * stride;
      if (p < HW) rg[u] = *reinterpret_cast<const typename V::raw*>(g + ((int64_t)n * HW + p) * C + gi * V::W);
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_s[c] = s[(int64_t)n * C + c];
    y_s[c] = y[(int64_t)n * C + c];
  }
  __syncthreads();
  // recompute the hidden layer (needed for dW2 and the ReLU mask)
  for (int j = warp; j < Cr; j += nwarps) {
    float v = 0.f;
    for (int c = lane_w; c < C; c += 32) v += w1[j * C + c] * s_s[c];
    v = warp_sum(v);
    if (lane_w == 0) {
      v += b1[j];
      z_s[j] = fmaxf(v, 0.f);
      dv_s[j] = v > 0.f ? 1.f : 0.f;   // ReLU mask for now
    }
  }
  __syncthreads();
  // sigmoid'(u) = sigmoid(u) * sigmoid(-u), from u itself: no (1 - y) cancellation near saturation
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float u = b2[c];
    for (int j = 0; j < Cr; ++j) u += w2[c * Cr + j] * z_s[j];
    const float sp = 1.f / (1.f + expf(-u)), sn = 1.f / (1.f + expf(u));
    du_s[c] = dysum[(int64_t)n * C + c] * sp * sn;
  }
  __syncthreads();
  for (int j = warp; j < Cr; j += nwarps) {
    float dz = 0.f;
    for (int c = lane_w; c < C; c += 32) dz += w2[c * Cr + j] * du_s[c];
    dz = warp_sum(dz);
    if (lane_w == 0) dv_s[j] *= dz;
  }
  __syncthreads();
  const float inv_hw = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float d = 0.f;
    for (int j = 0; j < Cr; ++j) d += w1[j * C + c] * dv_s[j];
    ds_s[c] = d * inv_hw;
  }
  if (blockIdx.x == 0) {  // parameter gradients, once per sample
    for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
      int c = i / Cr, j = i % Cr;
      atomicAdd(dw2 + i, du_s[c] * z_s[j]);           // w2 [C][Cr]
      int j1 = i / C, c1 = i % C;
      atomicAdd(dw1 + i, dv_s[j1] * s_s[c1]);         // w1 [Cr][C]
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(db2 + c, du_s[c]);
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) atomicAdd(db1 + j, dv_s[j]);
  }
  __syncthreads();
  float csum[V::W];
#pragma unroll
  for (int j = 0; j < V::W; ++j) csum[j] = 0.f;
  if (active) {
    float yv[V::W], dv[V::W];
#pragma unroll
    for (int j = 0; j < V::W; ++j) {
      yv[j] = y_s[gi * V::W + j];
      dv[j] = ds_s[gi * V::W + j];
    }
    for (int p0 = pfirst; p0 < HW; p0 += stride * kPix) {
      if (p0 != pfirst) {
#pragma unroll
        for (int u = 0; u < kPix; ++u) {
          const int p = p0 + u * stride;
          if (p < HW) rg[u] = *reinterpret_cast<const typename V::raw*>(g + ((int64_t)n * HW + p) * C + gi * V::W);
        }
      }
#pragma unroll
      for (int u = 0; u < kPix; ++u) {
        const int p = p0 + u * stride;
        if (p < HW) {
          float f[V::W];
          V::unpack(rg[u], f);
#pragma unroll
          for (int j = 0; j < V::W; ++j) f[j] = fmaf(f[j], yv[j], dv[j]);
          const typename V::raw packed = V::pack(f);
          *reinterpret_cast<typename V::raw*>(dt + ((int64_t)n * HW + p) * C + gi * V::W) = packed;
          if (colsum_dt) {
            float r[V::W];
            V::unpack(packed, r);   // sum what was stored (rounded)
#pragma unroll
            for (int j = 0; j < V::W; ++j) csum[j] += r[j];
          }
        }
      }
    }
  }
  if (colsum_dt) {
    if (active) {
#pragma unroll
      for (int j = 0; j < V::W; ++j) cs_s[lane * C + gi * V::W + j] = csum[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float tot = 0.f;
      for (int l = 0; l < lanes; ++l) tot += cs_s[l * C + c];
      atomicAdd(colsum_dt + c, tot);
    }
  }
}

struct CaLaunch {
  int threads, lanes, slabs;
};

static CaLaunch ca_launch(srb_ctx* ctx, int N, int HW, int C, int vw) {
  CaLaunch l;
  const int cg = C / vw;
  l.lanes = 256 / cg;
  if (l.lanes < 1) l.lanes = 1;
  l.threads = ((cg * l.lanes + 31) / 32) * 32;   // whole warps (extra threads idle in the streaming part)
  // each thread keeps kPix pixels in flight: one pass over the sample needs HW / (lanes*kPix) blocks;
  // cap the total at ~4 blocks per SM
  int want = srb_cdiv(HW, l.lanes * kPix);
  int cap = (ctx->num_sms * 4 + N - 1) / N;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  l.slabs = want;
  return l;
}

static int ca_check(int C, int Cr, int dtype, const char* who) {
  const int vw = dtype == SRB_F32 ? 4 : 8;
  SRB_REQUIRE(C % vw == 0 && C / vw <= 256 && Cr >= 1 && Cr <= 64,
              "%s: unsupported C=%d Cr=%d", who, C, Cr);
  return 0;
}

extern "C" int srb_ca_fwd(srb_ctx* ctx, int N, int H, int W, int C, int Cr, int dtype, const void* t, const void* skip,
                          float* pooled_sum, int compute_pool, const float* w1, const float* b1, const float* w2,
                          const float* b2, void* out, float* s_out, float* y_out, void* stream) {
  SRB_REQUIRE(ctx && t && pooled_sum && w1 && b1 && w2 && b2 && out && s_out && y_out, "srb_ca_fwd: null argument");
  int rc = ca_check(C, Cr, dtype, "srb_ca_fwd");
  if (rc) return rc;
  SRB_REQUIRE(N <= 65535, "srb_ca_fwd: batch too large");
  const int HW = H * W;
  const CaLaunch l = ca_launch(ctx, N, HW, C, dtype == SRB_F32 ? 4 : 8);
  dim3 grid(l.slabs, N);
  if (compute_pool) {
    SRB_CHECK_CUDA(cudaMemsetAsync(pooled_sum, 0, sizeof(float) * (size_t)N * C, S(stream)));
    const size_t rs = sizeof(float) * (size_t)l.lanes * C;
    if (dtype == SRB_F32)
      ca_reduce_kernel<float, false><<<grid, l.threads, rs, S(stream)>>>((const float*)t, nullptr, HW, C, pooled_sum);
    else
      ca_reduce_kernel<__nv_bfloat16, false><<<grid, l.threads, rs, S(stream)>>>((const __nv_bfloat16*)t, nullptr, HW, C, pooled_sum);
    SRB_LAUNCH_CHECK();
  }
  size_t smem = sizeof(float) * (2 * C + Cr);
  if (dtype == SRB_F32)
    ca_scale_kernel<float><<<grid, l.threads, smem, S(stream)>>>((const float*)t, (const float*)skip, pooled_sum, HW, C,
                                                                 Cr, w1, b1, w2, b2, (float*)out, s_out, y_out);
  else
    ca_scale_kernel<__nv_bfloat16><<<grid, l.threads, smem, S(stream)>>>(
        (const __nv_bfloat16*)t, (const __nv_bfloat16*)skip, pooled_sum, HW, C, Cr, w1, b1, w2, b2, (__nv_bfloat16*)out,
        s_out, y_out);
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_ca_bwd(srb_ctx* ctx, int N, int H, int W, int C, int Cr, int dtype, const void* g, const void* t,
                          const float* s, const float* y, const float* w1, const float* b1, const float* w2,
                          const float* b2, void* dt, float* dw1, float* db1, float* dw2, float* db2, float* colsum_dt,
                          float* scratch, int scratch_is_zero, int accumulate, void* stream) {
  SRB_REQUIRE(ctx && g && t && s && y && w1 && b1 && w2 && b2 && dt && dw1 && db1 && dw2 && db2 && scratch,
              "srb_ca_bwd: null argument");
  int rc = ca_check(C, Cr, dtype, "srb_ca_bwd");
  if (rc) return rc;
  const int HW = H * W;
  const CaLaunch l = ca_launch(ctx, N, HW, C, dtype == SRB_F32 ? 4 : 8);
  dim3 grid(l.slabs, N);
  cudaStream_t st = S(stream);
  if (!scratch_is_zero) SRB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)N * C, st));
  if (!accumulate) {
    SRB_CHECK_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * (size_t)C * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * (size_t)C * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * C, st));
    if (colsum_dt) SRB_CHECK_CUDA(cudaMemsetAsync(colsum_dt, 0, sizeof(float) * C, st));
  }
  const size_t rs = sizeof(float) * (size_t)l.lanes * C;
  const size_t smem = sizeof(float) * (4 * C + 2 * Cr) + (colsum_dt ? rs : 0);
  if (dtype == SRB_F32) {
    ca_reduce_kernel<float, true><<<grid, l.threads, rs, st>>>((const float*)g, (const float*)t, HW, C, scratch);
    SRB_LAUNCH_CHECK();
    ca_bwd_apply_kernel<float><<<grid, l.threads, smem, st>>>((const float*)g, s, y, scratch, HW, C, Cr, w1, b1, w2, b2,
                                                              (float*)dt, dw1, db1, dw2, db2, colsum_dt);
  } else {
    ca_reduce_kernel<__nv_bfloat16, true><<<grid, l.threads, rs, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)t,
                                                                        HW, C, scratch);
    SRB_LAUNCH_CHECK();
    ca_bwd_apply_kernel<__nv_bfloat16><<<grid, l.threads, smem, st>>>((const __nv_bfloat16*)g, s, y, scratch, HW, C, Cr,
                                                                      w1, b1, w2, b2, (__nv_bfloat16*)dt, dw1, db1, dw2,
                                                                      db2, colsum_dt);
  }
  SRB_LAUNCH_CHECK();
  return 0;
}
