// RCAN channel attention (reference models/rcan.py:10-29 CALayer, fused with the RCAB skip add
// rcan.py:54).  HBM/L2-bound streaming kernels; the two tiny FC layers (C -> C/r -> C) are
// recomputed by every block from the pooled sums (512 MACs for C=64, r=16) so that the gate never
// round-trips through global memory between a "gate" and a "scale" launch.
//
// forward :  s = mean_hw(t); z = relu(W1 s + b1); y = sigmoid(W2 z + b2); out = t*y + skip
// backward:  dy = sum_hw(g*t); du = dy*y*(1-y); dz = W2^T du; dv = dz*[v>0]; ds = W1^T dv
//            dt = g*y + ds/HW ;  dW2 += du z^T, db2 += du, dW1 += dv s^T, db1 += dv
#include "common.cuh"

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// sums over H*W per (n, c):  out[n][c] += sum  (out pre-zeroed).  MUL: sum of a*b instead of a.
template <typename T, bool MUL>
__global__ void ca_reduce_kernel(const T* __restrict__ a, const T* __restrict__ b, int HW, int C,
                                 float* __restrict__ out) {
  extern __shared__ double red[];  // [C]; cross-thread sums in fp64 (the per-thread partials are short)
  const int cg = C / 4;
  const int n = blockIdx.y;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.0;
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    for (int p = blockIdx.x * lanes + lane; p < HW; p += gridDim.x * lanes) {
      int64_t off = ((int64_t)n * HW + p) * C + g * 4;
      float4 v = ld4(a + off);
      if (MUL) {
        float4 u = ld4(b + off);
        v.x *= u.x; v.y *= u.y; v.z *= u.z; v.w *= u.w;
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(&red[g * 4 + 0], (double)acc.x);
    atomicAdd(&red[g * 4 + 1], (double)acc.y);
    atomicAdd(&red[g * 4 + 2], (double)acc.z);
    atomicAdd(&red[g * 4 + 3], (double)acc.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + (int64_t)n * C + i, (float)red[i]);
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
__global__ void ca_scale_kernel(const T* __restrict__ t, const T* __restrict__ skip, const float* __restrict__ sums,
                                int HW, int C, int Cr, const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2, T* __restrict__ out,
                                float* __restrict__ s_out, float* __restrict__ y_out) {
  extern __shared__ float sm[];
  float* s_s = sm;
  float* y_s = sm + C;
  float* z_s = sm + 2 * C;
  const int n = blockIdx.y;
  ca_gate(sums + (int64_t)n * C, 1.f / (float)HW, C, Cr, w1, b1, w2, b2, s_s, z_s, y_s);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      s_out[(int64_t)n * C + c] = s_s[c];
      y_out[(int64_t)n * C + c] = y_s[c];
    }
  }
  const int cg = C / 4;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  if (lane >= lanes) return;
  const float4 yv = make_float4(y_s[g * 4], y_s[g * 4 + 1], y_s[g * 4 + 2], y_s[g * 4 + 3]);
  for (int p = blockIdx.x * lanes + lane; p < HW; p += gridDim.x * lanes) {
    int64_t off = ((int64_t)n * HW + p) * C + g * 4;
    float4 v = ld4(t + off);
    v.x *= yv.x; v.y *= yv.y; v.z *= yv.z; v.w *= yv.w;
    if (skip) {
      float4 u = ld4(skip + off);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    st4(out + off, v);
  }
}

template <typename T>
__global__ void ca_bwd_apply_kernel(const T* __restrict__ g, const float* __restrict__ s, const float* __restrict__ y,
                                    const float* __restrict__ dysum, int HW, int C, int Cr,
                                    const float* __restrict__ w1, const float* __restrict__ b1,
                                    const float* __restrict__ w2, const float* __restrict__ b2, T* __restrict__ dt,
                                    float* __restrict__ dw1,
                                    float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
                                    float* __restrict__ colsum_dt) {
  extern __shared__ float sm[];
  float* s_s = sm;             // [C]
  float* y_s = sm + C;         // [C]
  float* du_s = sm + 2 * C;    // [C]
  float* ds_s = sm + 3 * C;    // [C]  (already divided by HW)
  float* cs_s = sm + 4 * C;    // [C]  column sums of dt
  float* z_s = sm + 5 * C;     // [Cr]
  float* dv_s = z_s + Cr;      // [Cr]
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane_w = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_s[c] = s[(int64_t)n * C + c];
    y_s[c] = y[(int64_t)n * C + c];
    cs_s[c] = 0.f;
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
  const int cg = C / 4;
  const int gi = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    const float4 yv = make_float4(y_s[gi * 4], y_s[gi * 4 + 1], y_s[gi * 4 + 2], y_s[gi * 4 + 3]);
    const float4 dv = make_float4(ds_s[gi * 4], ds_s[gi * 4 + 1], ds_s[gi * 4 + 2], ds_s[gi * 4 + 3]);
    for (int p = blockIdx.x * lanes + lane; p < HW; p += gridDim.x * lanes) {
      int64_t off = ((int64_t)n * HW + p) * C + gi * 4;
      float4 v = ld4(g + off);
      v.x = v.x * yv.x + dv.x; v.y = v.y * yv.y + dv.y; v.z = v.z * yv.z + dv.z; v.w = v.w * yv.w + dv.w;
      st4(dt + off, v);
      if (colsum_dt) {
        float4 r = ld4(dt + off);  // sum what was stored (rounded)
        csum.x += r.x; csum.y += r.y; csum.z += r.z; csum.w += r.w;
      }
    }
    if (colsum_dt) {
      atomicAdd(&cs_s[gi * 4 + 0], csum.x);
      atomicAdd(&cs_s[gi * 4 + 1], csum.y);
      atomicAdd(&cs_s[gi * 4 + 2], csum.z);
      atomicAdd(&cs_s[gi * 4 + 3], csum.w);
    }
  }
  if (colsum_dt) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(colsum_dt + c, cs_s[c]);
  }
}

static int ca_block_threads(int C) {
  int cg = C / 4;
  int lanes = 256 / cg;
  if (lanes < 1) lanes = 1;
  int th = cg * lanes;
  return ((th + 31) / 32) * 32;  // whole warps (extra threads idle in the streaming part)
}

static int ca_slabs(srb_ctx* ctx, int N, int HW, int threads, int C) {
  int lanes = threads / (C / 4);
  if (lanes < 1) lanes = 1;
  int want = (ctx->num_sms * 4 + N - 1) / N;           // ~4 blocks per SM over the whole batch
  int maxs = (HW + lanes * 4 - 1) / (lanes * 4);       // at least 4 pixels per thread
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  return want;
}

extern "C" int srb_ca_fwd(srb_ctx* ctx, int N, int H, int W, int C, int Cr, int dtype, const void* t, const void* skip,
                          float* pooled_sum, int compute_pool, const float* w1, const float* b1, const float* w2,
                          const float* b2, void* out, float* s_out, float* y_out, void* stream) {
  SRB_REQUIRE(ctx && t && pooled_sum && w1 && b1 && w2 && b2 && out && s_out && y_out, "srb_ca_fwd: null argument");
  SRB_REQUIRE(C % 4 == 0 && C <= 1024 && Cr >= 1 && Cr <= 64, "srb_ca_fwd: unsupported C=%d Cr=%d", C, Cr);
  SRB_REQUIRE(N <= 65535, "srb_ca_fwd: batch too large");
  const int HW = H * W;
  const int threads = ca_block_threads(C);
  const int slabs = ca_slabs(ctx, N, HW, threads, C);
  dim3 grid(slabs, N);
  if (compute_pool) {
    SRB_CHECK_CUDA(cudaMemsetAsync(pooled_sum, 0, sizeof(float) * (size_t)N * C, S(stream)));
    if (dtype == SRB_F32)
      ca_reduce_kernel<float, false><<<grid, threads, C * sizeof(double), S(stream)>>>((const float*)t, nullptr, HW, C, pooled_sum);
    else
      ca_reduce_kernel<__nv_bfloat16, false><<<grid, threads, C * sizeof(double), S(stream)>>>((const __nv_bfloat16*)t, nullptr, HW, C, pooled_sum);
    SRB_LAUNCH_CHECK();
  }
  size_t smem = sizeof(float) * (2 * C + Cr);
  if (dtype == SRB_F32)
    ca_scale_kernel<float><<<grid, threads, smem, S(stream)>>>((const float*)t, (const float*)skip, pooled_sum, HW, C, Cr,
                                                               w1, b1, w2, b2, (float*)out, s_out, y_out);
  else
    ca_scale_kernel<__nv_bfloat16><<<grid, threads, smem, S(stream)>>>((const __nv_bfloat16*)t, (const __nv_bfloat16*)skip,
                                                                       pooled_sum, HW, C, Cr, w1, b1, w2, b2,
                                                                       (__nv_bfloat16*)out, s_out, y_out);
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_ca_bwd(srb_ctx* ctx, int N, int H, int W, int C, int Cr, int dtype, const void* g, const void* t,
                          const float* s, const float* y, const float* w1, const float* b1, const float* w2,
                          const float* b2, void* dt, float* dw1, float* db1, float* dw2, float* db2, float* colsum_dt,
                          float* scratch, int scratch_is_zero, int accumulate, void* stream) {
  SRB_REQUIRE(ctx && g && t && s && y && w1 && b1 && w2 && b2 && dt && dw1 && db1 && dw2 && db2 && scratch,
              "srb_ca_bwd: null argument");
  SRB_REQUIRE(C % 4 == 0 && C <= 1024 && Cr >= 1 && Cr <= 64, "srb_ca_bwd: unsupported C=%d Cr=%d", C, Cr);
  const int HW = H * W;
  const int threads = ca_block_threads(C);
  const int slabs = ca_slabs(ctx, N, HW, threads, C);
  dim3 grid(slabs, N);
  cudaStream_t st = S(stream);
  if (!scratch_is_zero) SRB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)N * C, st));
  if (!accumulate) {
    SRB_CHECK_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * (size_t)C * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * (size_t)C * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * Cr, st));
    SRB_CHECK_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * C, st));
    if (colsum_dt) SRB_CHECK_CUDA(cudaMemsetAsync(colsum_dt, 0, sizeof(float) * C, st));
  }
  size_t smem = sizeof(float) * (5 * C + 2 * Cr);
  if (dtype == SRB_F32) {
    ca_reduce_kernel<float, true><<<grid, threads, C * sizeof(double), st>>>((const float*)g, (const float*)t, HW, C, scratch);
    SRB_LAUNCH_CHECK();
    ca_bwd_apply_kernel<float><<<grid, threads, smem, st>>>((const float*)g, s, y, scratch, HW, C, Cr, w1, b1, w2, b2,
                                                            (float*)dt, dw1, db1, dw2, db2, colsum_dt);
  } else {
    ca_reduce_kernel<__nv_bfloat16, true><<<grid, threads, C * sizeof(double), st>>>((const __nv_bfloat16*)g,
                                                                                     (const __nv_bfloat16*)t, HW, C, scratch);
    SRB_LAUNCH_CHECK();
    ca_bwd_apply_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>((const __nv_bfloat16*)g, s, y, scratch, HW, C, Cr, w1,
                                                                    b1, w2, b2, (__nv_bfloat16*)dt, dw1, db1, dw2, db2, colsum_dt);
  }
  SRB_LAUNCH_CHECK();
  return 0;
}
