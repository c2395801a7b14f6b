// Bandwidth-bound helpers: weight packing, boundary layout conversion, channel-slice copies/adds,
// pixel-unshuffle, column sums, L1 loss (+ seed gradient) and Adam.  All HBM/L2-bound: coalesced,
// vectorised where alignment allows, grid sized from the element count.
#include "common.cuh"

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int perm_channel(int co, int Cout, int r) {
  // nn.PixelShuffle(r) reads conv channel co = c'*r*r + ij; we store channels in (ij, c') order
  if (r <= 1) return co;
  int rr = r * r, Cp = Cout / rr;
  return (co % rr) * Cp + co / rr;
}

__global__ void pack_simt_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int mode, int shuffle,
                                 float* __restrict__ out) {
  int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the source OIHW
    int kw = i % k;
    int kh = (i / k) % k;
    int ci = (i / (k * k)) % Cin;
    int co = i / ((int64_t)k * k * Cin);
    int cop = perm_channel(co, Cout, shuffle);
    float v = w[i];
    if (mode == SRB_PACK_FWD) {
      out[(((int64_t)kh * k + kw) * Cin + ci) * Cout + cop] = v;
    } else {
      // dgrad conv: input channels = (permuted) co, output channels = ci, taps rotated by 180 degrees
      out[(((int64_t)(k - 1 - kh) * k + (k - 1 - kw)) * Cout + cop) * Cin + ci] = v;
    }
  }
}

__global__ void pack_umma_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int mode, int shuffle,
                                 __nv_bfloat16* __restrict__ out, int cin_e, int cout_e) {
  // out[chunk][kw][kh][row(cout_e)][64]; effective conv has cin_e inputs, cout_e outputs
  int nchunk = (cin_e + 63) / 64;
  int64_t total = (int64_t)nchunk * k * k * cout_e * 64;
  int rr = shuffle > 1 ? shuffle * shuffle : 1;
  int Cp = Cout / rr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kk = i % 64;
    int row = (i / 64) % cout_e;
    int kh = (i / ((int64_t)64 * cout_e)) % k;
    int kw = (i / ((int64_t)64 * cout_e * k)) % k;
    int chunk = i / ((int64_t)64 * cout_e * k * k);
    int cin_idx = chunk * 64 + kk;
    float v = 0.f;
    if (cin_idx < cin_e) {
      if (mode == SRB_PACK_FWD) {
        // row = permuted output channel -> original co
        int co = row;
        if (rr > 1) co = (row % Cp) * rr + row / Cp;
        v = w[(((int64_t)co * Cin + cin_idx) * k + kh) * k + kw];
      } else {
        // effective input channel cin_idx = permuted co; effective output row = ci
        int co = cin_idx;
        if (rr > 1) co = (cin_idx % Cp) * rr + cin_idx / Cp;
        v = w[(((int64_t)co * Cin + row) * k + (k - 1 - kh)) * k + (k - 1 - kw)];
      }
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

extern "C" size_t srb_packed_weight_bytes(int Cout, int Cin, int ksize, int packing, int mode) {
  int cin_e = mode == SRB_PACK_FWD ? Cin : Cout;
  int cout_e = mode == SRB_PACK_FWD ? Cout : Cin;
  if (packing == SRB_PACK_SIMT) return (size_t)Cout * Cin * ksize * ksize * sizeof(float);
  return (size_t)((cin_e + 63) / 64) * ksize * ksize * cout_e * 64 * sizeof(__nv_bfloat16);
}

extern "C" int srb_pack_weight(srb_ctx* ctx, const float* w, int Cout, int Cin, int k, int packing, int mode,
                               int shuffle, void* out, void* stream) {
  SRB_REQUIRE(ctx && w && out, "srb_pack_weight: null argument");
  SRB_REQUIRE(shuffle == 0 || (Cout % (shuffle * shuffle)) == 0, "srb_pack_weight: Cout %d not divisible by r^2", Cout);
  ctx->weights_dirty = 1;
  int64_t total = (int64_t)Cout * Cin * k * k;
  if (packing == SRB_PACK_SIMT) {
    int blocks = srb_cdiv(total, 256);
    if (blocks > 4096) blocks = 4096;
    pack_simt_kernel<<<blocks, 256, 0, S(stream)>>>(w, Cout, Cin, k, mode, shuffle, (float*)out);
  } else {
    int cin_e = mode == SRB_PACK_FWD ? Cin : Cout;
    int cout_e = mode == SRB_PACK_FWD ? Cout : Cin;
    int64_t tot2 = (int64_t)((cin_e + 63) / 64) * k * k * cout_e * 64;
    int blocks = srb_cdiv(tot2, 256);
    if (blocks > 4096) blocks = 4096;
    pack_umma_kernel<<<blocks, 256, 0, S(stream)>>>(w, Cout, Cin, k, mode, shuffle, (__nv_bfloat16*)out, cin_e, cout_e);
  }
  SRB_LAUNCH_CHECK();
  return 0;
}

__global__ void pack_bias_kernel(const float* __restrict__ b, int Cout, int shuffle, float* __restrict__ out) {
  int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co < Cout) out[perm_channel(co, Cout, shuffle)] = b[co];
}

extern "C" int srb_pack_bias(srb_ctx* ctx, const float* bias, int Cout, int shuffle, float* out, void* stream) {
  SRB_REQUIRE(ctx && bias && out, "srb_pack_bias: null argument");
  pack_bias_kernel<<<srb_cdiv(Cout, 128), 128, 0, S(stream)>>>(bias, Cout, shuffle, out);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// boundary layout conversion
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                    const float* __restrict__ add, T* __restrict__ y, int cs, int co) {
  int64_t npix = (int64_t)N * H * W;
  int64_t hw = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = p / hw, r = p % hw;
    for (int c = 0; c < C; ++c) {
      float v = x[(n * C + c) * hw + r];
      if (add) v += add[c];
      st_elem(y + p * cs + co + c, v);
    }
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int cs, int co, int N, int C, int H, int W,
                                    const float* __restrict__ add, float* __restrict__ y) {
  int64_t npix = (int64_t)N * H * W;
  int64_t hw = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = p / hw, r = p % hw;
    for (int c = 0; c < C; ++c) {
      float v = ld_elem(x + p * cs + co + c);
      if (add) v += add[c];
      y[(n * C + c) * hw + r] = v;
    }
  }
}

extern "C" int srb_nchw_to_nhwc(srb_ctx* ctx, const float* x, int N, int C, int H, int W, const float* add, int dtype,
                                void* y, int cs, int co, void* stream) {
  SRB_REQUIRE(ctx && x && y, "srb_nchw_to_nhwc: null argument");
  int64_t npix = (int64_t)N * H * W;
  int blocks = srb_cdiv(npix, 256);
  if (dtype == SRB_F32)
    nchw_to_nhwc_kernel<float><<<blocks, 256, 0, S(stream)>>>(x, N, C, H, W, add, (float*)y, cs, co);
  else
    nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, S(stream)>>>(x, N, C, H, W, add, (__nv_bfloat16*)y, cs, co);
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_nhwc_to_nchw(srb_ctx* ctx, const void* x, int cs, int co, int dtype, int N, int C, int H, int W,
                                const float* add, float* y, void* stream) {
  SRB_REQUIRE(ctx && x && y, "srb_nhwc_to_nchw: null argument");
  int64_t npix = (int64_t)N * H * W;
  int blocks = srb_cdiv(npix, 256);
  if (dtype == SRB_F32)
    nhwc_to_nchw_kernel<float><<<blocks, 256, 0, S(stream)>>>((const float*)x, cs, co, N, C, H, W, add, y);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, S(stream)>>>((const __nv_bfloat16*)x, cs, co, N, C, H, W, add, y);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// channel-slice copy / add, pixel-unshuffle  (4 channels per thread; requires C,cs,co % 4 == 0,
// scalar fallback otherwise)
// ---------------------------------------------------------------------------------------------
template <typename T, bool ADD>
__global__ void slice_kernel(const T* __restrict__ a, int a_cs, int a_co, const T* __restrict__ b, int b_cs, int b_co,
                             T* __restrict__ out, int o_cs, int o_co, int C4, int64_t npix) {
  int64_t total = npix * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C4;
    int c = (int)(i % C4) * 4;
    float4 v = ld4(a + p * a_cs + a_co + c);
    if (ADD) {
      float4 u = ld4(b + p * b_cs + b_co + c);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    st4(out + p * o_cs + o_co + c, v);
  }
}

template <typename T, bool ADD>
__global__ void slice_scalar_kernel(const T* __restrict__ a, int a_cs, int a_co, const T* __restrict__ b, int b_cs,
                                    int b_co, T* __restrict__ out, int o_cs, int o_co, int C, int64_t npix) {
  int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int c = (int)(i % C);
    float v = ld_elem(a + p * a_cs + a_co + c);
    if (ADD) v += ld_elem(b + p * b_cs + b_co + c);
    st_elem(out + p * o_cs + o_co + c, v);
  }
}

template <typename T, bool ADD>
static int launch_slice(const void* a, int a_cs, int a_co, const void* b, int b_cs, int b_co, void* out, int o_cs,
                        int o_co, int C, int64_t npix, cudaStream_t st) {
  bool vec = (C % 4 == 0) && (a_cs % 4 == 0) && (a_co % 4 == 0) && (o_cs % 4 == 0) && (o_co % 4 == 0) &&
             (!ADD || ((b_cs % 4 == 0) && (b_co % 4 == 0)));
  if (vec) {
    int64_t total = npix * (C / 4);
    int blocks = srb_cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    slice_kernel<T, ADD><<<blocks, 256, 0, st>>>((const T*)a, a_cs, a_co, (const T*)b, b_cs, b_co, (T*)out, o_cs, o_co,
                                                  C / 4, npix);
  } else {
    int64_t total = npix * C;
    int blocks = srb_cdiv(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    slice_scalar_kernel<T, ADD><<<blocks, 256, 0, st>>>((const T*)a, a_cs, a_co, (const T*)b, b_cs, b_co, (T*)out, o_cs,
                                                         o_co, C, npix);
  }
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_copy_channels(srb_ctx* ctx, const void* src, int s_cs, int s_co, void* dst, int d_cs, int d_co, int C,
                                 int64_t npix, int dtype, void* stream) {
  SRB_REQUIRE(ctx && src && dst, "srb_copy_channels: null argument");
  if (dtype == SRB_F32)
    return launch_slice<float, false>(src, s_cs, s_co, nullptr, 0, 0, dst, d_cs, d_co, C, npix, S(stream));
  return launch_slice<__nv_bfloat16, false>(src, s_cs, s_co, nullptr, 0, 0, dst, d_cs, d_co, C, npix, S(stream));
}

extern "C" int srb_add_channels(srb_ctx* ctx, const void* a, int a_cs, int a_co, const void* b, int b_cs, int b_co,
                                void* out, int o_cs, int o_co, int C, int64_t npix, int dtype, void* stream) {
  SRB_REQUIRE(ctx && a && b && out, "srb_add_channels: null argument");
  if (dtype == SRB_F32)
    return launch_slice<float, true>(a, a_cs, a_co, b, b_cs, b_co, out, o_cs, o_co, C, npix, S(stream));
  return launch_slice<__nv_bfloat16, true>(a, a_cs, a_co, b, b_cs, b_co, out, o_cs, o_co, C, npix, S(stream));
}

template <typename T>
__global__ void unshuffle_kernel(const T* __restrict__ g, int g_cs, int g_co, T* __restrict__ out, int o_cs, int o_co,
                                 int N, int H, int W, int Cp, int r) {
  // out[n,h,w, ij*Cp + c] = g[n, h*r+i, w*r+j, c]
  int rr = r * r;
  int64_t total = (int64_t)N * H * W * rr * Cp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c = idx % Cp;
    int ij = (idx / Cp) % rr;
    int64_t p = idx / ((int64_t)Cp * rr);
    int w = p % W;
    int h = (p / W) % H;
    int64_t n = p / ((int64_t)W * H);
    int i = ij / r, j = ij % r;
    int64_t gp = (n * (H * r) + (h * r + i)) * (int64_t)(W * r) + (w * r + j);
    out[p * o_cs + o_co + ij * Cp + c] = g[gp * g_cs + g_co + c];
  }
}

template <typename T>
__global__ void unshuffle_vec_kernel(const T* __restrict__ g, int g_cs, int g_co, T* __restrict__ out, int o_cs, int o_co,
                                     int N, int H, int W, int Cp, int r) {
  // one 16-byte channel vector per thread; consecutive threads walk channels, then (i,j), then pixels
  using V = VecT<T>;
  const int rr = r * r, cv = Cp / V::W;
  const int64_t total = (int64_t)N * H * W * rr * cv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V::W;
    const int ij = (int)((idx / cv) % rr);
    const int64_t p = idx / ((int64_t)cv * rr);
    const int w = (int)(p % W);
    const int h = (int)((p / W) % H);
    const int64_t n = p / ((int64_t)W * H);
    const int i = ij / r, j = ij % r;
    const int64_t gp = (n * (H * r) + (h * r + i)) * (int64_t)(W * r) + (w * r + j);
    *reinterpret_cast<typename V::raw*>(out + p * o_cs + o_co + ij * Cp + c) =
        *reinterpret_cast<const typename V::raw*>(g + gp * g_cs + g_co + c);
  }
}

extern "C" int srb_pixel_unshuffle(srb_ctx* ctx, const void* g, int g_cs, int g_co, void* out, int o_cs, int o_co, int N,
                                   int H, int W, int Cp, int r, int dtype, void* stream) {
  SRB_REQUIRE(ctx && g && out, "srb_pixel_unshuffle: null argument");
  const int vw = dtype == SRB_F32 ? 4 : 8;
  const bool vec = (Cp % vw == 0) && (g_cs % vw == 0) && (g_co % vw == 0) && (o_cs % vw == 0) && (o_co % vw == 0);
  int64_t total = (int64_t)N * H * W * r * r * (vec ? Cp / vw : Cp);
  int blocks = srb_cdiv(total, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == SRB_F32) {
    if (vec) unshuffle_vec_kernel<float><<<blocks, 256, 0, S(stream)>>>((const float*)g, g_cs, g_co, (float*)out, o_cs, o_co, N, H, W, Cp, r);
    else unshuffle_kernel<float><<<blocks, 256, 0, S(stream)>>>((const float*)g, g_cs, g_co, (float*)out, o_cs, o_co, N, H, W, Cp, r);
  } else {
    if (vec) unshuffle_vec_kernel<__nv_bfloat16><<<blocks, 256, 0, S(stream)>>>((const __nv_bfloat16*)g, g_cs, g_co, (__nv_bfloat16*)out, o_cs, o_co, N, H, W, Cp, r);
    else unshuffle_kernel<__nv_bfloat16><<<blocks, 256, 0, S(stream)>>>((const __nv_bfloat16*)g, g_cs, g_co, (__nv_bfloat16*)out, o_cs, o_co, N, H, W, Cp, r);
  }
  SRB_LAUNCH_CHECK();
  return 0;
}

// out = (act > 0) ? g : 0 over channel slices (ReLU backward, reference: autograd of nn.ReLU)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ g, int g_cs, int g_co, const T* __restrict__ act, int a_cs,
                                int a_co, T* __restrict__ out, int o_cs, int o_co, int C, int64_t npix) {
  int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int c = (int)(i % C);
    float a = ld_elem(act + p * a_cs + a_co + c);
    float v = ld_elem(g + p * g_cs + g_co + c);
    st_elem(out + p * o_cs + o_co + c, a > 0.f ? v : 0.f);
  }
}

// bf16, 8 channels (16 bytes) per thread: the RDN dense-block backward masks a 64-channel slice of the
// 576-channel gradient buffer per layer (128 launches per step; the scalar form took 11.5 us each)
__global__ void relu_bwd_vec8_kernel(const __nv_bfloat16* __restrict__ g, int g_cs, int g_co,
                                     const __nv_bfloat16* __restrict__ act, int a_cs, int a_co,
                                     __nv_bfloat16* __restrict__ out, int o_cs, int o_co, int vpp, int64_t nvec) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int64_t p = i / vpp;
  const int c = (int)(i - p * vpp) * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(act + p * a_cs + a_co + c);
  uint4 v = *reinterpret_cast<const uint4*>(g + p * g_cs + g_co + c);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16x2(aw[e]);
    if (!(f.x > 0.f)) vw[e] &= 0xFFFF0000u;
    if (!(f.y > 0.f)) vw[e] &= 0x0000FFFFu;
  }
  *reinterpret_cast<uint4*>(out + p * o_cs + o_co + c) = make_uint4(vw[0], vw[1], vw[2], vw[3]);
}

extern "C" int srb_relu_bwd(srb_ctx* ctx, const void* g, int g_cs, int g_co, const void* act, int a_cs, int a_co,
                            void* out, int o_cs, int o_co, int C, int64_t npix, int dtype, void* stream) {
  SRB_REQUIRE(ctx && g && act && out, "srb_relu_bwd: null argument");
  if (dtype == SRB_BF16 && C % 8 == 0 && !((g_cs | g_co | a_cs | a_co | o_cs | o_co) % 8) &&
      !((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    const int vpp = C / 8;
    const int64_t nvec = npix * vpp;
    if (nvec == 0) return 0;
    relu_bwd_vec8_kernel<<<srb_cdiv(nvec, 256), 256, 0, S(stream)>>>(
        (const __nv_bfloat16*)g, g_cs, g_co, (const __nv_bfloat16*)act, a_cs, a_co, (__nv_bfloat16*)out, o_cs, o_co, vpp, nvec);
    SRB_LAUNCH_CHECK();
    return 0;
  }
  int64_t total = npix * C;
  int blocks = srb_cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == SRB_F32)
    relu_bwd_kernel<float><<<blocks, 256, 0, S(stream)>>>((const float*)g, g_cs, g_co, (const float*)act, a_cs, a_co,
                                                          (float*)out, o_cs, o_co, C, npix);
  else
    relu_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, S(stream)>>>((const __nv_bfloat16*)g, g_cs, g_co,
                                                                  (const __nv_bfloat16*)act, a_cs, a_co,
                                                                  (__nv_bfloat16*)out, o_cs, o_co, C, npix);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per-channel sums
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int cs, int co, int C, int64_t npix, float* __restrict__ out,
                              float alpha, int shuffle) {
  // blockDim = (32 channels, 8 pixel lanes); grid.x over channel groups of 32, grid.y over pixel slabs
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    for (int64_t p = blockIdx.y * (int64_t)blockDim.y + threadIdx.y; p < npix; p += (int64_t)gridDim.y * blockDim.y)
      acc += ld_elem(x + p * cs + co + c);
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    int oc = c;  // x is in (ij, c') channel order when it is an un-shuffled gradient
    if (shuffle > 1) {
      int rr = shuffle * shuffle, Cp = C / rr;
      oc = (c % Cp) * rr + c / Cp;
    }
    atomicAdd(out + oc, s * alpha);
  }
}

int srb_colsum_launch(srb_ctx* ctx, const void* x, int cs, int co, int C, int64_t npix, int dtype, float* out,
                      int accumulate, float alpha, int shuffle, cudaStream_t st) {
  (void)ctx;
  if (!accumulate) SRB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  dim3 block(32, 8);
  int slabs = srb_cdiv(npix, 8 * 64);
  if (slabs > 512) slabs = 512;
  if (slabs < 1) slabs = 1;
  dim3 grid(srb_cdiv(C, 32), slabs);
  if (dtype == SRB_F32)
    colsum_kernel<float><<<grid, block, 0, st>>>((const float*)x, cs, co, C, npix, out, alpha, shuffle);
  else
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, cs, co, C, npix, out, alpha, shuffle);
  SRB_LAUNCH_CHECK();
  return 0;
}

// Many bias gradients in ONE launch (the deferred weight-gradient batch, conv.cu).  A CTA owns one
// 128 KB chunk of one item (flat grid over all chunks, so a 96x96x256 gradient gets 16x the CTAs of a
// 48x48x64 one); a thread reads 8 bf16 channels per 16-byte load, C/8 threads cover a pixel,
// 256/(C/8) pixels per iteration, four loads in flight; partial sums meet in shared memory and leave
// as C atomics per CTA.  (One colsum_kernel launch per bias was 15 x 14.7 us of the RCAN step.)
struct ColsumItem {
  const __nv_bfloat16* x;
  float* out;
  int64_t npix;
  int cs, co, C, shuffle;
  float alpha;
  int first_block, ppc;        // first CTA of this item, pixels per CTA
  int vpp;                     // 16-byte vectors per pixel: ceil(C / 8), a power of two (lanes >= C are padding)
};
constexpr int kColsumMaxItems = 64;
constexpr int kColsumChunkBytes = 128 * 1024;
struct ColsumBatch {
  ColsumItem it[kColsumMaxItems];
  int n;
};

__global__ void __launch_bounds__(256) colsum_batched_kernel(const __grid_constant__ ColsumBatch B) {
  __shared__ float red[256][9];
  int ii = 0;
  while (ii + 1 < B.n && (int)blockIdx.x >= B.it[ii + 1].first_block) ++ii;
  const ColsumItem& it = B.it[ii];
  const int vpp = it.vpp;                    // 16-byte vectors per pixel (power of two, <= 32)
  const int ppi = 256 / vpp;                 // pixels per CTA iteration
  const int v = threadIdx.x % vpp, pl = threadIdx.x / vpp;
  const int64_t p0 = (int64_t)((int)blockIdx.x - it.first_block) * it.ppc;
  const int64_t p1 = p0 + it.ppc < it.npix ? p0 + it.ppc : it.npix;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  const __nv_bfloat16* base = it.x + it.co + v * 8;
#pragma unroll 4
  for (int64_t p = p0 + pl; p < p1; p += ppi) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(base + p * it.cs));
    const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
    a[0] += f0.x; a[1] += f0.y; a[2] += f1.x; a[3] += f1.y; a[4] += f2.x; a[5] += f2.y; a[6] += f3.x; a[7] += f3.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = a[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (c < it.C) {
    const int cv = c >> 3, cj = c & 7;
    float s = 0.f;
    for (int i = 0; i < ppi; ++i) s += red[i * vpp + cv][cj];
    int oc = c;
    if (it.shuffle > 1) {
      const int rr = it.shuffle * it.shuffle, Cp = it.C / rr;
      oc = (c % Cp) * rr + c / Cp;
    }
    atomicAdd(it.out + oc, s * it.alpha);
  }
}

int srb_colsum_batched_ok(const void* x, int cs, int co, int C, int dtype) {
  if (dtype != SRB_BF16) return 0;
  if (C < 1 || C > 256) return 0;
  const int vpp = (C + 7) / 8;               // the padded vector lanes must exist inside the pixel (RGB tensors are 8 wide)
  if ((vpp & (vpp - 1)) || co + vpp * 8 > cs) return 0;
  if ((cs % 8) || (co % 8) || (reinterpret_cast<uintptr_t>(x) & 15)) return 0;
  return 1;
}

// items: parallel arrays; every item must satisfy srb_colsum_batched_ok
int srb_colsum_batched_launch(srb_ctx* ctx, int n, const void* const* xs, const int* cs, const int* co, const int* C,
                              const int64_t* npix, float* const* outs, const int* accumulate, const float* alpha,
                              const int* shuffle, cudaStream_t st) {
  (void)ctx;
  for (int i0 = 0; i0 < n; i0 += kColsumMaxItems) {
    const int m = n - i0 < kColsumMaxItems ? n - i0 : kColsumMaxItems;
    ColsumBatch B;
    int blocks = 0;
    for (int i = 0; i < m; ++i) {
      const int k = i0 + i;
      if (!accumulate[k]) SRB_CHECK_CUDA(cudaMemsetAsync(outs[k], 0, sizeof(float) * C[k], st));
      ColsumItem& it = B.it[i];
      it.x = reinterpret_cast<const __nv_bfloat16*>(xs[k]);
      it.out = outs[k];
      it.npix = npix[k];
      it.cs = cs[k];
      it.co = co[k];
      it.C = C[k];
      it.shuffle = shuffle[k];
      it.alpha = alpha[k];
      it.first_block = blocks;
      it.vpp = (C[k] + 7) / 8;
      it.ppc = kColsumChunkBytes / (it.vpp * 16);   // a multiple of the 256 / vpp pixels of one iteration
      blocks += srb_cdiv(npix[k] > 0 ? npix[k] : 1, it.ppc);
    }
    B.n = m;
    colsum_batched_kernel<<<blocks, 256, 0, st>>>(B);
    SRB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int srb_colsum(srb_ctx* ctx, const void* x, int cs, int co, int C, int64_t npix, int dtype, float* out,
                          int accumulate, void* stream) {
  SRB_REQUIRE(ctx && x && out, "srb_colsum: null argument");
  return srb_colsum_launch(ctx, x, cs, co, C, npix, dtype, out, accumulate, 1.0f, 0, S(stream));
}

// ---------------------------------------------------------------------------------------------
// L1 loss + seed gradient
// ---------------------------------------------------------------------------------------------
__global__ void l1_kernel(const float* __restrict__ sr, const float* __restrict__ hr, int64_t n, float inv_n,
                          float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = sr[i] - hr[i];
    acc += fabsf(d);
    if (grad) grad[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
  }
}

extern "C" int srb_l1_loss(srb_ctx* ctx, const float* sr, const float* hr, int64_t n, float* loss, float* grad,
                           void* stream) {
  SRB_REQUIRE(ctx && sr && hr && loss, "srb_l1_loss: null argument");
  SRB_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), S(stream)));
  int blocks = srb_cdiv(n, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  l1_kernel<<<blocks, 256, 0, S(stream)>>>(sr, hr, n, 1.0f / (float)n, loss, grad);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, amsgrad=False, maximize=False)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            int step_host, const int32_t* __restrict__ step_dev, float gscale) {
  int step = step_dev ? *step_dev : step_host;
  float bc1 = 1.f - powf(b1, (float)step);
  float bc2 = 1.f - powf(b2, (float)step);
  float step_size = lr / bc1;
  float inv_sqrt_bc2 = rsqrtf(bc2);
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    if (wd != 0.f) gi += wd * pi;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi = pi - step_size * (mi / denom);
  };
  // 16-byte lanes over the aligned body (the flat buffers of srb200.trainer.FlatParams are), scalars for the rest
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nthr) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    update(pv.x, gv.x, mv.x, vv.x);
    update(pv.y, gv.y, mv.y, vv.y);
    update(pv.z, gv.z, mv.z, vv.z);
    update(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nthr) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(pi, g[i], mi, vi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

__global__ void inc_kernel(int32_t* c) { *c += 1; }

extern "C" int srb_inc_counter(srb_ctx* ctx, int32_t* counter, void* stream) {
  SRB_REQUIRE(ctx && counter, "srb_inc_counter: null argument");
  inc_kernel<<<1, 1, 0, S(stream)>>>(counter);
  SRB_LAUNCH_CHECK();
  return 0;
}

__global__ void inc64_kernel(long long* c) { *c += 1; }

extern "C" int srb_inc_counter64(srb_ctx* ctx, int64_t* counter, void* stream) {
  SRB_REQUIRE(ctx && counter, "srb_inc_counter64: null argument");
  inc64_kernel<<<1, 1, 0, S(stream)>>>(reinterpret_cast<long long*>(counter));
  SRB_LAUNCH_CHECK();
  return 0;
}

extern "C" int srb_adam_step(srb_ctx* ctx, float* param, const float* grad, float* m, float* v, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step,
                             const int32_t* step_dev, float grad_scale, void* stream) {
  SRB_REQUIRE(ctx && param && grad && m && v, "srb_adam_step: null argument");
  int blocks = srb_cdiv(n, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, S(stream)>>>(param, grad, m, v, n, lr, beta1, beta2, eps, weight_decay, step, step_dev,
                                             grad_scale);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// table-driven re-pack of all conv weights of a model in ONE launch (per optimizer step)
// ---------------------------------------------------------------------------------------------
// 3x3 tensor-core packing through shared memory.  In OIHW the 64 input channels x 9 taps of one output channel are ONE
// contiguous run of 576 floats, and that run is what both packed forms slice: the forward form takes it as (row = co,
// K = ci), the input-gradient form as (K = co, row = ci) with the taps mirrored.  A CTA stages 32 such runs (coalesced
// reads, converted to bf16: 36 KB) and writes 16-byte packed vectors from the staged tile.  The gather form below read one
// float per 32-byte sector: 1 GB of sector traffic per RCAN step for 62 MB of parameters (0.10 ms).
constexpr int kPkOuter = 32;
constexpr int kPkPitch = 64 * 9 + 2;      // halfwords per staged run; +2 spreads the 8-run column reads over the banks
constexpr int kPkSubElems = kPkOuter * 64 * 9;

__device__ __forceinline__ uint16_t bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}

__device__ void pack_umma3_staged(const srb_pack_item& it, uint16_t* T) {
  const bool fwd = it.mode == SRB_PACK_FWD;
  const int Cout = it.Cout, Cin = it.Cin;
  const int rr = it.shuffle > 1 ? it.shuffle * it.shuffle : 1;
  const int Cp = Cout / rr;
  const int cin_e = fwd ? Cin : Cout, cout_e = fwd ? Cout : Cin;     // K extent / rows of the packed matrix
  const int nchunk = (cin_e + 63) / 64;
  const int nrb = fwd ? (cout_e + kPkOuter - 1) / kPkOuter : (cout_e + 63) / 64;
  const int nsub = fwd ? nchunk * nrb : nchunk * 2 * nrb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* out = reinterpret_cast<uint4*>(it.dst);
  for (int sb = blockIdx.x; sb < nsub; sb += gridDim.x) {
    // staged runs: outer = output channel (forward: the packed row; input gradient: the K index), inner = 64 input channels
    int chunk, hf = 0, outer0, inner0;
    if (fwd) {
      chunk = sb / nrb;
      outer0 = (sb % nrb) * kPkOuter;
      inner0 = chunk * 64;
    } else {
      chunk = sb / (2 * nrb);
      const int rem = sb % (2 * nrb);
      hf = rem / nrb;
      outer0 = chunk * 64 + hf * kPkOuter;
      inner0 = (rem % nrb) * 64;
    }
    const int n_inner = min(64, Cin - inner0);
    __syncthreads();      // the previous sub-block's readers are done with T
    {
      // the warp's 4 runs: all 72 loads of a lane are issued before the first conversion / shared-memory store
      float v[kPkOuter / 8][18];
#pragma unroll
      for (int i = 0; i < kPkOuter / 8; ++i) {
        const int oi = outer0 + warp + 8 * i;
        const bool valid = oi < Cout;
        const int co = !valid ? 0 : (rr > 1 ? (oi % Cp) * rr + oi / Cp : oi);
        const float* run = it.src + ((int64_t)co * Cin + inner0) * 9;
        const int len = valid ? n_inner * 9 : 0;
#pragma unroll
        for (int k = 0; k < 18; ++k) v[i][k] = lane + 32 * k < len ? __ldg(run + lane + 32 * k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kPkOuter / 8; ++i)
#pragma unroll
        for (int k = 0; k < 18; ++k) T[(warp + 8 * i) * kPkPitch + lane + 32 * k] = bf16_bits(v[i][k]);
    }
    __syncthreads();
    if (fwd) {
      for (int u = threadIdx.x; u < 9 * kPkOuter * 8; u += blockDim.x) {
        const int j = u & 7, o = (u >> 3) & (kPkOuter - 1), slot = u >> 8;      // slot = kw * 3 + kh of the packed buffer
        const int row = outer0 + o;
        if (row >= cout_e) continue;
        const int kw = slot / 3, kh = slot % 3;
        const uint16_t* t = T + o * kPkPitch + (j * 8) * 9 + kh * 3 + kw;
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (uint32_t)t[(2 * q) * 9] | ((uint32_t)t[(2 * q + 1) * 9] << 16);
        out[(((int64_t)(chunk * 3 + kw) * 3 + kh) * cout_e + row) * 8 + j] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      for (int u = threadIdx.x; u < 9 * 64 * 4; u += blockDim.x) {
        const int j4 = u & 3, r = (u >> 2) & 63, slot = u >> 8;
        const int row = inner0 + r;
        if (row >= cout_e) continue;
        const int kw = slot / 3, kh = slot % 3;
        const uint16_t* t = T + (j4 * 8) * kPkPitch + r * 9 + (2 - kh) * 3 + (2 - kw);      // mirrored taps
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (uint32_t)t[(2 * q) * kPkPitch] | ((uint32_t)t[(2 * q + 1) * kPkPitch] << 16);
        out[(((int64_t)(chunk * 3 + kw) * 3 + kh) * cout_e + row) * 8 + hf * 4 + j4] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) pack_table_kernel(const srb_pack_item* __restrict__ table, int staged) {
  __shared__ uint16_t stage[kPkOuter * kPkPitch];
  const srb_pack_item it = table[blockIdx.y];
  const int k = it.ksize;
  if (k == 0) {  // bias
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < it.Cout; co += gridDim.x * blockDim.x)
      reinterpret_cast<float*>(it.dst)[perm_channel(co, it.Cout, it.shuffle)] = it.src[co];
    return;
  }
  const int Cout = it.Cout, Cin = it.Cin;
  const int rr = it.shuffle > 1 ? it.shuffle * it.shuffle : 1;
  if (it.packing == SRB_PACK_UMMA && k == 3 && staged) {
    pack_umma3_staged(it, stage);
    return;
  }
  if (it.packing == SRB_PACK_UMMA) {
    // Output-centric: a thread produces 8 consecutive bf16 of the packed buffer (one 16-byte store,
    // fully coalesced) and gathers its 8 sources — strided fp32 reads that the nine taps / 64 channels
    // of a block share through L1/L2.  (The source-centric form wrote 2 bytes per 32-byte sector:
    // 229 us per RCAN step for 31 M elements.)
    const int cin_e = it.mode == SRB_PACK_FWD ? Cin : Cout;
    const int cout_e = it.mode == SRB_PACK_FWD ? Cout : Cin;
    const int nchunk = (cin_e + 63) / 64;
    const int Cp = Cout / rr;
    const int64_t nvec = (int64_t)nchunk * k * k * cout_e * 8;
    uint4* out = reinterpret_cast<uint4*>(it.dst);
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
      const int kk0 = (int)(v % 8) * 8;
      const int row = (int)((v / 8) % cout_e);
      const int kh = (int)((v / ((int64_t)8 * cout_e)) % k);
      const int kw = (int)((v / ((int64_t)8 * cout_e * k)) % k);
      const int chunk = (int)(v / ((int64_t)8 * cout_e * k * k));
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cin_idx = chunk * 64 + kk0 + j;
        float x = 0.f;
        if (cin_idx < cin_e) {
          if (it.mode == SRB_PACK_FWD) {
            int co = row;
            if (rr > 1) co = (row % Cp) * rr + row / Cp;
            x = it.src[(((int64_t)co * Cin + cin_idx) * k + kh) * k + kw];
          } else {
            int co = cin_idx;
            if (rr > 1) co = (cin_idx % Cp) * rr + cin_idx / Cp;
            x = it.src[(((int64_t)co * Cin + row) * k + (k - 1 - kh)) * k + (k - 1 - kw)];
          }
        }
        f[j] = x;
      }
      out[v] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    return;
  }
  const int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kw = i % k;
    const int kh = (i / k) % k;
    const int ci = (i / (k * k)) % Cin;
    const int co = i / ((int64_t)k * k * Cin);
    const int cop = perm_channel(co, Cout, it.shuffle);
    const float v = it.src[i];
    float* out = reinterpret_cast<float*>(it.dst);
    if (it.mode == SRB_PACK_FWD) out[(((int64_t)kh * k + kw) * Cin + ci) * Cout + cop] = v;
    else out[(((int64_t)(k - 1 - kh) * k + (k - 1 - kw)) * Cout + cop) * Cin + ci] = v;
  }
}

extern "C" int srb_pack_table(srb_ctx* ctx, const srb_pack_item* table_dev, int n, int64_t max_elems, void* stream) {
  SRB_REQUIRE(ctx && (table_dev || n == 0), "srb_pack_table: null argument");
  ctx->weights_dirty = 1;
  if (n == 0) return 0;
  SRB_REQUIRE(n <= 65535, "srb_pack_table: too many items (%d)", n);
  // grid.x is sized for a TYPICAL item (the caller passes the median filter size): every path of the kernel is a
  // grid-stride loop, so larger filters take more trips; one 256-thread CTA per 2048 elements = one 16-byte packed
  // vector per thread.  Sizing it for the largest filter launched 105 k CTAs per RCAN step, most without work.
  // (A 3x3 tensor-core item takes one CTA per 32 x 64 x 9 elements: pack_umma3_staged; its surplus CTAs exit at once.)
  int bx = srb_cdiv(max_elems, 256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  static const int staged = [] {      // SRB200_PACK_STAGED=0: the gather form for every item (A/B runs)
    const char* e = getenv("SRB200_PACK_STAGED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  pack_table_kernel<<<dim3(bx, n), 256, 0, S(stream)>>>(table_dev, staged);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---- stream delay ---------------------------------------------------------------------------------------------------
// One thread that waits `ns` nanoseconds.  Used at the head of the weight-gradient side stream (srb200.ops.WgradOverlap):
// the batched weight-gradient launch and the next backward chain both become runnable when the previous chain ends, and
// the chain's cluster launch reaches the SMs ~1.5 us later; if the weight-gradient CTAs are placed first they are spread
// over all GPCs and the 16 six-SM clusters no longer fit at once (chain launch 345 -> 580 us).
__global__ void delay_kernel(unsigned long long ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (true) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns) break;
    __nanosleep(100);
  }
}

extern "C" int srb_delay(srb_ctx* ctx, int64_t ns, void* stream) {
  SRB_REQUIRE(ctx && ns >= 0 && ns <= 1000000, "srb_delay: 0 <= ns <= 1e6");
  if (ns == 0) return 0;
  delay_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>((unsigned long long)ns);
  SRB_LAUNCH_CHECK();
  return 0;
}
