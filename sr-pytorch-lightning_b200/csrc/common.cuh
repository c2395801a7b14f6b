// Shared helpers for libsrb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/srb200.h"

struct srb_ctx {
  int device;
  int num_sms;
  int smem_optin;
  void* encode_tiled;  // PFN_cuTensorMapEncodeTiled, resolved at srb_create
  int wgrad_sm_budget; // > 0: batched weight-gradient launches use at most this many CTAs (srb_set_wgrad_sm_budget)
  int weights_dirty;   // a pack kernel was launched since the last fully serialised conv launch (conv_c64.cu)
  long long* trace;    // diagnostics: device buffer for per-CTA event clocks of conv_c64 (NULL = off)
  int no_pdl;          // SRB200_NO_PDL=1: never use programmatic dependent launch (A/B measurements)
};

void srb_set_error(const char* fmt, ...);

#define SRB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      srb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define SRB_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      srb_set_error(__VA_ARGS__); \
      return 2;                   \
    }                             \
  } while (0)

extern unsigned long long g_srb_launches;  // kernels launched by this library (bench.py's gpu_launches)
#define SRB_LAUNCH_CHECK()                   \
  do {                                       \
    __atomic_fetch_add(&g_srb_launches, 1ull, __ATOMIC_RELAXED); \
    SRB_CHECK_CUDA(cudaGetLastError());      \
  } while (0)

static inline int srb_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- element access -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ld_elem(const T* p);
template <> __device__ __forceinline__ float ld_elem<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_elem<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_elem(T* p, float v);
template <> __device__ __forceinline__ void st_elem<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 4 consecutive channels (16-B for fp32, 8-B for bf16); pointer must be aligned accordingly
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(a);
}

// 16-byte channel vectors: 4 fp32 or 8 bf16
template <typename T> struct VecT;
template <> struct VecT<float> {
  static constexpr int W = 4;
  typedef float4 raw;
  __device__ static void unpack(const raw& r, float (&f)[4]) { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
  __device__ static raw pack(const float (&f)[4]) { return make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct VecT<__nv_bfloat16> {
  static constexpr int W = 8;
  typedef uint4 raw;
  __device__ static void unpack(const raw& r, float (&f)[8]) {
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  __device__ static raw pack(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
