// Diagnostics: tcgen05.mma issue/throughput probe (not on the product path).  One CTA per SM
// issues `iters` back-to-back MMAs of shape 128 x N x 16 (bf16, SS operands, 128-B swizzle) from a
// resident shared-memory tile and reports SM cycles per MMA — the measured tensor-pipe ceiling for
// the tile shapes the conv kernels use (DESIGN.md "Kernels").
#include "common.cuh"
#include "ptx.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int N, int iters, int mn_major, int distinct,
                                                            long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  // zero the operand area so no NaN/denormal effects
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - ptx::smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&slot, 256);
    ptx::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lo = ptx::smem_desc_lo(base, mn_major ? 1024u : 16u);
    const uint32_t b_lo = ptx::smem_desc_lo(base + 48 * 1024, mn_major ? 1024u : 16u);
    constexpr uint32_t hi = ptx::smem_desc_hi_sw128(1024u);
    const uint32_t step = distinct > 1 ? (mn_major ? 128u : 2u) : 0u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) ptx::umma_bf16_lohi(tmem, a_lo + u * step, hi, b_lo + u * step, hi, idesc, 1u);
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

}  // namespace

/* cycles[blocks] <- SM cycles for `iters` MMAs on each of `blocks` CTAs (one per SM). */
extern "C" int srb_probe_umma(srb_ctx* ctx, int N, int iters, int mn_major, int distinct, int blocks,
                              long long* cycles_dev, void* stream) {
  SRB_REQUIRE(ctx && cycles_dev, "srb_probe_umma: null argument");
  SRB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && distinct >= 1 && distinct <= 4, "srb_probe_umma: bad arguments");
  const size_t smem = 97 * 1024 + 1024;
  static bool attr = false;
  if (!attr) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  umma_probe_kernel<<<blocks, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(N, iters, mn_major, distinct, cycles_dev);
  SRB_LAUNCH_CHECK();
  return 0;
}
