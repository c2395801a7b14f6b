// Hardware probes for design decisions (diagnostics, not on the product path).  Standalone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../sr-pytorch-lightning_b200/csrc probes.cu -o probes
// P1  UMMA row-shift: can a K-major SWIZZLE_128B operand start at an arbitrary 128-B row of a
//     TMA-written window (pitch P pixels), i.e. one input window serving all 9 taps of a 3x3 conv?
// P3  L2 -> SM TMA bandwidth with every SM pulling at once (distinct vs identical addresses).
// P4  launch floor of a dependent kernel chain inside a CUDA graph, with and without programmatic
//     dependent launch (PDL).
// P5  cluster co-residency (cudaOccupancyMaxActiveClusters).
// P6  inter-CTA signalling latency through L2 (release/acquire flag ping-pong, N-CTA gather).
// P7  distributed shared memory inside a cluster: remote store + remote mbarrier arrive ping-pong, halo exchange
//     between ring neighbours, 64-float all-gather (round 2: per-sample cluster chains).
//     Written in round 1 after the GPU budget was spent: compiles for sm_100a, NOT yet run.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ptx.cuh"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static void make_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(1);
  }
}

// ------------------------------------------------------------------------------------------------
// P1: row-shifted UMMA operand
// ------------------------------------------------------------------------------------------------
// window = ROWS x P pixels x 64 ch (bf16) loaded by ONE TMA (SW128) at a 1024-aligned base.
// For each tap (kh,kw): D[r][n] = sum_k A[r][k] * I[n][k], A row r = window pixel ((r/8)+kh)*P + (r%8)+kw.
// variant 0: base_offset field 0;  variant 1: base_offset = (start >> 7) & 7.
__global__ void __launch_bounds__(128, 1) p1_kernel(const __grid_constant__ CUtensorMap tm, int P, int rows, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t win_bytes = (uint32_t)rows * P * 128u;
  const uint32_t b_off = (win_bytes + 1023u) & ~1023u;
  // identity B: 64 rows (n) x 64 k, K-major SW128
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    const uint32_t off = n * 128 + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(gen + b_off + off) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_load, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&slot, 64);
    ptx::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_load, win_bytes);
    ptx::tma_load_3d(base, &tm, &bar_load, 0, 0, 0);
    ptx::mbar_wait(&bar_load, 0);
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int tap = 0; tap < 9; ++tap) {
      const int kh = tap / 3, kw = tap % 3;
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
        const uint32_t start = base + (uint32_t)(kh * P + kw) * 128u;
        const uint32_t a_lo = ptx::smem_desc_lo(start, 16u);
        uint32_t a_hi = ptx::smem_desc_hi_sw128((uint32_t)P * 128u);
        if (variant == 1) a_hi |= ((start >> 7) & 7u) << (49 - 32);
        const uint32_t b_lo = ptx::smem_desc_lo(base + b_off, 16u);
        const uint32_t b_hi = ptx::smem_desc_hi_sw128(1024u);
        const uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 0, 0);
        for (int k = 0; k < 4; ++k) ptx::umma_bf16_lohi(tmem, a_lo + k * 2u, a_hi, b_lo + k * 2u, b_hi, idesc, k != 0);
        ptx::umma_commit(&bar_mma);
      }
      ptx::mbar_wait(&bar_mma, phase);
      phase ^= 1u;
      ptx::tc_fence_after();
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      float* o = out + ((size_t)(variant * 9 + tap) * 128 + warp * 32 + lane) * 64;
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t acc[32];
        ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) o[c0 + j] = __uint_as_float(acc[j]);
      }
      ptx::tc_fence_before();
      __syncthreads();
    }
  }
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 64);
  }
}

static void run_p1(int P) {
  const int rows = 18, npx = rows * P;
  std::vector<__nv_bfloat16> h(npx * 64);
  auto val = [](int q, int c) { return (float)((q * 7 + c * 3) % 255 - 127); };
  for (int q = 0; q < npx; ++q)
    for (int c = 0; c < 64; ++c) h[q * 64 + c] = __float2bfloat16(val(q, c));
  __nv_bfloat16* dx;
  float* dout;
  CK(cudaMalloc(&dx, h.size() * 2));
  CK(cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const size_t out_n = 2 * 9 * 128 * 64;
  CK(cudaMalloc(&dout, out_n * 4));
  CK(cudaMemset(dout, 0xff, out_n * 4));
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, (cuuint64_t)P, (cuuint64_t)rows};
  cuuint64_t strides[2] = {128, (cuuint64_t)P * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)P, (cuuint32_t)rows};
  make_map(&tm, dx, 3, dims, strides, box);
  const size_t smem = (size_t)rows * P * 128 + 1024 + 8192 + 1024;
  CK(cudaFuncSetAttribute(p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p1_kernel<<<1, 128, smem>>>(tm, P, rows, dout);
  CK(cudaDeviceSynchronize());
  std::vector<float> o(out_n);
  CK(cudaMemcpy(o.data(), dout, out_n * 4, cudaMemcpyDeviceToHost));
  for (int variant = 0; variant < 2; ++variant) {
    printf("P1 pitch=%d base_offset=%s mismatches per tap:", P, variant ? "(start>>7)&7" : "0");
    for (int tap = 0; tap < 9; ++tap) {
      int bad = 0;
      const int kh = tap / 3, kw = tap % 3;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) {
          const int q = ((r / 8) + kh) * P + (r % 8) + kw;
          if (o[((size_t)(variant * 9 + tap) * 128 + r) * 64 + c] != val(q, c)) ++bad;
        }
      printf(" %d", bad);
    }
    printf("\n");
  }
  CK(cudaFree(dx));
  CK(cudaFree(dout));
}

// ------------------------------------------------------------------------------------------------
// P3: L2 -> SM TMA bandwidth
// ------------------------------------------------------------------------------------------------
// Each CTA streams `nbox` boxes of (64 ch x 8 px x rows) from a [npix_rows][8][64] tensor through a
// 4-stage ring; mode 0: CTA-distinct boxes; mode 1: all CTAs load the same boxes (weight-like).
__global__ void __launch_bounds__(64, 1) p3_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int nbox, int total_rows,
                                                   int mode, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[4], empty[4];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bytes = (uint32_t)box_rows * 8u * 128u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nbox; ++i) {
      const int s = i & 3;
      ptx::mbar_wait(&empty[s], ((i >> 2) & 1) ^ 1);
      ptx::mbar_arrive_expect_tx(&full[s], bytes);
      int row = mode == 1 ? (i * box_rows) % (total_rows - box_rows) : (int)(((long long)blockIdx.x * nbox + i) * box_rows % (total_rows - box_rows));
      ptx::tma_load_3d(base + s * bytes, &tm, &full[s], 0, 0, row);
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < nbox; ++i) {
      const int s = i & 3;
      ptx::mbar_wait(&full[s], (i >> 2) & 1);
      ptx::mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

static void run_p3(int num_sms) {
  const int total_rows = 16 * 48 * 48 / 8;  // 4.7 MB tensor viewed as [rows][8 px][64 ch]
  __nv_bfloat16* dx;
  CK(cudaMalloc(&dx, (size_t)total_rows * 8 * 128));
  CK(cudaMemset(dx, 0, (size_t)total_rows * 8 * 128));
  long long* dcy;
  CK(cudaMalloc(&dcy, 256 * 8));
  for (int box_rows : {18, 36}) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {64, 8, (cuuint64_t)total_rows};
    cuuint64_t strides[2] = {128, 1024};
    cuuint32_t box[3] = {64, 8, (cuuint32_t)box_rows};
    make_map(&tm, dx, 3, dims, strides, box);
    const size_t smem = 4 * (size_t)box_rows * 1024 + 2048;
    CK(cudaFuncSetAttribute(p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int mode = 0; mode < 2; ++mode) {
      for (int blocks : {1, 16, num_sms}) {
        const int nbox = box_rows == 18 ? 64 : 32;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        p3_kernel<<<blocks, 64, smem>>>(tm, box_rows, nbox, total_rows, mode, dcy);  // warm (L2 fill)
        CK(cudaEventRecord(e0));
        p3_kernel<<<blocks, 64, smem>>>(tm, box_rows, nbox, total_rows, mode, dcy);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> cy(blocks);
        CK(cudaMemcpy(cy.data(), dcy, blocks * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto c : cy) avg += (double)c;
        avg /= blocks;
        const double bytes_cta = (double)nbox * box_rows * 1024;
        printf("P3 box=%2d KB mode=%s blocks=%3d: %.1f B/cyc/SM (in-kernel), kernel %.2f us -> %.2f TB/s aggregate\n", box_rows,
               mode ? "same-addr" : "distinct ", blocks, bytes_cta / avg, ms * 1e3, bytes_cta * blocks / (ms * 1e-3) / 1e12);
      }
    }
  }
  CK(cudaFree(dx));
  CK(cudaFree(dcy));
}

// ------------------------------------------------------------------------------------------------
// P4: dependent-chain launch floor in a CUDA graph, with / without PDL
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1) p4_kernel(float* buf, int pdl, int work_iters) {
  extern __shared__ uint8_t smem_raw[];
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // "prologue" work that does not depend on the previous kernel
  if (threadIdx.x == 0) smem_raw[0] = 1;
  __syncthreads();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = buf[i];
  for (int k = 0; k < work_iters; ++k) v = v * 1.0001f + 0.5f;
  buf[i] = v + smem_raw[0];
}

static void run_p4(int num_sms) {
  float* buf;
  CK(cudaMalloc(&buf, 256 * 192 * 4));
  CK(cudaMemset(buf, 0, 256 * 192 * 4));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const int chain = 400;
  for (size_t smem : {(size_t)100 * 1024, (size_t)210 * 1024}) {
    CK(cudaFuncSetAttribute(p4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int work : {0, 2000}) {
      for (int pdl = 0; pdl < 2; ++pdl) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < chain; ++i) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(num_sms);
          cfg.blockDim = dim3(192);
          cfg.dynamicSmemBytes = smem;
          cfg.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at;
          cfg.numAttrs = pdl ? 1 : 0;
          CK(cudaLaunchKernelEx(&cfg, p4_kernel, buf, pdl, work));
        }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaStreamSynchronize(st));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, st));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("P4 graph chain smem=%3zu KB work=%4d pdl=%d: %.2f us per kernel\n", smem / 1024, work, pdl, ms * 1e3 / chain);
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
      }
    }
  }
  float h;
  CK(cudaMemcpy(&h, buf, 4, cudaMemcpyDeviceToHost));
  printf("P4 check value %.3f\n", h);
  CK(cudaFree(buf));
}

// ------------------------------------------------------------------------------------------------
// P5: cluster co-residency
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1) p5_kernel(int* x) {
  extern __shared__ uint8_t smem_raw[];
  if (x && threadIdx.x == 0) smem_raw[0] = (uint8_t)x[0];
}
static void run_p5() {
  CK(cudaFuncSetAttribute(p5_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (size_t smem : {(size_t)100 * 1024, (size_t)200 * 1024}) {
    CK(cudaFuncSetAttribute(p5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int cs : {1, 2, 3, 4, 6, 8, 9, 12, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 32);
      cfg.blockDim = dim3(192);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, p5_kernel, &cfg);
      printf("P5 smem=%3zu KB cluster=%2d: max active clusters %d (%d CTAs)%s\n", smem / 1024, cs, n, n * cs,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// P6: inter-CTA signalling latency through L2
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ping-pong between CTA 0 and CTA `peer` (all other CTAs idle)
__global__ void p6_pingpong(int* flags, int iters, int peer, long long* cycles) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      st_release(flags, i);
      while (ld_acquire(flags + 32) < i) {}
    }
    cycles[0] = clock64() - t0;
  } else if ((int)blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) {
      while (ld_acquire(flags) < i) {}
      st_release(flags + 32, i);
    }
  }
}
// gather: groups of `gsz` CTAs; each round every CTA writes 16 KB of data, fences, adds 1 to the
// group counter and waits until the counter reaches gsz * round  (the per-sample CA sync pattern)
__global__ void p6_gather(int* counters, float* data, int gsz, int iters, long long* cycles) {
  const int grp = blockIdx.x / gsz;
  int* ctr = counters + grp * 32;
  float* mine = data + (size_t)blockIdx.x * 4096;
  const long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) mine[k] = (float)i;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      red_release(ctr, 1);
      while (ld_acquire(ctr) < gsz * i) {}
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

static void run_p6(int num_sms) {
  int* flags;
  float* data;
  long long* dcy;
  CK(cudaMalloc(&flags, 4096 * 4));
  CK(cudaMalloc(&data, (size_t)256 * 4096 * 4));
  CK(cudaMalloc(&dcy, 256 * 8));
  const int iters = 2000;
  for (int peer : {1, 2, 73, 147}) {
    CK(cudaMemset(flags, 0, 4096 * 4));
    p6_pingpong<<<num_sms, 32>>>(flags, iters, peer, dcy);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, dcy, 8, cudaMemcpyDeviceToHost));
    printf("P6 ping-pong CTA0<->CTA%d: %.0f cycles one-way (st.release -> ld.acquire observed)\n", peer, (double)c / iters / 2);
  }
  for (int gsz : {1, 9, 18, 144}) {
    CK(cudaMemset(flags, 0, 4096 * 4));
    const int blocks = 144;
    p6_gather<<<blocks, 128>>>(flags, data, gsz, 200, dcy);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cy(blocks);
    CK(cudaMemcpy(cy.data(), dcy, blocks * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto c : cy) avg += (double)c;
    printf("P6 gather group=%3d (16 KB store + fence + red.release + acquire-poll): %.0f cycles per round\n", gsz, avg / blocks / 200);
  }
  CK(cudaFree(flags));
  CK(cudaFree(data));
  CK(cudaFree(dcy));
}

// ------------------------------------------------------------------------------------------------
// P7: distributed shared memory inside a thread-block cluster (the halo exchange / CALayer all-gather of the
//     per-sample cluster design)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = ptx::smem_u32(bar);
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) {      // ~2 s: a protocol bug must not hang the GPU
      printf("P7: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode 0: ping-pong of an 8-byte token between rank 0 and rank `peer` (remote store + remote mbarrier arrive)
// mode 1: halo exchange — every CTA writes `bytes` to its two ring neighbours' halo buffers (128 threads, 16-byte remote
//         stores), each thread then arrives on the neighbour's barrier (count 256 = 2 writers x 128 threads); a round ends
//         when the CTA's own barrier completes
// mode 2: all-gather of 64 floats: every CTA writes its 256 bytes into slot[rank] of every CTA (64 threads), arrives once
//         per peer (thread 0 after a CTA barrier + fence), waits for `csize` arrivals
__global__ void __launch_bounds__(128, 1) p7_kernel(int mode, int peer, int bytes, int iters, long long* cycles) {
  extern __shared__ __align__(16) uint8_t smem[];
  // two barriers, used alternately: a neighbour can be one round ahead (its arrivals for round i+1 may land before this
  // CTA has waited for round i), but never two — so a barrier is never completed twice unobserved
  __shared__ uint64_t bar[2];
  const uint32_t rank = cluster_ctarank();
  uint32_t csize;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const uint32_t buf = ptx::smem_u32(smem);                 // [0, 32 KB): halo-from-left | +32 KB: halo-from-right
  const uint32_t bar_a = ptx::smem_u32(&bar[0]);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) ptx::mbar_init(&bar[b], mode == 0 ? 1 : (mode == 1 ? 256 : csize));
    ptx::fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();
  const long long t0 = clock64();
  if (mode == 0) {
    if (threadIdx.x == 0 && (rank == 0 || (int)rank == peer)) {
      const uint32_t other = rank == 0 ? (uint32_t)peer : 0u;
      const uint32_t rbuf = mapa(buf, other), rbar = mapa(bar_a, other);
      for (int i = 0; i < iters; ++i) {
        const int b = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        if (rank == 0) {
          st_cluster_v4(rbuf, make_uint4(i, i, i, i));
          mbar_arrive_remote_release(rbar + 8u * b);
          mbar_wait_cluster(&bar[b], par);
        } else {
          mbar_wait_cluster(&bar[b], par);
          st_cluster_v4(rbuf, make_uint4(i, i, i, i));
          mbar_arrive_remote_release(rbar + 8u * b);
        }
      }
    }
  } else if (mode == 1) {
    const uint32_t left = (rank + csize - 1) % csize, right = (rank + 1) % csize;
    const uint32_t l_buf = mapa(buf + 32768u, left), r_buf = mapa(buf, right);      // I am the right / left neighbour there
    const uint32_t l_bar = mapa(bar_a, left), r_bar = mapa(bar_a, right);
    for (int i = 0; i < iters; ++i) {
      for (int o = threadIdx.x * 16; o < bytes; o += 128 * 16) {
        const uint4 v = make_uint4(i, o, rank, 0);
        st_cluster_v4(l_buf + o, v);
        st_cluster_v4(r_buf + o, v);
      }
      const int b = i & 1;
      mbar_arrive_remote_release(l_bar + 8u * b);
      mbar_arrive_remote_release(r_bar + 8u * b);
      mbar_wait_cluster(&bar[b], (uint32_t)(i >> 1) & 1u);
      __syncthreads();     // (a real kernel double-buffers the halo rows as well; the probe only times the protocol)
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      if (threadIdx.x < 16) {
        for (uint32_t pr = 0; pr < csize; ++pr)
          st_cluster_v4(mapa(buf + rank * 256u + threadIdx.x * 16u, pr), make_uint4(i, rank, threadIdx.x, 0));
      }
      __syncthreads();
      const int b = i & 1;
      if (threadIdx.x == 0)
        for (uint32_t pr = 0; pr < csize; ++pr) mbar_arrive_remote_release(mapa(bar_a, pr) + 8u * b);
      mbar_wait_cluster(&bar[b], (uint32_t)(i >> 1) & 1u);
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  __syncthreads();
  cluster_sync_all();      // no CTA exits while a peer may still write into its shared memory
}

static void run_p7() {
  long long* dcy;
  CK(cudaMalloc(&dcy, 256 * 8));
  const int smem = 65536;
  CK(cudaFuncSetAttribute(p7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(p7_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  auto launch = [&](int csize, int nclusters, int mode, int peer, int bytes, int iters, const char* what) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize * nclusters);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaMemset(dcy, 0, 256 * 8));
    CK(cudaLaunchKernelEx(&cfg, p7_kernel, mode, peer, bytes, iters, dcy));
    CK(cudaDeviceSynchronize());
    std::vector<long long> cy(csize * nclusters);
    CK(cudaMemcpy(cy.data(), dcy, cy.size() * 8, cudaMemcpyDeviceToHost));
    double mx = 0;
    for (auto c : cy) mx = c > mx ? (double)c : mx;
    printf("P7 cluster=%d x%2d %s: %.0f cycles per %s\n", csize, nclusters, what, mx / iters / (mode == 0 ? 2 : 1),
           mode == 0 ? "one-way hop" : "round");
  };
  const int iters = 2000;
  for (int csize : {2, 6}) {
    launch(csize, 1, 0, 1, 0, iters, "DSMEM ping-pong rank0<->rank1 (16-byte remote store + remote mbarrier arrive)");
    if (csize > 2) launch(csize, 1, 0, csize - 1, 0, iters, "DSMEM ping-pong rank0<->last rank");
  }
  for (int bytes : {2048, 6144, 16384})
    for (int ncl : {1, 16}) {
      char what[128];
      snprintf(what, sizeof what, "halo exchange, %5d B to each of 2 neighbours", bytes);
      launch(6, ncl, 1, 0, bytes, iters, what);
    }
  for (int ncl : {1, 16}) launch(6, ncl, 2, 0, 0, iters, "all-gather of 64 floats among the 6 CTAs");
  CK(cudaFree(dcy));
}


// ------------------------------------------------------------------------------------------------
// P8: tcgen05.mma.cta_group::2 throughput at N = 64 / 128 (M = 256 over a CTA pair).  In SS mode a 128 x 64 x 16 MMA
//     reads 4 KB of A + 2 KB of B from shared memory (48 cycles at 128 B/cycle against 32 of tensor work); a CTA pair
//     shares B (each CTA holds N/2 rows of it), i.e. 5 KB per CTA and MMA -> expected 40 cycles.
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) p8_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - ptx::smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  long long t0 = 0;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t adesc = ptx::smem_desc_sw128(base, 16u, 1024u);
    const uint64_t bdesc = ptx::smem_desc_sw128(base + 48 * 1024, 16u, 1024u);
    t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(adesc + (uint64_t)(u * 2)), "l"(bdesc + (uint64_t)(u * 2)), "r"(idesc), "r"(1u)
            : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     ptx::smem_u32(&bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  if (threadIdx.x == 0) {
    ptx::mbar_wait(&bar, 0);
    if (rank == 0) out[blockIdx.x / 2] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

static void run_p8(int num_sms) {
  long long* dcy;
  CK(cudaMalloc(&dcy, 256 * 8));
  const int smem = 98 * 1024;
  CK(cudaFuncSetAttribute(p8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4096;
  for (int pairs : {1, num_sms / 2})
    for (int N : {64, 128, 256}) {
      CK(cudaMemset(dcy, 0, 256 * 8));
      p8_kernel<<<pairs * 2, 128, smem>>>(N, iters, dcy);
      CK(cudaDeviceSynchronize());
      std::vector<long long> cy(pairs);
      CK(cudaMemcpy(cy.data(), dcy, pairs * 8, cudaMemcpyDeviceToHost));
      double mean = 0;
      for (auto c : cy) mean += (double)c / pairs;
      const double cyc = mean / iters, macs = 128.0 * N * 16;   // per SM
      printf("P8 cta_group::2 pairs=%3d M=256 N=%3d: %6.1f cyc/MMA -> %6.0f MAC/cyc/SM (%5.1f%% of 4096)\n", pairs, N, cyc,
             macs / cyc, 100.0 * macs / cyc / 4096.0);
    }
  CK(cudaFree(dcy));
}

// ------------------------------------------------------------------------------------------------
// P9: does a K-major SWIZZLE_128B A operand whose 8-row groups do NOT start on a 1024-byte swizzle atom (the row-shifted
//     window taps of the 3x3 convs: start = (kh * pitch + kw) * 128 B, stride between 8-row groups = pitch * 128 B) cost
//     tensor throughput?  128 x 64 x 16 MMAs back to back, A start offset / group stride varied, B fixed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) p9_kernel(int a_off_bytes, int a_sbo_bytes, int iters, int vary, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x)
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
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(128, 64, 0, 0);
    const uint32_t a_lo = ptx::smem_desc_lo(base + (uint32_t)a_off_bytes, 16u);
    const uint32_t b_lo = ptx::smem_desc_lo(base + 128 * 1024, 16u);
    const uint32_t hi_a = ptx::smem_desc_hi_sw128((uint32_t)a_sbo_bytes);
    constexpr uint32_t hi_b = ptx::smem_desc_hi_sw128(1024u);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)      // vary: walk the four 32-byte k-steps and three row shifts like a conv tap loop
        ptx::umma_bf16_lohi(tmem, a_lo + (vary ? (uint32_t)(u * 2 + ((i >> 2) % 3) * 8) : 0u), hi_a, b_lo + (vary ? (uint32_t)(u * 2) : 0u), hi_b,
                            idesc, 1u);
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

static void run_p9(int num_sms) {
  long long* dcy;
  CK(cudaMalloc(&dcy, 256 * 8));
  const int smem = 162 * 1024;
  CK(cudaFuncSetAttribute(p9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4096;
  struct Case { int off, sbo, vary; const char* what; };
  const Case cases[] = {
      {0, 1024, 0, "aligned start, groups 1024 B apart (dense tile)"},
      {128, 1024, 0, "start +128 B, groups 1024 B apart"},
      {0, 1280, 0, "aligned start, groups 1280 B apart (pitch 10)"},
      {128, 1280, 0, "start +128 B, pitch 10"},
      {0, 3328, 0, "aligned start, groups 3328 B apart (pitch 26)"},
      {128, 3328, 0, "start +128 B, pitch 26"},
      {3328 + 256, 3328, 0, "start (1 row + 2 px), pitch 26"},
      {0, 3328, 1, "pitch 26, k-steps and kw shifts walked as in the conv"},
      {0, 1280, 1, "pitch 10, k-steps and kw shifts walked as in the conv"},
      {0, 1024, 1, "dense tile, k-steps walked"},
  };
  for (const Case& c : cases) {
    CK(cudaMemset(dcy, 0, 256 * 8));
    p9_kernel<<<num_sms, 128, smem>>>(c.off, c.sbo, iters, c.vary, dcy);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cy(num_sms);
    CK(cudaMemcpy(cy.data(), dcy, num_sms * 8, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (auto v : cy) mean += (double)v / num_sms;
    printf("P9 %-62s: %6.1f cyc/MMA (128x64x16)\n", c.what, mean / iters);
  }
  CK(cudaFree(dcy));
}

// ------------------------------------------------------------------------------------------------
// P10: is the 128-byte swizzle of a TMA tensor STORE a function of the absolute shared-memory address (as it is for UMMA
//      operands, P1) or of the offset inside the box?  A buffer of 128-byte pixel rows is written with the absolute-address
//      swizzle (16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)); one box of 8 rows is stored to global from
//      a source that starts `row0` rows into the buffer (not 1024-byte aligned unless row0 % 8 == 0).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) p10_kernel(const __grid_constant__ CUtensorMap tm, int row0, int* bad_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - ptx::smem_u32(smem_raw));
  // logical value of (row r, 16-bit element e): r * 64 + e, as bf16 bit patterns are irrelevant -> use raw uint16
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {      // 64 rows x 8 chunks
    const int r = i >> 3, c = i & 7;
    uint16_t v[8];
    for (int k = 0; k < 8; ++k) v[k] = (uint16_t)(r * 64 + c * 8 + k);
    *reinterpret_cast<uint4*>(bp + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm)),
                 "r"(base + (uint32_t)row0 * 128u), "r"(0), "r"(0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  (void)bad_out;
}

static void run_p10() {
  uint16_t* dout;
  CK(cudaMalloc(&dout, 8 * 64 * 2));
  CK(cudaFuncSetAttribute(p10_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
  for (int row0 : {0, 8, 1, 3, 27, 29}) {
    CK(cudaMemset(dout, 0xFF, 8 * 64 * 2));
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, 8};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 8};
    make_map(&tm, dout, 2, dims, strides, box);
    p10_kernel<<<1, 128, 16 * 1024>>>(tm, row0, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> h(8 * 64);
    CK(cudaMemcpy(h.data(), dout, 8 * 64 * 2, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 8; ++r)
      for (int e = 0; e < 64; ++e) bad += h[r * 64 + e] != (uint16_t)((row0 + r) * 64 + e);
    printf("P10 TMA store of 8 rows from a source %2d rows (= %4d B) into an absolute-address-swizzled buffer: %d / 512 elements wrong%s\n", row0,
           row0 * 128, bad, bad == 0 ? "  -> swizzle follows the absolute address" : "");
  }
  CK(cudaFree(dout));
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  const char* only = argc > 1 ? argv[1] : "";
  auto want = [&](const char* n) { return !only[0] || strstr(only, n); };
  if (want("p1")) {
    run_p1(10);
    run_p1(16);
    run_p1(18);
  }
  if (want("p5")) run_p5();
  if (want("p6")) run_p6(prop.multiProcessorCount);
  if (want("p4")) run_p4(prop.multiProcessorCount);
  if (want("p3")) run_p3(prop.multiProcessorCount);
  if (want("p7")) run_p7();
  if (want("p8")) run_p8(prop.multiProcessorCount);
  if (want("p9")) run_p9(prop.multiProcessorCount);
  if (want("p10")) run_p10();
  return 0;
}
