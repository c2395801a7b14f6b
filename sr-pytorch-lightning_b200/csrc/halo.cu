// Halo exchange over NVLink peer memory for strip-parallel inference (include/srb200.h "spatial tiling";
// BASELINE.json configs[4]: EDSR x4 large 960x540 -> 3840x2160 on 8 GPUs).  The reference has no tiling
// (SRModel.predict_step, srmodel.py:375-380, pushes the whole frame through forward on one device).
//
// Each rank owns a row strip of the frame; a layer's output buffer is [1, t + rows + t, W, C] with t halo rows above and
// below.  Round 1 traded the halo rows with one NCCL send/recv pair per neighbour and layer, issued eagerly from Python:
// 69 layers x 2 pairs cost as much as a strip's convolutions (4.5x on 8 GPUs).  Here the strip buffers are CUDA-IPC
// allocations that the two neighbours map, and ONE kernel per layer
//   1. copies this rank's first / last t owned rows straight into the neighbours' halo rows (16-byte stores through
//      NVLink: 491 KB per neighbour for a 256-channel LR layer),
//   2. publishes them: __threadfence_system, then the last CTA to finish stores the frame number into the neighbours'
//      flag slot of this layer (st.release.sys),
//   3. waits for the same from its own neighbours (ld.acquire.sys, bounded),
// so the next convolution, ordered behind it on the stream, reads complete halos.  Step 0, before any of that: a
// ready-to-receive handshake.  The convolution that produced this layer wrote the WHOLE buffer, halo rows included (they
// hold incomplete sums), so a neighbour must not push into them before that convolution has finished: every rank first
// tells its neighbours "my layer-L buffer may be written" and pushes only after it has heard the same from them.  No host involvement: the whole strip
// forward (convs + exchanges) is captured as one CUDA graph.  Flags carry the frame number (monotonic), every layer has its
// own slot and its own buffer, so a rank that runs ahead into the next frame cannot overwrite anything still in use: a
// neighbour finishes frame f only after it has consumed this rank's halos of every layer of frame f.
#include "common.cuh"

namespace {

struct HaloParams {
  const uint4* src_top;      // this rank's first t owned rows
  const uint4* src_bot;      // this rank's last t owned rows
  uint4* dst_up;             // upper neighbour's bottom halo rows (peer memory) or NULL
  uint4* dst_dn;             // lower neighbour's top halo rows (peer memory) or NULL
  long long n16;             // 16-byte units per slab
  long long* flag_up;        // upper neighbour's slots of this layer, its "from below" pair {ready, data} (peer) or NULL
  long long* flag_dn;        // lower neighbour's "from above" pair (peer) or NULL
  const long long* wait_up;  // own "from above" pair (written by the upper neighbour) or NULL
  const long long* wait_dn;  // own "from below" pair or NULL
  const long long* frame;    // device counter: current frame number (>= 1)
  unsigned int* done;        // CTA completion counter (self-resetting)
};

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void wait_flag(const long long* w, long long f, const char* what) {
  const unsigned long long t0 = gtimer();
  unsigned int spins = 0;
  while (ld_acquire_sys(w) < f) {
    if ((++spins & 0xFFFu) == 0 && gtimer() - t0 > 10000000000ull) {      // 10 s: a peer died or the protocol is broken
      printf("srb200: halo exchange timed out waiting for %s (frame %lld, have %lld)\n", what, f, ld_acquire_sys(w));
      __trap();
    }
  }
}

__global__ void __launch_bounds__(256) halo_exchange_kernel(const HaloParams p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long f = *p.frame;
  // 0. handshake: my buffer of this layer is complete (the conv before this kernel on the stream has finished), you may
  //    write its halo rows; I push once you have said the same
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      if (p.flag_up) st_release_sys(p.flag_up, f);
      if (p.flag_dn) st_release_sys(p.flag_dn, f);
    }
    if (p.wait_up) wait_flag(p.wait_up, f, "the upper neighbour's ready flag");
    if (p.wait_dn) wait_flag(p.wait_dn, f, "the lower neighbour's ready flag");
  }
  __syncthreads();
  if (p.dst_up)
    for (long long i = i0; i < p.n16; i += stride) p.dst_up[i] = p.src_top[i];
  if (p.dst_dn)
    for (long long i = i0; i < p.n16; i += stride) p.dst_dn[i] = p.src_bot[i];
  __threadfence_system();                   // this thread's peer stores are performed system-wide
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(p.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  *p.done = 0;                               // ready for the next launch (stream order)
  __threadfence_system();
  if (p.flag_up) st_release_sys(p.flag_up + 1, f);
  if (p.flag_dn) st_release_sys(p.flag_dn + 1, f);
  if (p.wait_up) wait_flag(p.wait_up + 1, f, "the upper neighbour's rows");
  if (p.wait_dn) wait_flag(p.wait_dn + 1, f, "the lower neighbour's rows");
}

}  // namespace

extern "C" int srb_halo_exchange(srb_ctx* ctx, const srb_halo_desc* d, void* stream) {
  SRB_REQUIRE(ctx && d, "srb_halo_exchange: null argument");
  SRB_REQUIRE(d->slab_bytes > 0 && d->slab_bytes % 16 == 0, "srb_halo_exchange: slab size %lld must be a positive multiple of 16", (long long)d->slab_bytes);
  SRB_REQUIRE(d->frame && d->done, "srb_halo_exchange: frame counter / completion counter missing");
  SRB_REQUIRE((d->dst_up == nullptr) == (d->flag_up == nullptr) && (d->dst_up == nullptr) == (d->wait_up == nullptr),
              "srb_halo_exchange: upper neighbour needs all of dst / flag / wait, or none");
  SRB_REQUIRE((d->dst_dn == nullptr) == (d->flag_dn == nullptr) && (d->dst_dn == nullptr) == (d->wait_dn == nullptr),
              "srb_halo_exchange: lower neighbour needs all of dst / flag / wait, or none");
  SRB_REQUIRE(!d->dst_up || d->src_top, "srb_halo_exchange: src_top missing");
  SRB_REQUIRE(!d->dst_dn || d->src_bot, "srb_halo_exchange: src_bot missing");
  const void* ptrs[4] = {d->src_top, d->src_bot, d->dst_up, d->dst_dn};
  for (const void* q : ptrs) SRB_REQUIRE(((uintptr_t)q & 15) == 0, "srb_halo_exchange: slabs must be 16-byte aligned");
  HaloParams p;
  p.src_top = static_cast<const uint4*>(d->src_top);
  p.src_bot = static_cast<const uint4*>(d->src_bot);
  p.dst_up = static_cast<uint4*>(d->dst_up);
  p.dst_dn = static_cast<uint4*>(d->dst_dn);
  p.n16 = d->slab_bytes / 16;
  p.flag_up = reinterpret_cast<long long*>(d->flag_up);
  p.flag_dn = reinterpret_cast<long long*>(d->flag_dn);
  p.wait_up = reinterpret_cast<const long long*>(d->wait_up);
  p.wait_dn = reinterpret_cast<const long long*>(d->wait_dn);
  p.frame = reinterpret_cast<const long long*>(d->frame);
  p.done = reinterpret_cast<unsigned int*>(d->done);
  if (!p.dst_up && !p.dst_dn) return 0;      // a single strip: nothing to trade
  // few CTAs: the copy is 0.5-2 MB over one NVLink hop; the SMs stay free for the previous conv's tail (PDL-less, stream order)
  long long want = (p.n16 + 256 * 8 - 1) / (256 * 8);
  const int grid = (int)(want < 1 ? 1 : (want > 32 ? 32 : want));
  halo_exchange_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SRB_LAUNCH_CHECK();
  return 0;
}

// ---- CUDA-IPC backed device memory (the strip buffers and flag slots the neighbours map) -----------------------------
extern "C" int srb_ipc_alloc(srb_ctx* ctx, size_t bytes, void** ptr, unsigned char handle[64]) {
  SRB_REQUIRE(ctx && ptr && handle && bytes > 0, "srb_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  SRB_CHECK_CUDA(cudaSetDevice(ctx->device));
  void* q = nullptr;
  SRB_CHECK_CUDA(cudaMalloc(&q, bytes));
  cudaError_t e = cudaMemset(q, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, q);
  if (e != cudaSuccess) {
    cudaFree(q);
    srb_set_error("srb_ipc_alloc: %s", cudaGetErrorString(e));
    return 1;
  }
  memcpy(handle, &h, 64);
  *ptr = q;
  return 0;
}

extern "C" int srb_ipc_open(srb_ctx* ctx, const unsigned char handle[64], void** ptr) {
  SRB_REQUIRE(ctx && ptr && handle, "srb_ipc_open: bad argument");
  SRB_CHECK_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  SRB_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int srb_ipc_close(srb_ctx* ctx, void* ptr) {
  SRB_REQUIRE(ctx, "srb_ipc_close: null context");
  if (ptr) SRB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int srb_ipc_free(srb_ctx* ctx, void* ptr) {
  SRB_REQUIRE(ctx, "srb_ipc_free: null context");
  if (ptr) SRB_CHECK_CUDA(cudaFree(ptr));
  return 0;
}
