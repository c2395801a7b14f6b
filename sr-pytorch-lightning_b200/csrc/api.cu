// Context lifetime and error reporting of libsrb200 (C ABI in include/srb200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void srb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

unsigned long long g_srb_launches = 0;

extern "C" unsigned long long srb_launch_count(void) { return __atomic_load_n(&g_srb_launches, __ATOMIC_RELAXED); }

extern "C" int srb_abi_version(void) { return SRB_ABI_VERSION; }

extern "C" const char* srb_last_error(void) { return g_err; }

extern "C" int srb_create(int device, srb_ctx** out) {
  SRB_REQUIRE(out != nullptr, "srb_create: out is NULL");
  *out = nullptr;
  int count = 0;
  SRB_CHECK_CUDA(cudaGetDeviceCount(&count));
  SRB_REQUIRE(device >= 0 && device < count, "srb_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  SRB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  SRB_REQUIRE(prop.major == 10, "srb_create: libsrb200 is built for sm_100a only; device %d is sm_%d%d", device,
              prop.major, prop.minor);
  srb_ctx* c = new srb_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->wgrad_sm_budget = 0;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->encode_tiled = nullptr;
  c->weights_dirty = 1;
  c->trace = nullptr;
  c->no_pdl = getenv("SRB200_NO_PDL") != nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete c;
    srb_set_error("srb_create: cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return 3;
  }
  c->encode_tiled = fn;
  *out = c;
  return 0;
}

extern "C" int srb_destroy(srb_ctx* ctx) {
  delete ctx;
  return 0;
}

extern "C" int srb_num_sms(const srb_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
extern "C" int srb_set_wgrad_sm_budget(srb_ctx* ctx, int max_ctas) {
  if (!ctx || max_ctas < 0) return 1;
  ctx->wgrad_sm_budget = max_ctas;
  return 0;
}

/* diagnostics: conv_c64 writes 16 int64 per CTA (event clocks relative to CTA start, see conv_c64.cu)
 * into `buf` (device memory, >= 16 * num_sms int64) on every launch while it is set; NULL turns it off */
extern "C" int srb_debug_set_trace(srb_ctx* ctx, long long* buf) {
  SRB_REQUIRE(ctx, "srb_debug_set_trace: null context");
  ctx->trace = buf;
  return 0;
}
