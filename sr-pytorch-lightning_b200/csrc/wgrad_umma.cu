// tcgen05 weight-gradient kernel (placeholder until the MN-major split-K kernel lands: reports
// "not eligible" so the dispatcher uses the CUDA-core kernel).
#include "common.cuh"

int srb_wgrad_umma_ok(const srb_wgrad_desc*) { return 0; }

int srb_wgrad_umma(srb_ctx*, const srb_wgrad_desc*, const void*, const void*, float*, float*, cudaStream_t) {
  srb_set_error("srb_conv_wgrad(umma): not available in this build");
  return 5;
}
