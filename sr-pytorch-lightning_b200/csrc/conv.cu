// Dispatch of the convolution entry points (include/srb200.h) to the tcgen05 or CUDA-core kernels.
#include <stdlib.h>
#include <vector>

#include "common.cuh"

int srb_conv_simt(srb_ctx*, const srb_conv_desc*, const void*, const void*, const float*, const void*, const void*,
                  void*, void*, float*, cudaStream_t);
int srb_conv_umma(srb_ctx*, const srb_conv_desc*, const void*, const void*, const float*, const void*, const void*,
                  void*, void*, float*, cudaStream_t);
int srb_conv_umma_bn(const srb_conv_desc*);
int srb_wgrad_simt(srb_ctx*, const srb_wgrad_desc*, const void*, const void*, float*, float*, cudaStream_t);
int srb_wgrad_umma(srb_ctx*, const srb_wgrad_desc*, const void*, const void*, float*, float*, cudaStream_t);
int srb_wgrad_umma_ok(const srb_wgrad_desc*);
int srb_wgrad_umma_batched(srb_ctx*, const srb_wgrad_desc*, const void* const*, const void* const*, float* const*, int,
                           cudaStream_t);
int srb_colsum_launch(srb_ctx*, const void*, int, int, int, int64_t, int, float*, int, float, int, cudaStream_t);
int srb_colsum_batched_ok(const void* x, int cs, int co, int C, int dtype);
int srb_colsum_batched_launch(srb_ctx*, int n, const void* const* xs, const int* cs, const int* co, const int* C,
                              const int64_t* npix, float* const* outs, const int* accumulate, const float* alpha,
                              const int* shuffle, cudaStream_t st);

static int check_conv_desc(const srb_conv_desc* d, const void* x, const void* w, const void* res, const void* mask,
                           void* y, void* y2, float* colsum) {
  SRB_REQUIRE(d && x && w && y, "srb_conv: null argument");
  SRB_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "srb_conv: empty shape N=%d H=%d W=%d Cin=%d Cout=%d",
              d->N, d->H, d->W, d->Cin, d->Cout);
  SRB_REQUIRE(d->dtype == SRB_F32 || d->dtype == SRB_BF16, "srb_conv: bad dtype %d", d->dtype);
  SRB_REQUIRE(d->ksize >= 1 && (d->ksize & 1), "srb_conv: kernel size must be odd, got %d", d->ksize);
  SRB_REQUIRE(d->x_co + d->Cin <= d->x_cs, "srb_conv: input channel slice [%d,%d) exceeds stride %d", d->x_co,
              d->x_co + d->Cin, d->x_cs);
  const int rr = d->shuffle > 1 ? d->shuffle * d->shuffle : 1;
  SRB_REQUIRE(d->shuffle == 0 || (d->shuffle >= 2 && d->shuffle <= 8), "srb_conv: shuffle must be 0 or 2..8");
  SRB_REQUIRE(d->Cout % rr == 0, "srb_conv: Cout %d not divisible by r^2=%d", d->Cout, rr);
  SRB_REQUIRE(d->y_co + d->Cout / rr <= d->y_cs, "srb_conv: output channel slice exceeds stride");
  SRB_REQUIRE(!(d->flags & SRB_RESIDUAL) || res, "srb_conv: RESIDUAL without residual pointer");
  SRB_REQUIRE(!(d->flags & SRB_MASK) || mask, "srb_conv: MASK without mask pointer");
  SRB_REQUIRE(!(d->flags & SRB_OUT2) || y2, "srb_conv: OUT2 without y2 pointer");
  SRB_REQUIRE(!(d->flags & SRB_COLSUM) || (colsum && d->shuffle <= 1 && (d->colsum_groups == 1 || d->colsum_groups == d->N)),
              "srb_conv: COLSUM needs colsum pointer, no shuffle and groups in {1, N}");
  return 0;
}

extern "C" int srb_conv(srb_ctx* ctx, const srb_conv_desc* d, const void* x, const void* w, const float* bias,
                        const void* res, const void* mask, void* y, void* y2, float* colsum, void* stream) {
  SRB_REQUIRE(ctx, "srb_conv: null context");
  int rc = check_conv_desc(d, x, w, res, mask, y, y2, colsum);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int backend = d->backend;
  if (backend == SRB_BACKEND_AUTO) backend = srb_conv_umma_bn(d) ? SRB_BACKEND_UMMA : SRB_BACKEND_SIMT;
  if (backend == SRB_BACKEND_UMMA) return srb_conv_umma(ctx, d, x, w, bias, res, mask, y, y2, colsum, st);
  return srb_conv_simt(ctx, d, x, w, bias, res, mask, y, y2, colsum, st);
}

/* 1 if srb_conv would take the tcgen05 path for this descriptor with backend AUTO (the caller
 * needs to know which weight packing to provide). */
extern "C" int srb_conv_uses_umma(const srb_conv_desc* d) {
  if (!d) return 0;
  if (d->backend == SRB_BACKEND_SIMT) return 0;
  return srb_conv_umma_bn(d) != 0;
}

extern "C" int srb_conv_wgrad(srb_ctx* ctx, const srb_wgrad_desc* d, const void* x, const void* gy, float* dw,
                              float* dbias, void* stream) {
  SRB_REQUIRE(ctx && d && x && gy && dw, "srb_conv_wgrad: null argument");
  SRB_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "srb_conv_wgrad: empty shape");
  SRB_REQUIRE(d->dtype == SRB_F32 || d->dtype == SRB_BF16, "srb_conv_wgrad: bad dtype %d", d->dtype);
  SRB_REQUIRE(d->x_co + d->Cin <= d->x_cs && d->g_co + d->Cout <= d->g_cs, "srb_conv_wgrad: channel slice exceeds stride");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int backend = d->backend;
  if (backend == SRB_BACKEND_AUTO) backend = srb_wgrad_umma_ok(d) ? SRB_BACKEND_UMMA : SRB_BACKEND_SIMT;
  if (backend == SRB_BACKEND_UMMA) return srb_wgrad_umma(ctx, d, x, gy, dw, dbias, st);
  return srb_wgrad_simt(ctx, d, x, gy, dw, dbias, st);
}

extern "C" int srb_wgrad_uses_umma(const srb_wgrad_desc* d) {
  if (!d || d->backend == SRB_BACKEND_SIMT) return 0;
  return srb_wgrad_umma_ok(d);
}

/* Many weight gradients in one call: the tcgen05-eligible ones share batched launches (one CTA per
 * 64x64x9 block, whole pixel reduction per CTA — see wgrad_umma.cu), the rest go to the CUDA-core
 * kernel one by one. */
extern "C" int srb_conv_wgrad_batched(srb_ctx* ctx, const srb_wgrad_item* items, int n, void* stream) {
  SRB_REQUIRE(ctx && (items || n == 0), "srb_conv_wgrad_batched: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::vector<srb_wgrad_desc> descs;
  std::vector<const void*> xs, gys;
  std::vector<float*> dws;
  // bias gradients of the batch share one launch (misc.cu colsum_batched_kernel); SRB200_NO_BATCHED_COLSUM=1
  // keeps one launch per bias (A/B measurements)
  static const bool batch_colsums = [] {
    const char* e = getenv("SRB200_NO_BATCHED_COLSUM");
    return !(e && e[0] && e[0] != '0');
  }();
  std::vector<const void*> cx;
  std::vector<int> ccs, cco, cC, cacc, cshuf;
  std::vector<int64_t> cnp;
  std::vector<float*> cout_;
  std::vector<float> calpha;
  for (int i = 0; i < n; ++i) {
    const srb_wgrad_item& it = items[i];
    if (it.gy && it.dbias && !it.dw) {      // bias gradient only (its weight gradient ran elsewhere): column sums of gy
      const int64_t npix = (int64_t)it.d.N * it.d.H * it.d.W;
      if (batch_colsums && srb_colsum_batched_ok(it.gy, it.d.g_cs, it.d.g_co, it.d.Cout, it.d.dtype)) {
        cx.push_back(it.gy);
        ccs.push_back(it.d.g_cs);
        cco.push_back(it.d.g_co);
        cC.push_back(it.d.Cout);
        cnp.push_back(npix);
        cout_.push_back(it.dbias);
        cacc.push_back(it.d.accumulate);
        calpha.push_back(it.d.alpha);
        cshuf.push_back(it.d.shuffle);
      } else {
        int rc = srb_colsum_launch(ctx, it.gy, it.d.g_cs, it.d.g_co, it.d.Cout, npix, it.d.dtype, it.dbias, it.d.accumulate,
                                   it.d.alpha, it.d.shuffle, st);
        if (rc) return rc;
      }
      continue;
    }
    SRB_REQUIRE(it.x && it.gy && it.dw, "srb_conv_wgrad_batched: item %d has a null pointer", i);
    const bool umma = it.d.backend != SRB_BACKEND_SIMT && srb_wgrad_umma_ok(&it.d);
    if (umma) {
      descs.push_back(it.d);
      xs.push_back(it.x);
      gys.push_back(it.gy);
      dws.push_back(it.dw);
      if (it.dbias) {
        const int64_t npix = (int64_t)it.d.N * it.d.H * it.d.W;
        if (batch_colsums && srb_colsum_batched_ok(it.gy, it.d.g_cs, it.d.g_co, it.d.Cout, it.d.dtype)) {
          cx.push_back(it.gy);
          ccs.push_back(it.d.g_cs);
          cco.push_back(it.d.g_co);
          cC.push_back(it.d.Cout);
          cnp.push_back(npix);
          cout_.push_back(it.dbias);
          cacc.push_back(it.d.accumulate);
          calpha.push_back(it.d.alpha);
          cshuf.push_back(it.d.shuffle);
        } else {
          int rc = srb_colsum_launch(ctx, it.gy, it.d.g_cs, it.d.g_co, it.d.Cout, npix, it.d.dtype, it.dbias,
                                     it.d.accumulate, it.d.alpha, it.d.shuffle, st);
          if (rc) return rc;
        }
      }
    } else {
      int rc = srb_conv_wgrad(ctx, &it.d, it.x, it.gy, it.dw, it.dbias, stream);
      if (rc) return rc;
    }
  }
  if (!cx.empty()) {
    int rc = srb_colsum_batched_launch(ctx, (int)cx.size(), cx.data(), ccs.data(), cco.data(), cC.data(), cnp.data(),
                                       cout_.data(), cacc.data(), calpha.data(), cshuf.data(), st);
    if (rc) return rc;
  }
  if (descs.empty()) return 0;
  return srb_wgrad_umma_batched(ctx, descs.data(), xs.data(), gys.data(), dws.data(), (int)descs.size(), st);
}
