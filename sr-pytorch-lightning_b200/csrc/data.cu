// GPU data path for training batches: replaces the per-sample PIL work of
// `_SRDataset._get_item` / `_get_patch` in the reference's srdata.py:57-169 — random aligned LR / HR crop, rotation by a
// multiple of 90 degrees, horizontal / vertical flip, `TF.to_tensor` (uint8 HWC -> fp32 CHW / 255) — for images that are
// resident in HBM as uint8 HWC RGB.  ONE launch builds the whole batch; the random choices stay on the host (the reference
// draws them from Python's `random`: srdata.py:77-92,165-166), a step uploads n 56-byte items.
//
// Semantics, in the reference's order: crop (PIL pads with black where the box leaves the image — and it can: `_get_patch`
// reads `lr_image.size` as (h, w) although PIL returns (w, h), srdata.py:152-153,165-169), `TF.rotate(angle)` (PIL:
// counter-clockwise; exact transposes for square patches and multiples of 90), `TF.hflip`, `TF.vflip`, `to_tensor`.
// HBM-bound and tiny: 16 patches are 0.4 MB read, 5.6 MB written (fp32) — one wave of CTAs, a few microseconds.
#include "common.cuh"

namespace {

struct PatchItem {      // mirrors srb_patch_item
  const uint8_t* lr_img;
  const uint8_t* hr_img;
  int32_t lr_h, lr_w, hr_h, hr_w;
  int32_t lr_top, lr_left;
  int32_t angle;          // 0, 90, 180, 270 (counter-clockwise, as PIL)
  int32_t hflip, vflip;
  int32_t pad;
};
static_assert(sizeof(PatchItem) == sizeof(srb_patch_item), "srb_patch_item layout");

// source pixel of output pixel (y, x) of a P x P patch after rotate -> hflip -> vflip
__device__ __forceinline__ void source_of(int y, int x, int P, int angle, int hflip, int vflip, int& sy, int& sx) {
  if (vflip) y = P - 1 - y;
  if (hflip) x = P - 1 - x;
  switch (angle) {
    case 90: sy = x; sx = P - 1 - y; break;            // PIL ROTATE_90: out(y, x) = in(x, P-1-y)
    case 180: sy = P - 1 - y; sx = P - 1 - x; break;
    case 270: sy = P - 1 - x; sx = y; break;           // PIL ROTATE_270: out(y, x) = in(P-1-x, y)
    default: sy = y; sx = x; break;
  }
}

__global__ void __launch_bounds__(256) patch_batch_kernel(const PatchItem* __restrict__ items, int lr_ps, int scale,
                                                          float* __restrict__ lr_out, float* __restrict__ hr_out) {
  const PatchItem it = items[blockIdx.y];
  const int which = blockIdx.z;                        // 0: LR patch, 1: HR patch
  const int P = which ? lr_ps * scale : lr_ps;
  const uint8_t* img = which ? it.hr_img : it.lr_img;
  const int H = which ? it.hr_h : it.lr_h, W = which ? it.hr_w : it.lr_w;
  const int top = which ? it.lr_top * scale : it.lr_top, left = which ? it.lr_left * scale : it.lr_left;
  float* out = (which ? hr_out : lr_out) + (size_t)blockIdx.y * 3 * P * P;
  if (which ? hr_out == nullptr : lr_out == nullptr) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P * P; i += gridDim.x * blockDim.x) {
    const int y = i / P, x = i - y * P;
    int sy, sx;
    source_of(y, x, P, it.angle, it.hflip, it.vflip, sy, sx);
    const int iy = top + sy, ix = left + sx;
    float r = 0.f, g = 0.f, b = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const uint8_t* px = img + ((size_t)iy * W + ix) * 3;
      r = (float)px[0] / 255.0f;                       // TF.to_tensor: .to(float32).div(255), IEEE division
      g = (float)px[1] / 255.0f;
      b = (float)px[2] / 255.0f;
    }
    out[i] = r;
    out[(size_t)P * P + i] = g;
    out[(size_t)2 * P * P + i] = b;
  }
}

}  // namespace

extern "C" int srb_patch_batch(srb_ctx* ctx, const srb_patch_item* items_dev, int n, int lr_patch, int scale, float* lr_out,
                               float* hr_out, void* stream) {
  SRB_REQUIRE(ctx && items_dev && n > 0 && lr_patch > 0 && scale > 0, "srb_patch_batch: bad argument");
  SRB_REQUIRE(lr_out || hr_out, "srb_patch_batch: no output");
  const int P = lr_patch * scale;
  dim3 grid((unsigned)((P * P + 255) / 256 > 64 ? 64 : (P * P + 255) / 256), (unsigned)n, 2);
  patch_batch_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const PatchItem*>(items_dev), lr_patch,
                                                                              scale, lr_out, hr_out);
  SRB_LAUNCH_CHECK();
  return 0;
}
