// Host-side, exact geometry + coefficient tables of the frame preprocessing that load_video applies to decoded uint8 frames
// (valley/util/data_util.py:271-281):  Resize(256) -> CenterCrop(224).
//
//   * Resize keeps interpolation='nearest' by default and the PIL branch of resize_clip maps that to PIL.Image.BILINEAR
//     (valley/data/video_transform.py:63-66, :269-277): Pillow's two-pass 8-bit convolution.  The tables below restate Pillow's
//     src/libImaging/Resample.c (precompute_coeffs for the triangle filter, normalize_coeffs_8bpc with PRECISION_BITS = 22):
//     doubles on the host, int32 fixed-point weights for the device -- the device side is then pure integer arithmetic and
//     bit-exact.
//   * short side -> 256, long side int(256 * long / short) (video_transform.py:74-81); untouched when the short side is 256 (:56-58)
//   * crop origin int(round((size - 224) / 2.)) with round-half-to-even (video_transform.py:542-543)
// No GPU involved: checked against the oracle on CPU (tests/test_host_logic.py).
#include <cmath>
#include <stdint.h>

#include "../../include/valley_b200.h"

extern "C" void vly_set_error_(const char*);

extern "C" int vly_preprocess_plan(int H, int W, int* new_h, int* new_w, int* crop_y, int* crop_x) {
  if (H <= 0 || W <= 0 || !new_h || !new_w || !crop_y || !crop_x) {
    vly_set_error_("vly_preprocess_plan: bad argument");
    return VLY_ERR_INVALID;
  }
  const int size = 256, crop = 224;
  int nh = H, nw = W;
  if (!((W <= H && W == size) || (H <= W && H == size))) {
    if (W < H) { nw = size; nh = (int)((double)((int64_t)size * H) / (double)W); }
    else { nh = size; nw = (int)((double)((int64_t)size * W) / (double)H); }
  }
  *new_h = nh; *new_w = nw;
  *crop_y = (int)std::nearbyint((nh - crop) / 2.0);     // default rounding mode: half to even, like Python's round()
  *crop_x = (int)std::nearbyint((nw - crop) / 2.0);
  return VLY_OK;
}

extern "C" int vly_resample_coeffs(int in_size, int out_size, int* ksize_out, int32_t* xmin, int32_t* count, int32_t* kk) {
  if (in_size <= 0 || out_size <= 0 || !ksize_out) {
    vly_set_error_("vly_resample_coeffs: bad argument");
    return VLY_ERR_INVALID;
  }
  const int PRECISION_BITS = 32 - 8 - 2;
  double scale = (double)((float)in_size - 0.0f) / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  *ksize_out = ksize;
  if (ksize > 64) {
    vly_set_error_("vly_resample_coeffs: down-scaling factor above 31 is not supported");
    return VLY_ERR_INVALID;
  }
  if (!xmin || !count || !kk) return VLY_OK;             // size query
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    double w[64];                                         // ksize = 2 * ceil(in / out) + 1 <= 64, checked above
    for (int x = 0; x < n; ++x) {
      double a = (x + lo - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    int32_t* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) k[x] = 0;
    for (int x = 0; x < n; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      k[x] = v < 0 ? (int32_t)(-0.5 + v * (1 << PRECISION_BITS)) : (int32_t)(0.5 + v * (1 << PRECISION_BITS));
    }
    xmin[xx] = lo;
    count[xx] = n;
  }
  return VLY_OK;
}
