// Frame preprocessing on the device: decoded uint8 frames -> CLIP-normalised pixels, the tensor the vision tower
// consumes.  Replaces the per-frame CPU/PIL pipeline of load_video (valley/util/data_util.py:271-281).
//
// Byte/integer work, HBM-bound and tiny next to the ViT: two passes like Pillow (horizontal, then vertical, each rounding to
// uint8 -- the intermediate rounding is part of the reference's result), restricted to what the 224x224 centre crop needs:
//   pass 1: only the input rows the cropped output rows reach, only the 224 cropped output columns
//   pass 2: vertical taps over that strip, then (u8 / 255 - mean) / std in IEEE fp32 (div.rn, no FMA contraction) and the
//           frames-first [T,3,224,224] layout in fp32 / fp16 / bf16
// The fixed-point weights come from the host tables (host_preprocess.cpp); nothing here is floating point until the last line.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace vly {

struct PreprocParams {
  const uint8_t* frames;     // [T, H, W, 3]
  int T, H, W;
  int row0, rows;            // input rows [row0, row0 + rows) feed the cropped output rows
  int crop_x, crop_y;        // crop origin in the resized image
  int ksize_h, ksize_v;
  const int32_t *xmin_h, *cnt_h, *kk_h;   // horizontal tables, indexed by the RESIZED column
  const int32_t *ymin_v, *cnt_v, *kk_v;   // vertical tables, indexed by the RESIZED row
  uint8_t* strip;            // [T, rows, 224, 3]
  void* out;                 // [T, 3, 224, 224]
  int out_dtype;             // vly_dtype
  float mean[3], std[3];
};

constexpr int kPrecisionBits = 22;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid (rows, T), 224 threads: one output column of one strip row, three channels
__global__ void __launch_bounds__(224) preprocess_horizontal_kernel(const PreprocParams p) {
  const int r = blockIdx.x, t = blockIdx.y, x = threadIdx.x;
  const int xr = p.crop_x + x;
  const int lo = __ldg(p.xmin_h + xr), n = __ldg(p.cnt_h + xr);
  const int32_t* k = p.kk_h + (size_t)xr * p.ksize_h;
  const uint8_t* src = p.frames + (((size_t)t * p.H + p.row0 + r) * p.W + lo) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int j = 0; j < n; ++j) {
    const int w = __ldg(k + j);
    s0 += src[3 * j + 0] * w;
    s1 += src[3 * j + 1] * w;
    s2 += src[3 * j + 2] * w;
  }
  uint8_t* dst = p.strip + (((size_t)t * p.rows + r) * 224 + x) * 3;
  dst[0] = clip8(s0);
  dst[1] = clip8(s1);
  dst[2] = clip8(s2);
}

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// grid (224, T), 224 threads: one output pixel, three channels
template <typename OutT>
__global__ void __launch_bounds__(224) preprocess_vertical_kernel(const PreprocParams p) {
  const int y = blockIdx.x, t = blockIdx.y, x = threadIdx.x;
  const int yr = p.crop_y + y;
  const int lo = __ldg(p.ymin_v + yr) - p.row0, n = __ldg(p.cnt_v + yr);
  const int32_t* k = p.kk_v + (size_t)yr * p.ksize_v;
  const uint8_t* src = p.strip + (((size_t)t * p.rows + lo) * 224 + x) * 3;
  int s[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int j = 0; j < n; ++j) {
    const int w = __ldg(k + j);
    const uint8_t* q = src + (size_t)j * 224 * 3;
    s[0] += q[0] * w;
    s[1] += q[1] * w;
    s[2] += q[2] * w;
  }
  OutT* out = reinterpret_cast<OutT*>(p.out);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // ClipToTensor: float(u8).div(255); Normalize: sub(mean).div(std) -- three separately rounded fp32 operations
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8(s[c]), 255.f), p.mean[c]), p.std[c]);
    out[(((size_t)t * 3 + c) * 224 + y) * 224 + x] = from_f32<OutT>(v);
  }
}

}  // namespace vly
