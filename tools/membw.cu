// Read-only HBM streaming calibration: (a) LDG.128 grid-stride sum, (b) TMA 1-D bulk copies into a smem ring with
// consumers that only touch the barrier (no math), (c) same with a consumer that reads the smem (LDS) and does FMAs.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../valley_b200/csrc/common.cuh"
using namespace vly;

__global__ void ldg_sum(const uint4* __restrict__ p, size_t n, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = ldg_nc_v4(p + i), b = ldg_nc_v4(p + i + stride), c = ldg_nc_v4(p + i + 2 * stride), d = ldg_nc_v4(p + i + 3 * stride);
    acc += __uint_as_float(a.x ^ b.y ^ c.z ^ d.w);
  }
  if (acc == 123.456f) *out = acc;
}

template <int MATH>
__global__ void __launch_bounds__(544, 1) bulk_ring(const uint8_t* __restrict__ p, size_t bytes_per_cta, int n_stages, int stage_bytes, int copy_bytes, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)n_stages * stage_bytes);
  uint64_t* empty = full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); }
    fence_barrier_init();
  }
  __syncthreads();
  const uint8_t* base = p + (size_t)blockIdx.x * bytes_per_cta;
  const int n_iter = (int)(bytes_per_cta / stage_bytes);
  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], stage_bytes);
        for (int c = 0; c < stage_bytes; c += copy_bytes) bulk_load_1d(sm + (size_t)st * stage_bytes + c, base + (size_t)it * stage_bytes + c, copy_bytes, &full[st]);
        if (++st == n_stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    int st = 0; uint32_t ph = 0;
    float acc[4] = {0, 0, 0, 0};
    const int ct = tid - 32;
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(&full[st], ph);
      if (MATH) {
        for (int r = 0; r < stage_bytes / 8192; ++r) {
          const uint4 w = *reinterpret_cast<const uint4*>(sm + (size_t)st * stage_bytes + r * 8192 + ct * 16);
          acc[r & 3] = fmaf(bf16_lo(w.x), 1.01f, acc[r & 3]); acc[r & 3] = fmaf(bf16_hi(w.x), 1.02f, acc[r & 3]);
          acc[r & 3] = fmaf(bf16_lo(w.y), 1.03f, acc[r & 3]); acc[r & 3] = fmaf(bf16_hi(w.y), 1.04f, acc[r & 3]);
          acc[r & 3] = fmaf(bf16_lo(w.z), 1.05f, acc[r & 3]); acc[r & 3] = fmaf(bf16_hi(w.z), 1.06f, acc[r & 3]);
          acc[r & 3] = fmaf(bf16_lo(w.w), 1.07f, acc[r & 3]); acc[r & 3] = fmaf(bf16_hi(w.w), 1.08f, acc[r & 3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == n_stages) { st = 0; ph ^= 1; }
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] == 123.456f) *out = acc[0];
  }
}

int main() {
  const size_t N = (size_t)8 << 30;   // 8 GiB
  uint8_t* d; float* o;
  cudaMalloc(&d, N); cudaMalloc(&o, 4); cudaMemset(d, 1, N);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int blocks : {148 * 2, 148 * 4, 148 * 8}) for (int thr : {256, 512}) {
    ldg_sum<<<blocks, thr>>>((const uint4*)d, N / 16, o);
    cudaEventRecord(e0); ldg_sum<<<blocks, thr>>>((const uint4*)d, N / 16, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg_sum blocks=%d thr=%d: %.1f GB/s\n", blocks, thr, N / ms / 1e6);
  }
  const size_t per_cta = (N / 148) / (1 << 20) * (1 << 20);
  for (int math : {0, 1}) for (int stages : {3, 6}) for (int copy : {4096, 8192, 32768}) {
    const int stage_bytes = 32768;
    const size_t smem = (size_t)stages * stage_bytes + 256;
    auto k = math ? bulk_ring<1> : bulk_ring<0>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<148, 544, smem>>>(d, per_cta, stages, stage_bytes, copy, o);
    cudaEventRecord(e0); k<<<148, 544, smem>>>(d, per_cta, stages, stage_bytes, copy, o); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("bulk_ring math=%d stages=%d copy=%d: %.1f GB/s  (%s)\n", math, stages, copy, per_cta * 148.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  // copy kernel reference (read+write bytes), like MEASURED_PEAKS
  uint8_t* d2; cudaMalloc(&d2, N / 2);
  cudaMemcpy(d2, d, N / 2, cudaMemcpyDeviceToDevice);
  cudaEventRecord(e0); cudaMemcpy(d2, d, N / 2, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpy D2D 4 GiB: %.1f GB/s (read+write)\n", (double)N / ms / 1e6);
  return 0;
}
