// Streaming-rate calibration of the decode step's weight ring, with the access pattern of decode_step_kernel's producer:
// CTA c streams work units c, c + grid, c + 2 grid, ... of a row-major [N, K] bf16 matrix; a unit is ROWS consecutive rows,
// cut into ceil(K / KC) ring stages; each stage is ROWS bulk copies of KC*2 bytes (rows padded by `pad` bytes in smem).
// Consumers only wait / arrive (no math): this is the ceiling a (ROWS, KC, stages, in-flight) geometry can reach on this part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ringbw tools/ringbw.cu && tools/ringbw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../valley_b200/csrc/common.cuh"
using namespace vly;

__global__ void __launch_bounds__(576, 1) ring_stream(const __nv_bfloat16* __restrict__ W, int N, int K, int ROWS, int KC, int pad, int n_stages,
                                                      int n_inflight, int one_copy, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int row_stride = KC * 2 + pad;
  const int stage_b = ROWS * row_stride;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)n_stages * stage_b);
  uint64_t* empty = full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); }
    fence_barrier_init();
  }
  __syncthreads();
  const int n_groups = N / ROWS, n_slices = (K + KC - 1) / KC;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0, wst = 0, issued = 0;
      uint32_t ph = 0, wph = 0;
      for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        for (int s = 0; s < n_slices; ++s) {
          const int kc = min(KC, K - s * KC);
          mbar_wait(&empty[st], ph ^ 1);
          if (issued >= n_inflight) {
            mbar_wait(&full[wst], wph);
            if (++wst == n_stages) { wst = 0; wph ^= 1; }
          }
          ++issued;
          mbar_expect_tx(&full[st], (uint32_t)ROWS * kc * 2);
          uint8_t* dst = sm + (size_t)st * stage_b;
          const __nv_bfloat16* src = W + (size_t)g * ROWS * K + (size_t)s * KC;
          if (one_copy && kc == K && pad == 0) bulk_load_1d(dst, src, (uint32_t)ROWS * kc * 2, &full[st]);
          else
            for (int r = 0; r < ROWS; ++r) bulk_load_1d(dst + r * row_stride, src + (size_t)r * K, (uint32_t)kc * 2, &full[st]);
          if (++st == n_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp <= 16) {
    int st = 0;
    uint32_t ph = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x)
      for (int s = 0; s < n_slices; ++s) {
        mbar_wait(&full[st], ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == n_stages) { st = 0; ph ^= 1; }
      }
  }
  if (out == nullptr && tid == 0) sm[0] = 1;
}

int main() {
  const size_t BYTES = (size_t)6 << 30;
  uint8_t* d;
  cudaMalloc(&d, BYTES);
  cudaMemset(d, 1, BYTES);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(ring_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { const char* name; int K, ROWS, KC, pad; };
  const Cfg cfgs[] = {
      {"7B  B=1 K=4096  4x4096 (now)", 4096, 4, 4096, 0},        {"7B  B=1 K=11008 4x4096 (now: 4096+4096+2816)", 11008, 4, 4096, 0},
      {"7B  B=1 K=11008 4x3712 (3 even slices, 29 KB)", 11008, 4, 3712, 0}, {"7B  B=1 K=11008 2x5504 (2 even, 22 KB)", 11008, 2, 5504, 0},
      {"13B B=4 K=5120  8x2048 (now: 2048+2048+1024)", 5120, 8, 2048, 64}, {"13B B=4 K=5120  8x2560 (2 even, 41 KB)", 5120, 8, 2560, 64},
      {"13B B=4 K=5120  4x2560 (2 even, 20 KB)", 5120, 4, 2560, 64},        {"13B B=4 K=5120  4x5120 (1 slice, 41 KB)", 5120, 4, 5120, 64},
      {"13B B=4 K=5120  8x1280 (4 even, 20 KB)", 5120, 8, 1280, 64},
      {"13B B=4 K=13824 8x2048 (now: 6x2048+1536)", 13824, 8, 2048, 64},  {"13B B=4 K=13824 8x2304 (6 even, 37 KB)", 13824, 8, 2304, 64},
      {"13B B=4 K=13824 8x1728 (8 even, 28 KB)", 13824, 8, 1728, 64},      {"13B B=4 K=13824 4x3456 (4 even, 28 KB)", 13824, 4, 3456, 64},
      {"13B B=4 K=13824 4x4608 (3 even, 37 KB)", 13824, 4, 4608, 64},      {"13B B=1 K=5120  4x5120 (1 slice, 40 KB)", 5120, 4, 5120, 0},
      {"13B B=1 K=13824 4x4608 (3 even, 36 KB)", 13824, 4, 4608, 0},       {"13B B=1 K=13824 4x3456 (4 even, 27 KB)", 13824, 4, 3456, 0},
  };
  for (const Cfg& c : cfgs) {
    const int stage_b = c.ROWS * (c.KC * 2 + c.pad);
    const int N = (int)(BYTES / ((size_t)c.K * 2) / (148 * c.ROWS) * (148 * c.ROWS));
    printf("%s  stage %d B\n", c.name, stage_b);
    for (int stages = 2; stages <= 7; ++stages) {
      const size_t smem = (size_t)stages * stage_b + 256;
      if (smem > 200 * 1024) break;
      for (int infl = 2; infl <= stages; ++infl) {
        if (infl < stages && infl != 3 && infl != 4) continue;
        float best = 0.f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          ring_stream<<<148, 576, smem>>>((const __nv_bfloat16*)d, N, c.K, c.ROWS, c.KC, c.pad, stages, infl, 1, (float*)d);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          const float gbs = (float)((double)N * c.K * 2 / ms / 1e6);
          if (rep > 0 && gbs > best) best = gbs;
        }
        printf("    stages %d inflight %d: %7.1f GB/s  (%s)\n", stages, infl, best, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
