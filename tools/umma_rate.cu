// Issue-rate probe for the "diagonal-block" UMMA of decode_step_umma_kernel: no TMA, no HBM -- NI issuer threads of one CTA per SM
// issue tcgen05.mma over operands already resident in shared memory, and the kernel reports cycles per MMA and the weight bytes per
// second per SM that rate would retire (M = 64: one MMA = 8 rows x 8 panels x 16 k = 2 KB of weights; M = 128: 16 panels = 4 KB).
// HBM delivers ~48.6 GB/s per SM (7.2 TB/s / 148): a consumer that cannot go faster than that cannot drain the ring after a grid
// barrier, so every barrier bubble is exposed.
//   variants: M x N = 64 x 64 | 128 x 128;  NI = 1..4 issuers;  ACC = 1 | 2 accumulators alternated by each issuer (is the chain of
//   accumulating MMAs serialised on the accumulator?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../valley_b200/csrc/common.cuh"
using namespace vly;

struct RP {
  int ni, nacc, iters, stage_bytes;   // iters = MMAs per issuer
  long long* cycles;                  // [grid]
};

template <int M>
__global__ void __launch_bounds__(576, 1) umma_rate_kernel(const RP p) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* xsw = sm;                       // 80 KB "activation" block
  uint8_t* ring = sm + 80 * 1024;          // 120 KB of "weight" stages
  __shared__ uint64_t done[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&done[i], 1);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long t0 = clock64();
  if (warp >= 1 && warp <= p.ni && lane == 0) {
    const int me = warp - 1;
    constexpr uint32_t idesc = make_idesc_bf16(M, M);
    constexpr int GROUP_BYTES = M * 128;                  // M rows x 128 B: 8 KB (M = 64) or 16 KB (M = 128) per panel group
    const int groups = 120 * 1024 / GROUP_BYTES;
    const uint64_t wd0 = make_smem_desc_sw128(smem_u32(ring), 16, 1024);
    const uint64_t xd0 = make_smem_desc_sw128(smem_u32(xsw), 16, 1024);
    const int xgroups = 80 * 1024 / GROUP_BYTES;
    int g = me, xg = me;
    for (int i = 0; i < p.iters; i += 4) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(me * p.nacc + ((i >> 2) % p.nacc)) * M;
      const uint64_t o = (uint64_t)g * (GROUP_BYTES / 16), ox = (uint64_t)xg * (GROUP_BYTES / 16);
      tc_mma_bf16(d_tmem, wd0 + o, xd0 + ox, idesc, 1u);
      tc_mma_bf16(d_tmem, wd0 + o + 2, xd0 + ox + 2, idesc, 1u);
      tc_mma_bf16(d_tmem, wd0 + o + 4, xd0 + ox + 4, idesc, 1u);
      tc_mma_bf16(d_tmem, wd0 + o + 6, xd0 + ox + 6, idesc, 1u);
      g += p.ni; if (g >= groups) g -= groups;
      xg += p.ni; if (xg >= xgroups) xg -= xgroups;
    }
    tc_commit(&done[me]);
    mbar_wait(&done[me], 0);
  }
  __syncthreads();
  if (tid == 0) p.cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  cudaFuncSetAttribute(umma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(umma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int M : {64, 128})
    for (int ni = 1; ni <= 4; ++ni)
      for (int nacc : {1, 2}) {
        if (M == 128 && ni * nacc * 128 > 512) continue;
        RP p;
        p.ni = ni; p.nacc = nacc; p.iters = 20000; p.stage_bytes = 0; p.cycles = cyc;
        float best_ms = 1e9f;
        long long h[148];
        for (int rep = 0; rep < 3; ++rep) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          if (M == 64) umma_rate_kernel<64><<<148, 576, 200 * 1024>>>(p);
          else umma_rate_kernel<128><<<148, 576, 200 * 1024>>>(p);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best_ms) best_ms = ms;
        }
        cudaError_t err = cudaGetLastError();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
        const double total_mma = (double)p.iters * ni;
        const double cyc_per = mean / total_mma;
        const double bytes_per = M == 64 ? 2048.0 : 4096.0;
        // wall-clock based rate (kernel time includes ~10 us of setup; 20000 MMAs per issuer make it negligible)
        const double gbs_sm = total_mma * bytes_per / (best_ms * 1e-3) / 1e9;
        printf("M=N=%3d issuers %d accumulators/issuer %d: %6.1f cycles/MMA (per CTA, all issuers)  -> %6.1f GB/s of weights per SM (%.2f ms)  %s\n", M, ni,
               nacc, cyc_per, gbs_sm, best_ms, cudaGetErrorString(err));
        if (err != cudaSuccess) return 1;
      }
  return 0;
}
