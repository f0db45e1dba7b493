// Streaming-rate calibration of the decode step's weight ring, with the access pattern of decode_step_kernel's producer:
// CTA c streams work units c, c + grid, c + 2 grid, ... of a row-major [N, K] bf16 matrix; a unit is ROWS consecutive rows,
// cut into ceil(K / KC) ring stages; each stage is ROWS bulk copies of KC*2 bytes (rows padded by `pad` bytes in smem).
// Consumer models (what hands a landed stage back to the producer, and how fast):
//   mode 0 : wait + arrive only                         -> the ceiling of a (ROWS, KC, stages, in-flight) geometry
//   mode 1 : the CUDA-core consumer of the B = 1 path   (512 threads, 16 B of every row per thread, fp32 FMAs, x from smem)
//   mode 2 : the mma.sync consumer of the B = 2..4 path with ONE accumulator per warp (a dependent HMMA chain per stage)
//   mode 3 : the same with FOUR independent accumulators
//   mode 4 : HMMA only (operands stay in registers: no shared-memory loads)      -- is it the tensor pipe ...
//   mode 5 : the shared-memory loads only (no HMMA)                              -- ... or the LDS traffic?
//   mode 6 : x fragments cached in registers (one LDS.128 of weights per two HMMAs)
//   mode 7 : weights as the A operand (16 weight rows x 16 k per HMMA: half the HMMAs per byte), x fragments in registers
//   +8     : plus the per-unit handoff of the 16 per-warp partials to a finalize warp (4 mbarrier slots), as in the kernel
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ringbw tools/ringbw.cu && tools/ringbw [quick]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../valley_b200/csrc/common.cuh"
using namespace vly;

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct RP {
  const __nv_bfloat16* W;
  int N, K, ROWS, KC, pad, n_stages, n_inflight, mode, xs_stride, xrows;
  float* out;
};

__global__ void __launch_bounds__(576, 1) ring_stream(const RP p) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int row_stride = p.KC * 2 + p.pad;
  const int stage_b = p.ROWS * row_stride;
  uint8_t* ring = sm;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ring + (size_t)p.n_stages * stage_b);      // [4][xs_stride]
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + (size_t)p.xrows * p.xs_stride * 2);
  uint64_t* empty = full + 8;
  uint64_t* red_full = empty + 8;
  uint64_t* red_empty = red_full + 4;
  float* red = reinterpret_cast<float*>(red_empty + 4);                                             // [4][16][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mode = p.mode & 7;
  const bool handoff = (p.mode & 8) != 0;
  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); }
    for (int i = 0; i < 4; ++i) { mbar_init(&red_full[i], 16); mbar_init(&red_empty[i], 1); }
    fence_barrier_init();
  }
  for (int i = tid; i < p.xrows * p.xs_stride; i += blockDim.x) xs[i] = __float2bfloat16(0.001f * (i & 63));
  __syncthreads();
  const int n_groups = p.N / p.ROWS, n_slices = (p.K + p.KC - 1) / p.KC;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0, wst = 0, issued = 0, confirmed = 0;
      uint32_t ph = 0, wph = 0;
      for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        for (int s = 0; s < n_slices; ++s) {
          const int kc = min(p.KC, p.K - s * p.KC);
          mbar_wait(&empty[st], ph ^ 1);
          while (issued - confirmed >= p.n_inflight) {
            mbar_wait(&full[wst], wph);
            if (++wst == p.n_stages) { wst = 0; wph ^= 1; }
            ++confirmed;
          }
          ++issued;
          mbar_expect_tx(&full[st], (uint32_t)p.ROWS * kc * 2);
          uint8_t* dst = ring + (size_t)st * stage_b;
          const __nv_bfloat16* src = p.W + (size_t)g * p.ROWS * p.K + (size_t)s * p.KC;
          if (kc == p.K && p.pad == 0) bulk_load_1d(dst, src, (uint32_t)p.ROWS * kc * 2, &full[st]);
          else
            for (int r = 0; r < p.ROWS; ++r) bulk_load_1d(dst + r * row_stride, src + (size_t)r * p.K, (uint32_t)kc * 2, &full[st]);
          if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp <= 16) {
    const int cw = warp - 1, ct = tid - 32;
    int st = 0;
    uint32_t ph = 0;
    unsigned unit_no = 0;
    float sink = 0.f;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++unit_no) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0}, d3[4] = {0, 0, 0, 0};
      for (int s = 0; s < n_slices; ++s) {
        const int kc = min(p.KC, p.K - s * p.KC);
        mbar_wait(&full[st], ph);
        const uint8_t* stg = ring + (size_t)st * stage_b;
        if (mode == 1) {
          for (int c8 = ct * 8; c8 < kc; c8 += 4096) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)s * p.KC + c8);
            const float xf[8] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y), bf16_lo(xv.z), bf16_hi(xv.z), bf16_lo(xv.w), bf16_hi(xv.w)};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (r < p.ROWS) {
                const uint4 wv = *reinterpret_cast<const uint4*>(stg + r * row_stride + c8 * 2);
                const float wf[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y), bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[r] = fmaf(wf[e], xf[e], acc[r]);
              }
            }
          }
        } else if (mode >= 2 && mode <= 6) {
          const int gq = lane >> 2, tq = lane & 3;
          const uint8_t* wrow = stg + gq * row_stride;
          const __nv_bfloat16* xrow = xs + (size_t)(gq & 3) * p.xs_stride + (size_t)s * p.KC;
          const bool w_ok = gq < p.ROWS, x_ok = gq < 4;
          int it = 0;
          uint4 wreg = make_uint4(lane, 1, 2, 3), xreg = make_uint4(4, lane, 6, 7);
#pragma unroll 4
          for (int k = cw * 32; k < kc; k += 512, ++it) {
            uint4 wb = wreg, xa = xreg;
            if (mode != 4) {
              if (w_ok) wb = *reinterpret_cast<const uint4*>(wrow + (k + tq * 8) * 2);
              if (mode != 6 && x_ok) xa = *reinterpret_cast<const uint4*>(xrow + k + tq * 8);
            }
            if (mode == 5) {
              d0[0] += __uint_as_float(wb.x ^ xa.y);
            } else if (mode == 2) {
              mma16816(d0, xa.x, 0u, xa.y, 0u, wb.x, wb.y);
              mma16816(d0, xa.z, 0u, xa.w, 0u, wb.z, wb.w);
            } else if (it & 1) {
              mma16816(d2, xa.x, 0u, xa.y, 0u, wb.x, wb.y);
              mma16816(d3, xa.z, 0u, xa.w, 0u, wb.z, wb.w);
            } else {
              mma16816(d0, xa.x, 0u, xa.y, 0u, wb.x, wb.y);
              mma16816(d1, xa.z, 0u, xa.w, 0u, wb.z, wb.w);
            }
          }
        } else if (mode == 7) {
          // A = weights: lane (g, t) loads 16 B of weight rows g and g + 8 (ROWS = 16 per unit); B = x from registers
          const int gq = lane >> 2, tq = lane & 3;
          const uint8_t* w0 = stg + gq * row_stride;
          const uint8_t* w1 = stg + (gq + 8) * row_stride;
          uint4 xreg = make_uint4(4, lane, 6, 7);
          int it = 0;
#pragma unroll 4
          for (int k = cw * 32; k < kc; k += 512, ++it) {
            const uint4 a = *reinterpret_cast<const uint4*>(w0 + (k + tq * 8) * 2);
            const uint4 b = *reinterpret_cast<const uint4*>(w1 + (k + tq * 8) * 2);
            if (it & 1) {
              mma16816(d2, a.x, b.x, a.y, b.y, xreg.x, xreg.y);
              mma16816(d3, a.z, b.z, a.w, b.w, xreg.z, xreg.w);
            } else {
              mma16816(d0, a.x, b.x, a.y, b.y, xreg.x, xreg.y);
              mma16816(d1, a.z, b.z, a.w, b.w, xreg.z, xreg.w);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
      const float v = acc[0] + acc[1] + acc[2] + acc[3] + d0[0] + d0[1] + d1[0] + d1[1] + d2[0] + d2[1] + d3[0] + d3[1];
      if (handoff) {
        const int slot = unit_no & 3;
        const uint32_t round = (unit_no >> 2) & 1;
        mbar_wait(&red_empty[slot], round ^ 1);
        red[(slot * 16 + cw) * 32 + lane] = v;
        __syncwarp();
        if (lane == 0) mbar_arrive(&red_full[slot]);
      } else {
        sink += v;
      }
    }
    if (sink == 123.456f) *p.out = sink;
  } else if (handoff) {
    unsigned unit_no = 0;
    float sink = 0.f;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++unit_no) {
      const int slot = unit_no & 3;
      const uint32_t round = (unit_no >> 2) & 1;
      mbar_wait(&red_full[slot], round);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) t += red[(slot * 16 + w) * 32 + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&red_empty[slot]);
      sink += t;
    }
    if (sink == 123.456f) *p.out = sink;
  }
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  const size_t BYTES = (size_t)6 << 30;
  uint8_t* d;
  cudaMalloc(&d, BYTES);
  cudaMemset(d, 1, BYTES);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(ring_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { const char* name; int K, ROWS, KC, pad, xrows; };   // xrows: activation rows staged in smem (smem left for the ring)
  const Cfg cfgs[] = {
      {"7B  B=1 K=4096  4x4096", 4096, 4, 4096, 0, 1},           {"7B  B=1 K=11008 4x3672", 11008, 4, 3672, 0, 1},
      {"13B B=1 K=5120  4x2560", 5120, 4, 2560, 0, 1},           {"13B B=1 K=5120  4x5120 (1 slice, 40 KB)", 5120, 4, 5120, 0, 1},
      {"13B B=1 K=5120  4x4096 (old: 4096+1024)", 5120, 4, 4096, 0, 1},
      {"13B B=1 K=13824 4x4608", 13824, 4, 4608, 0, 1},          {"13B B=1 K=13824 4x3456", 13824, 4, 3456, 0, 1},
      {"13B B=4 K=5120  8x2048 (old: 2048+2048+1024)", 5120, 8, 2048, 64, 4}, {"13B B=4 K=5120  8x2560", 5120, 8, 2560, 64, 4},
      {"13B B=4 K=5120  4x2560", 5120, 4, 2560, 64, 4},          {"13B B=4 K=5120  8x1728", 5120, 8, 1728, 64, 4},
      {"13B B=4 K=13824 8x2048 (old)", 13824, 8, 2048, 64, 4},   {"13B B=4 K=13824 8x2304", 13824, 8, 2304, 64, 4},
      {"13B B=4 K=13824 8x1728", 13824, 8, 1728, 64, 4},         {"13B B=4 K=13824 4x2304", 13824, 4, 2304, 64, 4},
      {"7B  B=4 K=4096  8x2048 (old)", 4096, 8, 2048, 64, 4},
      {"13B B=4 K=5120  8x2560 (x not in smem)", 5120, 8, 2560, 64, 0},   {"13B B=4 K=5120  16x1280 (weights as A)", 5120, 16, 1280, 64, 0},
      {"13B B=4 K=5120  16x2560 (weights as A, 82 KB)", 5120, 16, 2560, 64, 0}, {"13B B=4 K=13824 8x2304 (x not in smem)", 13824, 8, 2304, 64, 0},
  };
  for (const Cfg& c : cfgs) {
    const int stage_b = c.ROWS * (c.KC * 2 + c.pad);
    const int N = (int)(BYTES / ((size_t)c.K * 2) / (148 * c.ROWS) * (148 * c.ROWS));
    const int kmax = c.K >= 8192 ? c.K : (c.K == 4096 ? 11008 : 13824);      // the activation block is sized for the model's largest K
    const int xs_stride = (((kmax * 2 + 127) & ~127) + c.pad) / 2;
    const size_t x_bytes = (size_t)c.xrows * xs_stride * 2, x_budget = x_bytes;
    printf("%s  stage %d B\n", c.name, stage_b);
    const int modes_b1[] = {0, 1, 9}, modes_b4[] = {0, 2, 4, 5, 6, 14}, modes_x0[] = {0, 4, 6, 14}, modes_a[] = {0, 7, 15};
    const int* modes = !c.pad ? modes_b1 : (c.ROWS == 16 ? modes_a : (c.xrows == 0 ? modes_x0 : modes_b4));
    const int n_modes = !c.pad ? 3 : (c.ROWS == 16 ? 3 : (c.xrows == 0 ? 4 : 6));
    for (int stages = 2; stages <= 8; ++stages) {
      if ((size_t)stages * stage_b + x_budget + 9 * 1024 > 227 * 1024) break;          // what would fit next to the real kernel's other buffers
      const size_t smem = (size_t)stages * stage_b + x_bytes + 256 + 8192 + 128;
      if (smem > 227 * 1024) break;
      if (quick && stages != 3 && stages != 5) continue;
      for (int infl = 3; infl <= stages; ++infl) {
        if (infl != stages && infl != 3 && infl != 4) continue;
        if (quick && infl != stages && infl != 3) continue;
        printf("    stages %d inflight %d:", stages, infl);
        for (int mi = 0; mi < n_modes; ++mi) {
          RP p;
          p.W = (const __nv_bfloat16*)d; p.N = N; p.K = c.K; p.ROWS = c.ROWS; p.KC = c.KC; p.pad = c.pad; p.n_stages = stages; p.n_inflight = infl;
          p.mode = modes[mi]; p.xs_stride = xs_stride; p.xrows = c.xrows; p.out = (float*)d;
          float best = 0.f;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            ring_stream<<<148, 576, smem>>>(p);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const float gbs = (float)((double)N * c.K * 2 / ms / 1e6);
            if (rep > 0 && gbs > best) best = gbs;
          }
          printf("  mode %2d %7.1f", modes[mi], best);
        }
        printf("  GB/s (%s)\n", cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
