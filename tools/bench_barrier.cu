// Grid-barrier micro-benchmark (experiment tool for the decode step kernel; not on the product path).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_barrier tools/bench_barrier.cu && tools/bench_barrier
// One CTA per SM, 544 participating threads like decode_step_kernel's consumers.  Variants:
//   flat      : every CTA does red.release on ONE counter, thread 0 polls it with ld.acquire       (what decode_mega.cuh does)
//   hier      : CTAs arrive on one of G group counters (atom, last arriver forwards to the root), everyone polls the root
//   cluster2/4: barrier.cluster inside a 2-/4-CTA cluster, ONE red + poll per cluster, barrier.cluster to release the rest
// Prints microseconds per barrier (N barriers in one cooperative launch, max over CTAs of clock64 deltas / SM clock).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t ld_acq(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_rel(unsigned int* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory"); }
__device__ __forceinline__ uint32_t atom_acq_rel(unsigned int* p) {
  uint32_t v;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kThreads = 544;

// mode 0 flat, 1 hierarchical (groups of `gsz` CTAs)
__global__ void __launch_bounds__(kThreads, 1) barrier_kernel(unsigned int* counters, int n_iter, int mode, int gsz, long long* cycles) {
  const int tid = threadIdx.x;
  unsigned int* root = counters;                 // [0]
  unsigned int* grp = counters + 32 + (blockIdx.x / gsz) * 32;   // one 128-byte line per group
  const int n_groups = (gridDim.x + gsz - 1) / gsz;
  const int my_group_size = min(gsz, (int)gridDim.x - (int)(blockIdx.x / gsz) * gsz);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 1; it <= n_iter; ++it) {
    __syncthreads();
    if (tid == 0) {
      if (mode == 0) {
        red_rel(root);
        while ((int)(ld_acq(root) - (unsigned)it * gridDim.x) < 0) {
        }
      } else {
        const uint32_t prev = atom_acq_rel(grp);
        if ((prev + 1) % (unsigned)my_group_size == 0) red_rel(root);      // last arriver of the group forwards
        while ((int)(ld_acq(root) - (unsigned)it * n_groups) < 0) {
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

// cluster-assisted: one global arrival per cluster
template <int CS>
__global__ void __launch_bounds__(kThreads, 1) barrier_cluster_kernel(unsigned int* counters, int n_iter, long long* cycles) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  unsigned int* root = counters;
  const unsigned n_clusters = gridDim.x / CS;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 1; it <= n_iter; ++it) {
    cl.sync();                                            // everyone in the cluster has arrived (and its writes are visible cluster-wide)
    if (cl.block_rank() == 0 && tid == 0) {
      __threadfence();
      red_rel(root);
      while ((int)(ld_acq(root) - (unsigned)it * n_clusters) < 0) {
      }
    }
    cl.sync();                                            // release the rest of the cluster
  }
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

static double run(int mode, int gsz, int cs, int grid, int n_iter, unsigned int* d_cnt, long long* d_cyc, double sm_ghz) {
  CK(cudaMemset(d_cnt, 0, 4096 * 4));
  if (cs == 0) {
    void* args[] = {&d_cnt, &n_iter, &mode, &gsz, &d_cyc};
    CK(cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(grid), dim3(kThreads), args, 0, 0));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid / cs * cs);
    cfg.blockDim = dim3(kThreads);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    if (cs == 2) CK(cudaLaunchKernelEx(&cfg, barrier_cluster_kernel<2>, d_cnt, n_iter, d_cyc));
    else CK(cudaLaunchKernelEx(&cfg, barrier_cluster_kernel<4>, d_cnt, n_iter, d_cyc));
  }
  CK(cudaDeviceSynchronize());
  static long long h[1024];
  CK(cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx / n_iter / (sm_ghz * 1e3);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int grid = p.multiProcessorCount;
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz / 1e6;
  unsigned int* d_cnt;
  long long* d_cyc;
  CK(cudaMalloc(&d_cnt, 4096 * 4));
  CK(cudaMalloc(&d_cyc, 1024 * 8));
  const int n_iter = 2000;
  printf("%s, %d SMs, %.3f GHz, %d barriers per launch\n", p.name, grid, ghz, n_iter);
  for (int rep = 0; rep < 2; ++rep) {
    printf("flat                : %.3f us\n", run(0, 1, 0, grid, n_iter, d_cnt, d_cyc, ghz));
    for (int g : {4, 8, 16, 37}) printf("hier (groups of %2d) : %.3f us\n", g, run(1, g, 0, grid, n_iter, d_cnt, d_cyc, ghz));
    printf("cluster of 2        : %.3f us\n", run(0, 1, 2, grid, n_iter, d_cnt, d_cyc, ghz));
    printf("cluster of 4        : %.3f us\n", run(0, 1, 4, grid, n_iter, d_cnt, d_cyc, ghz));
  }
  return 0;
}
