// libvalley_b200.so -- host side of the C ABI declared in include/valley_b200.h.
// Owns: packed weights, workspace, KV caches, CUDA-graph of the decode step.  No torch, no CPU fallback.
#include <cuda_runtime.h>
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/valley_b200.h"
#include "attention_tc.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "simt_kernels.cuh"
#include "decode_kernels.cuh"
#include "decode_mega.cuh"
#include "decode_umma.cuh"
#include "preprocess.cuh"
#include "pooling_kernels.cuh"

using namespace vly;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
extern "C" void vly_set_error_(const char* m) { snprintf(g_err, sizeof(g_err), "%s", m); }
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* vly_last_error(void) { return g_err; }
extern "C" const char* vly_version(void) { return "valley_b200 0.1 (sm_100a)"; }

#define CK(expr)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return fail(VLY_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
  } while (0)
#define CKL() CK(cudaGetLastError())
#define TRY(expr)              \
  do {                         \
    int r_ = (expr);           \
    if (r_ != VLY_OK) return r_; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// data structures
// ------------------------------------------------------------------------------------------------
struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};
static int ensure(Buf& b, size_t bytes) {
  if (b.bytes >= bytes) return VLY_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  CK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return VLY_OK;
}

struct Staged {
  std::vector<int64_t> shape;
  void* dev = nullptr;
  bool is_f32 = false;  // vectors are kept as fp32 holding bf16-rounded values; matrices as bf16
  int64_t numel = 0;
};

struct VitLayerW {
  bf16 *wqkv, *wo, *w1, *w2;
  float *qkv_cs, *qkv_b, *bo, *c1, *b1, *b2;
};
struct LlamaLayerW {
  bf16 *wqkv, *wo, *wgu, *wdown;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct vly_ctx {
  vly_config cfg;
  int num_sms = 148;
  PFN_encodeTiled encode = nullptr;
  std::mutex mu;
  std::map<std::string, Staged> staged;
  bool finalized = false;
  bool has_vit = false, has_llm = false;
  std::vector<void*> owned;  // every packed weight allocation
  // ViT
  int kpad = 0;
  bf16* patch_w = nullptr;
  float *cls = nullptr, *pos = nullptr, *pre_g = nullptr, *pre_b = nullptr;
  std::vector<VitLayerW> vit;
  bf16* proj_w = nullptr;
  float* proj_b = nullptr;
  // LLaMA
  bf16* embed = nullptr;
  std::vector<LlamaLayerW> layers;
  bf16* lm_head = nullptr;
  float2* rope = nullptr;
  // workspace
  Buf w_col, w_patch, w_qkv, w_ctx, w_h, w_stats, w_pool, w_x, w_q, w_attn, w_hb, w_pstats;
  cudaStream_t cap_stream = nullptr;
  int64_t launches = 0;
  // fused all-gather state
  bf16* g_buf = nullptr;          // [g_rows, vit_hidden] + flags
  int64_t g_rows = 0;
  int* g_flags = nullptr;
  bf16* g_peer_buf[8] = {};
  int* g_peer_flags[8] = {};
  int g_world = 0, g_rank = 0, g_epoch = 0;
  int* g_timeout = nullptr;       // pinned + mapped: set by gather_wait_kernel when a peer never signalled; read by the host
  Buf w_xlocal;
  // pooling variants (valley_model.py:40-52, :205-213)
  float* pool_U = nullptr;                     // temporal_importance: W_proj^T w_pool, [256, vit_hidden] fp32
  struct DeltaW {                              // temporal_transformer: one post-LN nn.TransformerEncoderLayer + position_matrix
    bf16 *in_w = nullptr, *out_w = nullptr, *l1_w = nullptr, *l2_w = nullptr, *pos = nullptr;
    float *in_b = nullptr, *out_b = nullptr, *l1_b = nullptr, *l2_b = nullptr, *n1_g = nullptr, *n1_b = nullptr, *n2_g = nullptr, *n2_b = nullptr;
    int ffn = 0, max_pos = 0;
  } delta;
  Buf w_score, w_pall, w_xp, w_dkv, w_dq, w_datt, w_dx1, w_df1, w_dx2;
  // frame preprocessing: strip + coefficient tables of the last geometry seen
  Buf w_strip, w_tables;
  int pre_H = 0, pre_W = 0;
  PreprocParams pre = {};
};

struct vly_kv;
static int sync_len(vly_kv* kv);
struct vly_kv {
  vly_ctx* ctx;
  int B, Smax;
  bf16* cache = nullptr;  // [L][2][B][nH][Smax][128]
  int host_len = 0;
  bool len_dirty = false;   // a stop token may have ended vly_generate early: host_len is re-read from the device on next use
  int* h_len = nullptr;     // pinned: d_len is copied here on the generating stream, len_event marks the copy
  cudaEvent_t len_event = nullptr;
  int* d_len = nullptr;   // device scalar
  int* d_step = nullptr;
  // decode workspace
  bf16 *x = nullptr, *q = nullptr, *attn = nullptr, *hb = nullptr;
  float* part_o = nullptr;
  float2* part_ml = nullptr;
  unsigned int* counters = nullptr;   // [B*nH] + 1 (argmax)
  float* part_val = nullptr;
  int* part_idx = nullptr;
  float* logits = nullptr;            // [B, V]
  long long* cur_tokens = nullptr;    // [B]
  long long* gen_tokens = nullptr;    // [B, Smax]
  int nsplit = 1, gemv_grid = 0;
  int stage_bytes = 0;                 // ring slot of the persistent decode kernel: max over its phases (pick_phase_geometry)
  int n_grid_syncs = 0;                // grid barriers per decode launch
  bool umma = false;                   // B = 2..4 on the tcgen05 consumer (decode_umma.cuh)
  int umma_x_cols = 0;                 // columns of the swizzled activation block of that kernel
  void* d_tmaps = nullptr;             // CUtensorMap[] of the weight matrices (one per phase geometry)
  PhaseDesc* d_phases = nullptr;
  int n_phases = 0;
  unsigned int* grid_counter = nullptr;
  long long* dbg = nullptr;
  SampleState* d_sample = nullptr;    // token selection state read by every decode step (sampling.cuh)
  bool sample_dirty = false;          // device state is not the plain-greedy default
  uint32_t* key_bits = nullptr;       // [B, Smax/32] attention_mask bits (1 = attend); all ones unless vly_kv_set_key_mask
  bool masked = false;
  int mask_words() const { return Smax / 32; }
  cudaGraphExec_t graph = nullptr;     // one decode step
  cudaGraphExec_t graph_n = nullptr;   // kGraphSteps steps in one graph (fewer graph launches, kernel->kernel edges inside)
  int graph_nodes = 0;
  size_t layer_stride() const { return (size_t)2 * B * ctx->cfg.num_attention_heads * Smax * 128; }
  bf16* k_layer(int l) const { return cache + (size_t)l * layer_stride(); }
  bf16* v_layer(int l) const { return k_layer(l) + layer_stride() / 2; }
};

// after an eos-terminated vly_generate only the device knows how many steps ran (one blocking 4-byte read, off the hot path)
// (the copy was enqueued on the stream that ran the generation -- torch streams are non-blocking, so a legacy-stream
//  cudaMemcpy here would not be ordered after it)
static int sync_len(vly_kv* kv) {
  if (!kv->len_dirty) return VLY_OK;
  CK(cudaSetDevice(kv->ctx->cfg.device));
  CK(cudaEventSynchronize(kv->len_event));
  kv->host_len = *reinterpret_cast<volatile int*>(kv->h_len);
  kv->len_dirty = false;
  return VLY_OK;
}

// decode implementation: 2 = one persistent cooperative kernel per step (default), 1 = per-op TMA-ring kernels + PDL,
// 0 = per-op register-streaming kernels (generation 1).  VLY_DECODE=v1|ring|mega overrides (A/B measurements).
static int decode_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLY_DECODE");
    v = 2;
    if (e && !strcmp(e, "v1")) v = 0;
    if (e && !strcmp(e, "ring")) v = 1;
    if (getenv("VLY_DECODE_V1")) v = 0;
  }
  return v;
}
static bool use_decode_v1() { return decode_mode() == 0; }

static inline int cdiv(long long a, long long b) { return int((a + b - 1) / b); }

// cudaFuncSetAttribute is per DEVICE: a second vly_ctx on another GPU of the same process must opt in again, so what has been
// set is tracked per (device, kernel) -- not in function-local statics.
static std::mutex g_attr_mu;
static std::map<std::pair<int, const void*>, size_t> g_attr_set;
template <typename Kern>
static int ensure_smem_attr(int device, Kern kern, size_t bytes, bool max_carveout = false) {
  std::lock_guard<std::mutex> lk(g_attr_mu);
  size_t& have = g_attr_set[std::make_pair(device, (const void*)kern)];
  if (bytes > have || (have == 0 && max_carveout)) {
    if (bytes > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (max_carveout) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    have = bytes > 0 ? bytes : 1;
  }
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------------
static int make_tmap_2d(vly_ctx* c, CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_rows, bool swizzle128 = true) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VLY_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d (ptr=%p inner=%llu rows=%llu stride=%llu box=%u,%u)", (int)r, ptr,
                (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner, box_rows);
  return VLY_OK;
}
static int make_tmap_3d(vly_ctx* c, CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                        uint32_t b0, uint32_t b1) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VLY_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return VLY_OK;
}

// ---- launch helper: optional programmatic dependent launch (the kernel's prologue overlaps the previous kernel's tail; kernels
// launched this way call griddepcontrol.wait before they touch the previous kernel's data).  VLY_NO_PDL=1 disables it. ----
static bool pdl_enabled() {
  static const bool on = getenv("VLY_NO_PDL") == nullptr;
  return on;
}
template <typename Kern, typename... Args>
static cudaError_t launch_ex(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ------------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------------
// TMA-store epilogue (gemm_tc.cuh, TEPI): bf16 row-major outputs whose rows the TMA unit can address (16-byte aligned base and pitch)
template <int EPI>
static bool gemm_tepi_ok(const GemmParams& p) {
  static const int env = getenv("VLY_GEMM_TEPI") ? atoi(getenv("VLY_GEMM_TEPI")) : 1;
  if (!env) return false;
  // (measured: it pays where the epilogue, not the main loop, sets the pace -- K <= 2048: ViT F = 64 10.24 -> 9.47 ms; with the
  //  LLaMA widths, K = 4096 .. 13824, the 5-stage ring it needs costs more than the stores save: 13B prefill 46.6 -> 47.4 ms)
  // ... except the residual epilogue of a LONG K = 4096 GEMM with few column tiles (the ViT's fc2 at >= 48 frames: 9.59 -> 9.20 ms at
  // 64 frames), where the row-wise residual loads + stores still rival the main loop
  if (p.K > 2048 && env < 2 && !(EPI == EPI_BIAS_RES_STATS && p.K <= 4096 && p.M >= 12000)) return false;
  // The GEMM that also pushes its tiles to the peers (fused all-gather) keeps the register-store epilogue: that combination is the one
  // verified bit-identical to NCCL on 2 AND 8 GPUs.  With the staged epilogue it was bit-identical on 2 GPUs, but the one 8-GPU
  // run taken with it reported a mismatch (profiles/bench_r02_n8_tepi_gather_mismatch.json).  Suspected: the residual-box
  // refill in the staged epilogue is issued before the ld.shared of that box are known to have retired; behind 7 peers' NVLink stores
  // they can still be queued when the TMA write lands.  To be fixed and re-verified on 8 GPUs before this is switched on.
  if (p.n_peers > 0 && env < 2) return false;
  if (!(EPI == EPI_BIAS || EPI == EPI_LN_BIAS || EPI == EPI_LN_BIAS_GELU || EPI == EPI_BIAS_RES_STATS || EPI == EPI_RMS_SWIGLU)) return false;
  if ((p.ldo & 7) || (reinterpret_cast<uintptr_t>(p.out) & 15)) return false;
  if (EPI == EPI_BIAS_RES_STATS && ((p.ldr & 7) || (reinterpret_cast<uintptr_t>(p.residual) & 15))) return false;
  return true;
}

template <int BN, int EPI>
static int launch_gemm_t(vly_ctx* c, const bf16* A, long long lda, const bf16* W, long long ldw, GemmParams p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  constexpr bool kTepiMode = (EPI == EPI_BIAS || EPI == EPI_LN_BIAS || EPI == EPI_LN_BIAS_GELU || EPI == EPI_BIAS_RES_STATS || EPI == EPI_RMS_SWIGLU);
  CUtensorMap ta, tb, to, tr;
  TRY(make_tmap_2d(c, &ta, A, p.K, p.M, lda * 2, 64, 128));
  TRY(make_tmap_2d(c, &tb, W, p.K, p.N, ldw * 2, 64, BN));
  to = ta; tr = ta;                                        // (placeholders when the epilogue stores from registers)
  const bool tepi = gemm_tepi_ok<EPI>(p);
  if (tepi) {
    const int n_out = (EPI == EPI_RMS_SWIGLU) ? p.N / 2 : p.N;
    TRY(make_tmap_2d(c, &to, p.out, n_out, p.M, p.ldo * 2, 64, 32));
    if (EPI == EPI_BIAS_RES_STATS) TRY(make_tmap_2d(c, &tr, p.residual, p.N, p.M, p.ldr * 2, 64, 32));
  }
  p.num_m_tiles = cdiv(p.M, 128);
  p.num_n_tiles = cdiv(p.N, BN);
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  // CTA-pair mode (cta_group::2, 256 x 256 tile per pair): halves the B traffic per SM (64 instead of 96 B/cycle/SM of
  // L2->SM operand traffic).  Used when there is enough work to fill the pairs; VLY_GEMM_CG2=0/1 overrides.
  if constexpr (BN == 256) {
    static const int cg2_env = getenv("VLY_GEMM_CG2") ? atoi(getenv("VLY_GEMM_CG2")) : -1;
    const int pairs = c->num_sms / 2;
    const int pair_tiles = cdiv(p.num_m_tiles, 2) * p.num_n_tiles;
    // (measured: pairs win from one full wave of pair tiles on -- ViT, 32 frames: 5.23 -> 5.10 ms; below that -- 16 frames, 68 tiles of
    //  the N = 1024 GEMMs -- the single-CTA tiles spread better: 3.05 vs 3.09 ms)
    //  An odd, small number of 128-row tiles wastes the pair's second half (M = 333: 3 tiles -> 4): those wait for two full waves.
    const bool small_odd = (p.num_m_tiles & 1) && p.num_m_tiles < 7;
    const bool use_cg2 = cg2_env >= 0 ? (cg2_env == 1) : (pair_tiles >= (small_odd ? 2 : 1) * pairs);
    if (use_cg2) {
      CUtensorMap tb2;
      TRY(make_tmap_2d(c, &tb2, W, p.K, p.N, ldw * 2, 64, BN / 2));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (pair_tiles < pairs ? pair_tiles : pairs));
      cfg.blockDim = dim3(Cfg::THREADS);
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled() ? 2 : 1;
      if constexpr (kTepiMode) {
        if (tepi) {
          constexpr int smem = gemm_smem_bytes<BN, true>();
          TRY(ensure_smem_attr(c->cfg.device, gemm_tc_kernel<BN, EPI, true, true>, smem));
          cfg.dynamicSmemBytes = smem;
          CK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, true, true>, ta, tb2, to, tr, p));
          c->launches++;
          return VLY_OK;
        }
      }
      TRY(ensure_smem_attr(c->cfg.device, gemm_tc_kernel<BN, EPI, true, false>, Cfg::SMEM_BYTES));
      cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      CK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, true, false>, ta, tb2, to, tr, p));
      c->launches++;
      return VLY_OK;
    }
  }
  const int grid = tiles < c->num_sms ? tiles : c->num_sms;
  if constexpr (kTepiMode && BN == 128) {
    if (tepi) {
      constexpr int smem = gemm_smem_bytes<BN, true>();
      TRY(ensure_smem_attr(c->cfg.device, gemm_tc_kernel<BN, EPI, false, true>, smem));
      CK(launch_ex(gemm_tc_kernel<BN, EPI, false, true>, dim3(grid), dim3(Cfg::THREADS), smem, st, true, ta, tb, to, tr, p));
      c->launches++;
      return VLY_OK;
    }
  }
  TRY(ensure_smem_attr(c->cfg.device, gemm_tc_kernel<BN, EPI, false, false>, Cfg::SMEM_BYTES));
  CK(launch_ex(gemm_tc_kernel<BN, EPI, false, false>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, true, ta, tb, to, tr, p));
  c->launches++;
  return VLY_OK;
}
template <int EPI>
static int launch_gemm(vly_ctx* c, int bn, const bf16* A, long long lda, const bf16* W, long long ldw, const GemmParams& p,
                       cudaStream_t st) {
  if (bn == 256) return launch_gemm_t<256, EPI>(c, A, lda, W, ldw, p, st);
  return launch_gemm_t<128, EPI>(c, A, lda, W, ldw, p, st);
}
static inline int pick_bn(int N) { return (N % 256 == 0) ? 256 : 128; }
// Tile width by wave efficiency: a persistent launch runs ceil(tiles / SMs) rounds, so what counts is how full the last round
// is.  M = 2056 (8 frames), N = 3072: 128 x 256 tiles -> 204 tiles = 2 rounds at 69 %; 128 x 128 -> 408 tiles = 3 half-size rounds
// at 92 %.  256-wide tiles are kept when they are within 5 % (better operand reuse per SM).  M-dependent, so callers that
// exchange row statistics compute it ONCE per (N, M) and pass it around.
static inline int pick_bn_m(const vly_ctx* c, int N, int M) {
  if (N % 256 != 0) return 128;
  static const int env_bn = getenv("VLY_GEMM_BN") ? atoi(getenv("VLY_GEMM_BN")) : 0;      // 128 / 256: force (A/B measurements)
  if (env_bn == 128 || env_bn == 256) return env_bn;
  const long long t256 = (long long)cdiv(M, 128) * (N / 256), t128 = 2 * t256;
  if (t256 >= 3LL * c->num_sms) return 256;            // many rounds: the last one matters little, operand reuse matters more
  {
    // CTA pairs (256 x 256 tiles, launch_gemm_t) sustain ~1.5x the rate of single 128 x 128 tiles (ncu, LLaMA-13B prefill: 1.22 vs
    // 0.82 PFLOP/s), so a reasonably full pair launch beats a perfectly full 128-wide one: M = 1332, N = 5120 -- 120 pair tiles on 74
    // pairs, 11 of 12 row tiles real = 0.74 -- took 3.2 ms off the 13B prefill (47.9 -> 44.7).  Below ~0.7 (8 ViT frames: 0.69) the
    // narrow tiles stay.
    const int mt = cdiv(M, 128), pairs = c->num_sms / 2;
    const long long pt = (long long)cdiv(mt, 2) * (N / 256);
    const bool small_odd = (mt & 1) && mt < 7;
    if (pt >= (small_odd ? 2 : 1) * pairs) {
      const double e_pair = (double)pt / ((double)cdiv(pt, pairs) * pairs) * ((double)mt / (2.0 * cdiv(mt, 2)));
      if (e_pair >= 0.70) return 256;
    }
  }
  const double e256 = (double)t256 / ((double)cdiv(t256, c->num_sms) * c->num_sms);
  const double e128 = (double)t128 / ((double)cdiv(t128, c->num_sms) * c->num_sms);
  return (e128 > e256 * 1.05) ? 128 : 256;
}

// ------------------------------------------------------------------------------------------------
// weight packing kernels
// ------------------------------------------------------------------------------------------------
// One CTA per SOURCE row r.  dst row: mode 0 -> off + r; mode 1 (RoPE pair interleave inside 128-wide heads) ->
// off + h*128 + (d < 64 ? 2d : 2(d-64)+1); mode 2 (gate/up interleave) -> 2r + off.
__global__ void pack_rows_kernel(const bf16* __restrict__ src, int K, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ bias_in, bf16* __restrict__ dst, int Kdst, int mode, int off,
                                 float* __restrict__ colsum, float* __restrict__ bias_out) {
  __shared__ float r1[8], r2[8];
  const int r = blockIdx.x;
  int dr;
  if (mode == 0) dr = off + r;
  else if (mode == 1) {
    const int h = r >> 7, d = r & 127;
    dr = off + h * 128 + (d < 64 ? 2 * d : 2 * (d - 64) + 1);
  } else dr = 2 * r + off;
  float cs = 0.f, bb = 0.f;
  for (int k = threadIdx.x; k < Kdst; k += blockDim.x) {
    float wf = 0.f;
    if (k < K) {
      const float w = __bfloat162float(src[(size_t)r * K + k]);
      wf = gamma ? bf16_round(w * gamma[k]) : w;
      if (beta) bb += w * beta[k];
    }
    dst[(size_t)dr * Kdst + k] = __float2bfloat16_rn(wf);
    cs += wf;
  }
  cs = warp_sum(cs);
  bb = warp_sum(bb);
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = cs;
    r2[threadIdx.x >> 5] = bb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += r1[i];
      b += r2[i];
    }
    if (colsum) colsum[dr] = a;
    if (bias_out) bias_out[dr] = (bias_in ? bias_in[r] : 0.f) + b;
  }
}

// rope[pos, j] = (cos, sin) of pos * theta^(-2j/128), computed in fp32 like HF (modeling_llama.py:124-135) and rounded
// to bf16 (cos.to(x.dtype)).
__global__ void rope_table_kernel(float2* rope, int max_pos, float theta, int head_dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = head_dim / 2;
  if (i >= max_pos * half) return;
  const int pos = i / half, j = i % half;
  const float inv = 1.0f / powf(theta, float(2 * j) / float(head_dim));
  const float fr = float(pos) * inv;
  rope[i] = make_float2(bf16_round(cosf(fr)), bf16_round(sinf(fr)));
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

// inputs_embeds -> x (copy) + row statistics
__global__ void __launch_bounds__(128) copy_rows_stats_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                              float2* __restrict__ stats, int stats_nt, int H) {
  __shared__ float red[4];
  const int row = blockIdx.x;
  float s = 0.f, sq = 0.f;
  for (int c = threadIdx.x * 8; c < H; c += 128 * 8) {
    const uint4 w = *reinterpret_cast<const uint4*>(in + (size_t)row * H + c);
    *reinterpret_cast<uint4*>(out + (size_t)row * H + c) = w;
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = bf16_lo(ww[i]), b = bf16_hi(ww[i]);
      s += a + b;
      sq += a * a + b * b;
    }
  }
  s = block_sum_128(s, red);
  sq = block_sum_128(sq, red);
  if (threadIdx.x < stats_nt) stats[(size_t)row * stats_nt + threadIdx.x] = threadIdx.x == 0 ? make_float2(s, sq) : make_float2(0.f, 0.f);
}

// cache [B,nH,Smax,128] -> HF layout [B,nH,len,128]; K is stored RoPE-pair-interleaved and is de-interleaved here.
__global__ void kv_export_kernel(const bf16* __restrict__ cache, bf16* __restrict__ out, int Smax, int len, int deinterleave) {
  const int bh = blockIdx.y, s = blockIdx.x, c = threadIdx.x;  // 128 threads
  const int d = deinterleave ? ((c & 1) ? 64 + (c >> 1) : (c >> 1)) : c;
  out[((size_t)bh * len + s) * 128 + d] = cache[((size_t)bh * Smax + s) * 128 + c];
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
extern "C" int vly_create(const vly_config* cfg, vly_ctx** out) {
  if (!cfg || !out) return fail(VLY_ERR_INVALID, "vly_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(VLY_ERR_CUDA, "vly_create: no CUDA device visible -- this library has no CPU fallback");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(VLY_ERR_CUDA, "vly_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  if (cfg->hidden_size % 128 || cfg->hidden_size / cfg->num_attention_heads != 128)
    return fail(VLY_ERR_INVALID, "vly_create: head_dim must be 128 (hidden %d, heads %d)", cfg->hidden_size, cfg->num_attention_heads);
  if (cfg->intermediate_size % 64) return fail(VLY_ERR_INVALID, "vly_create: intermediate_size must be a multiple of 64");
  if (cfg->vit_hidden != 1024 || cfg->vit_hidden / cfg->vit_heads != 64 || cfg->vit_mlp % 256)
    return fail(VLY_ERR_INVALID, "vly_create: vision tower must be ViT-L width (1024, head_dim 64); the reference hard-codes 1024 (valley_model.py:192)");
  vly_ctx* c = new vly_ctx();
  c->cfg = *cfg;
  c->num_sms = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
    delete c;
    return fail(VLY_ERR_CUDA, "vly_create: cuTensorMapEncodeTiled entry point not found");
  }
  c->encode = (PFN_encodeTiled)fn;
  if (cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return fail(VLY_ERR_CUDA, "vly_create: cudaStreamCreate failed");
  }
  *out = c;
  return VLY_OK;
}

extern "C" void vly_destroy(vly_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  for (auto& kvp : c->staged) cudaFree(kvp.second.dev);
  for (void* p : c->owned) cudaFree(p);
  Buf* bufs[] = {&c->w_col, &c->w_patch, &c->w_qkv, &c->w_ctx, &c->w_h, &c->w_stats, &c->w_pool, &c->w_x, &c->w_q, &c->w_attn, &c->w_hb, &c->w_pstats,
                 &c->w_xlocal, &c->w_score, &c->w_pall, &c->w_xp, &c->w_dkv, &c->w_dq, &c->w_datt, &c->w_dx1, &c->w_df1, &c->w_dx2, &c->w_strip,
                 &c->w_tables};
  for (Buf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->g_timeout) cudaFreeHost(c->g_timeout);
  delete c;
}

extern "C" int vly_num_sms(vly_ctx* c, int* out) {
  if (!c || !out) return fail(VLY_ERR_INVALID, "null");
  *out = c->num_sms;
  return VLY_OK;
}
extern "C" int vly_kernel_launch_count(vly_ctx* c, int64_t* out) {
  if (!c || !out) return fail(VLY_ERR_INVALID, "null");
  *out = c->launches;
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
static bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

extern "C" int vly_load_weight(vly_ctx* c, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim) {
  if (!c || !name || !dev_ptr || !shape || ndim < 1 || ndim > 4) return fail(VLY_ERR_INVALID, "vly_load_weight: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->finalized) return fail(VLY_ERR_STATE, "vly_load_weight(%s): weights already finalised", name);
  CK(cudaSetDevice(c->cfg.device));
  Staged s;
  s.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    s.shape.push_back(shape[i]);
    s.numel *= shape[i];
  }
  const std::string nm(name);
  s.is_f32 = (ndim == 1) || ends_with(nm, "position_embedding.weight");
  auto it = c->staged.find(nm);
  if (it != c->staged.end()) {
    cudaFree(it->second.dev);
    c->staged.erase(it);
  }
  CK(cudaMalloc(&s.dev, (size_t)s.numel * (s.is_f32 ? 4 : 2)));
  const int blocks = (int)std::min<long long>((s.numel + 255) / 256, 4096);
  if (s.is_f32) {
    if (dtype == VLY_F32) convert_to_f32_bf16rounded_kernel<float><<<blocks, 256>>>((const float*)dev_ptr, (float*)s.dev, s.numel);
    else if (dtype == VLY_BF16) convert_to_f32_bf16rounded_kernel<bf16><<<blocks, 256>>>((const bf16*)dev_ptr, (float*)s.dev, s.numel);
    else if (dtype == VLY_F16) convert_to_f32_bf16rounded_kernel<__half><<<blocks, 256>>>((const __half*)dev_ptr, (float*)s.dev, s.numel);
    else return fail(VLY_ERR_INVALID, "vly_load_weight(%s): unknown dtype %d", name, dtype);
  } else {
    if (dtype == VLY_F32) convert_to_bf16_kernel<float><<<blocks, 256>>>((const float*)dev_ptr, (bf16*)s.dev, s.numel);
    else if (dtype == VLY_BF16) CK(cudaMemcpyAsync(s.dev, dev_ptr, (size_t)s.numel * 2, cudaMemcpyDeviceToDevice, 0));
    else if (dtype == VLY_F16) convert_to_bf16_kernel<__half><<<blocks, 256>>>((const __half*)dev_ptr, (bf16*)s.dev, s.numel);
    else return fail(VLY_ERR_INVALID, "vly_load_weight(%s): unknown dtype %d", name, dtype);
  }
  CKL();
  CK(cudaStreamSynchronize(0));  // the caller may free its tensor right after we return
  c->staged[nm] = s;
  return VLY_OK;
}

template <typename T>
static int dalloc(vly_ctx* c, T** p, size_t n) {
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  c->owned.push_back(*p);
  return VLY_OK;
}

static int get_staged(vly_ctx* c, const std::string& name, bool f32, int64_t numel, void** out) {
  auto it = c->staged.find(name);
  if (it == c->staged.end()) return fail(VLY_ERR_STATE, "vly_finalize_weights: missing tensor '%s'", name.c_str());
  if (it->second.is_f32 != f32 || it->second.numel != numel)
    return fail(VLY_ERR_INVALID, "vly_finalize_weights: tensor '%s' has %lld elements (expected %lld)", name.c_str(),
                (long long)it->second.numel, (long long)numel);
  *out = it->second.dev;
  return VLY_OK;
}
static void drop_staged(vly_ctx* c, const std::string& name) {
  auto it = c->staged.find(name);
  if (it != c->staged.end()) {
    cudaFree(it->second.dev);
    c->staged.erase(it);
  }
}

extern "C" int vly_finalize_weights(vly_ctx* c) {
  if (!c) return fail(VLY_ERR_INVALID, "null ctx");
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->finalized) return VLY_OK;
  CK(cudaSetDevice(c->cfg.device));
  const vly_config& g = c->cfg;
  const std::string vp = "model.vision_tower.vision_model.";
  c->has_vit = c->staged.count(vp + "embeddings.patch_embedding.weight") > 0;
  c->has_llm = c->staged.count("model.embed_tokens.weight") > 0;
  if (!c->has_vit && !c->has_llm) return fail(VLY_ERR_STATE, "vly_finalize_weights: no weights loaded");

  if (c->has_vit) {
    const int D = g.vit_hidden, M = g.vit_mlp, P = g.vit_patch, KK = 3 * P * P;
    c->kpad = ((KK + 63) / 64) * 64;
    const int tokens = (g.vit_image / P) * (g.vit_image / P) + 1;
    void *pw, *cls, *pos, *pg, *pb;
    TRY(get_staged(c, vp + "embeddings.patch_embedding.weight", false, (int64_t)D * KK, &pw));
    TRY(get_staged(c, vp + "embeddings.class_embedding", true, D, &cls));
    TRY(get_staged(c, vp + "embeddings.position_embedding.weight", true, (int64_t)tokens * D, &pos));
    TRY(get_staged(c, vp + "pre_layrnorm.weight", true, D, &pg));
    TRY(get_staged(c, vp + "pre_layrnorm.bias", true, D, &pb));
    TRY(dalloc(c, &c->patch_w, (size_t)D * c->kpad));
    pack_rows_kernel<<<D, 256>>>((bf16*)pw, KK, nullptr, nullptr, nullptr, c->patch_w, c->kpad, 0, 0, nullptr, nullptr);
    CKL();
    TRY(dalloc(c, &c->cls, D));
    TRY(dalloc(c, &c->pos, (size_t)tokens * D));
    TRY(dalloc(c, &c->pre_g, D));
    TRY(dalloc(c, &c->pre_b, D));
    CK(cudaMemcpy(c->cls, cls, D * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(c->pos, pos, (size_t)tokens * D * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(c->pre_g, pg, D * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(c->pre_b, pb, D * 4, cudaMemcpyDeviceToDevice));
    drop_staged(c, vp + "embeddings.patch_embedding.weight");
    c->vit.resize(g.vit_layers);
    for (int l = 0; l < g.vit_layers; ++l) {
      const std::string q = vp + "encoder.layers." + std::to_string(l) + ".";
      VitLayerW& w = c->vit[l];
      void *g1, *b1n, *g2, *b2n;
      TRY(get_staged(c, q + "layer_norm1.weight", true, D, &g1));
      TRY(get_staged(c, q + "layer_norm1.bias", true, D, &b1n));
      TRY(get_staged(c, q + "layer_norm2.weight", true, D, &g2));
      TRY(get_staged(c, q + "layer_norm2.bias", true, D, &b2n));
      TRY(dalloc(c, &w.wqkv, (size_t)3 * D * D));
      TRY(dalloc(c, &w.qkv_cs, 3 * D));
      TRY(dalloc(c, &w.qkv_b, 3 * D));
      const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
      for (int i = 0; i < 3; ++i) {
        void *ww, *bb;
        TRY(get_staged(c, q + "self_attn." + nm[i] + ".weight", false, (int64_t)D * D, &ww));
        TRY(get_staged(c, q + "self_attn." + nm[i] + ".bias", true, D, &bb));
        pack_rows_kernel<<<D, 256>>>((bf16*)ww, D, (float*)g1, (float*)b1n, (float*)bb, w.wqkv, D, 0, i * D, w.qkv_cs, w.qkv_b);
        CKL();
        CK(cudaDeviceSynchronize());
        drop_staged(c, q + "self_attn." + nm[i] + ".weight");
      }
      void *wo, *bo, *w1, *b1, *w2, *b2;
      TRY(get_staged(c, q + "self_attn.out_proj.weight", false, (int64_t)D * D, &wo));
      TRY(get_staged(c, q + "self_attn.out_proj.bias", true, D, &bo));
      TRY(get_staged(c, q + "mlp.fc1.weight", false, (int64_t)M * D, &w1));
      TRY(get_staged(c, q + "mlp.fc1.bias", true, M, &b1));
      TRY(get_staged(c, q + "mlp.fc2.weight", false, (int64_t)D * M, &w2));
      TRY(get_staged(c, q + "mlp.fc2.bias", true, D, &b2));
      TRY(dalloc(c, &w.wo, (size_t)D * D));
      TRY(dalloc(c, &w.bo, D));
      TRY(dalloc(c, &w.w1, (size_t)M * D));
      TRY(dalloc(c, &w.c1, M));
      TRY(dalloc(c, &w.b1, M));
      TRY(dalloc(c, &w.w2, (size_t)D * M));
      TRY(dalloc(c, &w.b2, D));
      CK(cudaMemcpy(w.wo, wo, (size_t)D * D * 2, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(w.bo, bo, D * 4, cudaMemcpyDeviceToDevice));
      pack_rows_kernel<<<M, 256>>>((bf16*)w1, D, (float*)g2, (float*)b2n, (float*)b1, w.w1, D, 0, 0, w.c1, w.b1);
      CKL();
      CK(cudaMemcpy(w.w2, w2, (size_t)D * M * 2, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(w.b2, b2, D * 4, cudaMemcpyDeviceToDevice));
      CK(cudaDeviceSynchronize());
      drop_staged(c, q + "self_attn.out_proj.weight");
      drop_staged(c, q + "mlp.fc1.weight");
      drop_staged(c, q + "mlp.fc2.weight");
    }
    if (c->staged.count("model.mm_projector.weight")) {
      void *pjw, *pjb;
      TRY(get_staged(c, "model.mm_projector.weight", false, (int64_t)g.hidden_size * D, &pjw));
      TRY(get_staged(c, "model.mm_projector.bias", true, g.hidden_size, &pjb));
      TRY(dalloc(c, &c->proj_w, (size_t)g.hidden_size * D));
      TRY(dalloc(c, &c->proj_b, g.hidden_size));
      CK(cudaMemcpy(c->proj_w, pjw, (size_t)g.hidden_size * D * 2, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(c->proj_b, pjb, g.hidden_size * 4, cudaMemcpyDeviceToDevice));
      drop_staged(c, "model.mm_projector.weight");
      const int H = g.hidden_size, NP = (g.vit_image / g.vit_patch) * (g.vit_image / g.vit_patch);
      if (g.patch_pooling_method == VLY_POOL_TEMPORAL_IMPORTANCE) {       // valley_model.py:40-43
        void* pw;
        TRY(get_staged(c, "model.pooling_layer.weight", false, (int64_t)NP * H, &pw));
        TRY(dalloc(c, &c->pool_U, (size_t)NP * D));
        fold_importance_kernel<<<NP, 256>>>((const bf16*)pw, c->proj_w, c->pool_U, H, D);      // the bias cancels in the softmax
        CKL();
        CK(cudaDeviceSynchronize());
        drop_staged(c, "model.pooling_layer.weight");
      } else if (g.patch_pooling_method == VLY_POOL_TEMPORAL_TRANSFORMER) {   // valley_model.py:45-52
        const std::string q = "model.transformer_delta_encoder.layers.0.";
        vly_ctx::DeltaW& w = c->delta;
        auto it = c->staged.find(q + "linear1.weight");
        if (it == c->staged.end()) return fail(VLY_ERR_STATE, "vly_finalize_weights: missing tensor '%slinear1.weight'", q.c_str());
        w.ffn = (int)(it->second.numel / H);
        auto ip = c->staged.find("model.position_matrix");
        if (ip == c->staged.end()) return fail(VLY_ERR_STATE, "vly_finalize_weights: missing tensor 'model.position_matrix'");
        w.max_pos = (int)(ip->second.numel / H);
        struct { const char* name; bool f32; int64_t n; void** dst; } items[] = {
            {"self_attn.in_proj_weight", false, (int64_t)3 * H * H, (void**)&w.in_w}, {"self_attn.in_proj_bias", true, 3 * H, (void**)&w.in_b},
            {"self_attn.out_proj.weight", false, (int64_t)H * H, (void**)&w.out_w},   {"self_attn.out_proj.bias", true, H, (void**)&w.out_b},
            {"linear1.weight", false, (int64_t)w.ffn * H, (void**)&w.l1_w},           {"linear1.bias", true, w.ffn, (void**)&w.l1_b},
            {"linear2.weight", false, (int64_t)H * w.ffn, (void**)&w.l2_w},           {"linear2.bias", true, H, (void**)&w.l2_b},
            {"norm1.weight", true, H, (void**)&w.n1_g}, {"norm1.bias", true, H, (void**)&w.n1_b},
            {"norm2.weight", true, H, (void**)&w.n2_g}, {"norm2.bias", true, H, (void**)&w.n2_b}};
        for (auto& it2 : items) {
          void* src;
          TRY(get_staged(c, q + it2.name, it2.f32, it2.n, &src));
          const size_t bytes = (size_t)it2.n * (it2.f32 ? 4 : 2);
          CK(cudaMalloc(it2.dst, bytes));
          c->owned.push_back(*it2.dst);
          CK(cudaMemcpy(*it2.dst, src, bytes, cudaMemcpyDeviceToDevice));
          drop_staged(c, q + it2.name);
        }
        void* pm;
        TRY(get_staged(c, "model.position_matrix", false, (int64_t)w.max_pos * H, &pm));
        TRY(dalloc(c, &w.pos, (size_t)w.max_pos * H));
        CK(cudaMemcpy(w.pos, pm, (size_t)w.max_pos * H * 2, cudaMemcpyDeviceToDevice));
        drop_staged(c, "model.position_matrix");
      }
    }
  }

  if (c->has_llm) {
    const int H = g.hidden_size, I = g.intermediate_size, V = g.vocab_size, L = g.num_hidden_layers;
    void* emb;
    TRY(get_staged(c, "model.embed_tokens.weight", false, (int64_t)V * H, &emb));
    TRY(dalloc(c, &c->embed, (size_t)V * H));
    CK(cudaMemcpy(c->embed, emb, (size_t)V * H * 2, cudaMemcpyDeviceToDevice));
    drop_staged(c, "model.embed_tokens.weight");
    c->layers.resize(L);
    for (int l = 0; l < L; ++l) {
      const std::string q = "model.layers." + std::to_string(l) + ".";
      LlamaLayerW& w = c->layers[l];
      void *g1, *g2;
      TRY(get_staged(c, q + "input_layernorm.weight", true, H, &g1));
      TRY(get_staged(c, q + "post_attention_layernorm.weight", true, H, &g2));
      TRY(dalloc(c, &w.wqkv, (size_t)3 * H * H));
      const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
      for (int i = 0; i < 3; ++i) {
        void* ww;
        TRY(get_staged(c, q + "self_attn." + nm[i] + ".weight", false, (int64_t)H * H, &ww));
        pack_rows_kernel<<<H, 256>>>((bf16*)ww, H, (float*)g1, nullptr, nullptr, w.wqkv, H, i < 2 ? 1 : 0, i * H, nullptr, nullptr);
        CKL();
        CK(cudaDeviceSynchronize());
        drop_staged(c, q + "self_attn." + nm[i] + ".weight");
      }
      void *wo, *wg, *wu, *wd;
      TRY(get_staged(c, q + "self_attn.o_proj.weight", false, (int64_t)H * H, &wo));
      TRY(get_staged(c, q + "mlp.gate_proj.weight", false, (int64_t)I * H, &wg));
      TRY(get_staged(c, q + "mlp.up_proj.weight", false, (int64_t)I * H, &wu));
      TRY(get_staged(c, q + "mlp.down_proj.weight", false, (int64_t)H * I, &wd));
      TRY(dalloc(c, &w.wo, (size_t)H * H));
      TRY(dalloc(c, &w.wgu, (size_t)2 * I * H));
      TRY(dalloc(c, &w.wdown, (size_t)H * I));
      CK(cudaMemcpy(w.wo, wo, (size_t)H * H * 2, cudaMemcpyDeviceToDevice));
      pack_rows_kernel<<<I, 256>>>((bf16*)wg, H, (float*)g2, nullptr, nullptr, w.wgu, H, 2, 0, nullptr, nullptr);
      pack_rows_kernel<<<I, 256>>>((bf16*)wu, H, (float*)g2, nullptr, nullptr, w.wgu, H, 2, 1, nullptr, nullptr);
      CKL();
      CK(cudaMemcpy(w.wdown, wd, (size_t)H * I * 2, cudaMemcpyDeviceToDevice));
      CK(cudaDeviceSynchronize());
      drop_staged(c, q + "self_attn.o_proj.weight");
      drop_staged(c, q + "mlp.gate_proj.weight");
      drop_staged(c, q + "mlp.up_proj.weight");
      drop_staged(c, q + "mlp.down_proj.weight");
    }
    void *nf, *lm;
    TRY(get_staged(c, "model.norm.weight", true, H, &nf));
    TRY(get_staged(c, "lm_head.weight", false, (int64_t)V * H, &lm));
    TRY(dalloc(c, &c->lm_head, (size_t)V * H));
    pack_rows_kernel<<<V, 256>>>((bf16*)lm, H, (float*)nf, nullptr, nullptr, c->lm_head, H, 0, 0, nullptr, nullptr);
    CKL();
    CK(cudaDeviceSynchronize());
    drop_staged(c, "lm_head.weight");
    TRY(dalloc(c, &c->rope, (size_t)g.max_position_embeddings * 64));
    rope_table_kernel<<<cdiv((long long)g.max_position_embeddings * 64, 256), 256>>>(c->rope, g.max_position_embeddings, g.rope_theta, 128);
    CKL();
  }
  CK(cudaDeviceSynchronize());
  for (auto& kvp : c->staged) cudaFree(kvp.second.dev);
  c->staged.clear();
  c->finalized = true;
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// ViT encode
// ------------------------------------------------------------------------------------------------
static long long* g_attn_dbg = nullptr;
extern "C" int vly_debug_attn_counters(long long* host_out, int n) {
  if (!g_attn_dbg) return -1;
  return cudaMemcpy(host_out, g_attn_dbg, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
static int launch_vit_attention(vly_ctx* c, const bf16* qkv, int F, bf16* out, cudaStream_t st) {
  const vly_config& g = c->cfg;
  const int D = g.vit_hidden, tokens = (g.vit_image / g.vit_patch) * (g.vit_image / g.vit_patch) + 1;
  if (tokens != 257) return fail(VLY_ERR_INVALID, "vit attention kernel is specialised for 257 tokens (got %d)", tokens);
  TRY(ensure_smem_attr(c->cfg.device, vit_attention_kernel, VitAttnCfg::SMEM_BYTES));
  CUtensorMap tq, tkv;
  TRY(make_tmap_2d(c, &tq, qkv, 3 * D, (uint64_t)F * tokens, (uint64_t)3 * D * 2, 64, 128));
  TRY(make_tmap_2d(c, &tkv, qkv, 3 * D, (uint64_t)F * tokens, (uint64_t)3 * D * 2, 64, 136));
  VitAttnParams p;
  p.F = F;
  p.tokens = tokens;
  p.heads = g.vit_heads;
  p.D = D;
  p.ctx = out;
  p.qkv = qkv;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  {
    static long long* dbg = nullptr;
    if (getenv("VLY_ATTN_DBG") && !dbg) { CK(cudaMalloc((void**)&dbg, 148 * 16 * 8)); CK(cudaMemset(dbg, 0, 148 * 16 * 8)); }
    p.dbg = dbg;
    g_attn_dbg = dbg;
  }
  const int items = F * g.vit_heads;
  const int grid = items < c->num_sms ? items : c->num_sms;
  static const bool v1 = getenv("VLY_VIT_ATTN_V1") != nullptr;
  if (!v1) {
    TRY(ensure_smem_attr(c->cfg.device, vit_attention_pp_kernel, VitAttnPPCfg::SMEM_BYTES));
    CUtensorMap tx;
    TRY(make_tmap_2d(c, &tx, qkv, 3 * D, (uint64_t)F * tokens, (uint64_t)3 * D * 2, 64, 1, /*swizzle128=*/false));   // single rows, read linearly
    CK(launch_ex(vit_attention_pp_kernel, dim3(grid), dim3(VitAttnPPCfg::THREADS), VitAttnPPCfg::SMEM_BYTES, st, true, tq, tx, p));
    c->launches++;
    return VLY_OK;
  }
  vit_attention_kernel<<<grid, VitAttnCfg::THREADS, VitAttnCfg::SMEM_BYTES, st>>>(tq, tkv, p);
  c->launches++;
  CKL();
  return VLY_OK;
}

static int vit_layers_needed(const vly_config& g, int select_layer, int* out) {
  const int idx = select_layer >= 0 ? select_layer : g.vit_layers + 1 + select_layer;
  if (idx < 0 || idx > g.vit_layers) return fail(VLY_ERR_INVALID, "select_layer %d out of range for %d layers", select_layer, g.vit_layers);
  *out = idx;
  return VLY_OK;
}

static int vit_encode_impl(vly_ctx* c, const void* pixels, int pixel_dtype, int F, int select_layer, void* out_dev, void* stream,
                           bool gather, long long gather_frame_off, int gather_frame_stride);

extern "C" int vly_vit_encode(vly_ctx* c, const void* pixels, int pixel_dtype, int F, int select_layer, void* out_dev, void* stream) {
  if (!c || !pixels || !out_dev || F <= 0) return fail(VLY_ERR_INVALID, "vly_vit_encode: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  return vit_encode_impl(c, pixels, pixel_dtype, F, select_layer, out_dev, stream, false, 0, 1);
}

static int vit_encode_impl(vly_ctx* c, const void* pixels, int pixel_dtype, int F, int select_layer, void* out_dev, void* stream,
                           bool gather, long long gather_frame_off, int gather_frame_stride) {
  if (!c->finalized || !c->has_vit) return fail(VLY_ERR_STATE, "vly_vit_encode: vision weights not loaded/finalised");
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const vly_config& g = c->cfg;
  const int D = g.vit_hidden, MLP = g.vit_mlp, P = g.vit_patch, IMG = g.vit_image;
  const int NP = (IMG / P) * (IMG / P), tokens = NP + 1;
  int n_layers;
  TRY(vit_layers_needed(g, select_layer, &n_layers));
  const int CH = 256;  // frames per chunk: bounds the workspace (qkv 405 MB, mlp 540 MB) and keeps tiles plentiful
  const int fc_max = F < CH ? F : CH;
  const size_t Mmax = (size_t)fc_max * tokens;
  TRY(ensure(c->w_col, (size_t)fc_max * NP * c->kpad * 2));
  TRY(ensure(c->w_patch, (size_t)fc_max * NP * D * 2));
  TRY(ensure(c->w_qkv, Mmax * 3 * D * 2));
  TRY(ensure(c->w_ctx, Mmax * D * 2));
  TRY(ensure(c->w_h, Mmax * MLP * 2));
  const int nt_max = cdiv(D, 128);
  TRY(ensure(c->w_stats, Mmax * nt_max * sizeof(float2)));
  const size_t px_elem = pixel_dtype == VLY_F32 ? 4 : 2;
  for (int f0 = 0; f0 < F; f0 += CH) {
    const int fc = (F - f0) < CH ? (F - f0) : CH;
    const int M = fc * tokens;
    const int bn_d = pick_bn_m(c, D, M), nt = cdiv(D, bn_d);       // every N = D GEMM of this chunk (they exchange row statistics)
    const char* px = (const char*)pixels + (size_t)f0 * 3 * IMG * IMG * px_elem;
    bf16* x = (bf16*)out_dev + (size_t)f0 * tokens * D;
    bf16* col = (bf16*)c->w_col.p;
    const int blocks = 148 * 8;
    if (pixel_dtype == VLY_F32) im2col_kernel<float><<<blocks, 256, 0, st>>>((const float*)px, col, fc, IMG, P, c->kpad);
    else if (pixel_dtype == VLY_BF16) im2col_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)px, col, fc, IMG, P, c->kpad);
    else if (pixel_dtype == VLY_F16) im2col_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)px, col, fc, IMG, P, c->kpad);
    else return fail(VLY_ERR_INVALID, "vly_vit_encode: unknown pixel dtype %d", pixel_dtype);
    c->launches++;
    CKL();
    {  // patch embedding GEMM (conv2d stride=kernel=14, no bias)
      GemmParams p = {};
      p.M = fc * NP; p.N = D; p.K = c->kpad;
      p.out = c->w_patch.p; p.ldo = D;
      TRY(launch_gemm<EPI_BIAS>(c, pick_bn_m(c, D, fc * NP), col, c->kpad, c->patch_w, c->kpad, p, st));
    }
    float2* stats = (float2*)c->w_stats.p;
    vit_embed_ln_kernel<<<M, 128, 0, st>>>((bf16*)c->w_patch.p, c->cls, c->pos, c->pre_g, c->pre_b, x, stats, nt, tokens, D, g.vit_eps);
    c->launches++;
    CKL();
    for (int l = 0; l < n_layers; ++l) {
      const VitLayerW& w = c->vit[l];
      {  // LN1 -> q,k,v
        GemmParams p = {};
        p.M = M; p.N = 3 * D; p.K = D;
        p.out = c->w_qkv.p; p.ldo = 3 * D;
        p.bias = w.qkv_b; p.colsum = w.qkv_cs;
        p.stats_in = stats; p.stats_in_nt = nt; p.inv_dim = 1.f / D; p.eps = g.vit_eps;
        TRY(launch_gemm<EPI_LN_BIAS>(c, pick_bn_m(c, 3 * D, M), x, D, w.wqkv, D, p, st));
      }
      TRY(launch_vit_attention(c, (bf16*)c->w_qkv.p, fc, (bf16*)c->w_ctx.p, st));
      {  // out_proj + residual
        GemmParams p = {};
        p.M = M; p.N = D; p.K = D;
        p.out = x; p.ldo = D; p.bias = w.bo; p.residual = x; p.ldr = D; p.stats_out = stats;
        TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, bn_d, (bf16*)c->w_ctx.p, D, w.wo, D, p, st));
      }
      {  // LN2 -> fc1 -> quick_gelu
        GemmParams p = {};
        p.M = M; p.N = MLP; p.K = D;
        p.out = c->w_h.p; p.ldo = MLP; p.bias = w.b1; p.colsum = w.c1;
        p.stats_in = stats; p.stats_in_nt = nt; p.inv_dim = 1.f / D; p.eps = g.vit_eps;
        TRY(launch_gemm<EPI_LN_BIAS_GELU>(c, pick_bn_m(c, MLP, M), x, D, w.w1, D, p, st));
      }
      {  // fc2 + residual (+ on the last layer of a sharded encode: push every tile to all ranks' gather buffers)
        GemmParams p = {};
        p.M = M; p.N = D; p.K = MLP;
        p.out = x; p.ldo = D; p.bias = w.b2; p.residual = x; p.ldr = D; p.stats_out = stats;
        if (gather && l == n_layers - 1) {
          p.n_peers = c->g_world;
          for (int q = 0; q < c->g_world; ++q) p.peer_out[q] = c->g_peer_buf[q];
          p.peer_tokens = tokens;
          p.peer_frame_stride = gather_frame_stride;
          p.peer_frame_off = gather_frame_off + (long long)f0 * gather_frame_stride;
        }
        TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, bn_d, (bf16*)c->w_h.p, MLP, w.w2, MLP, p, st));
      }
    }
    if (gather && n_layers == 0) return fail(VLY_ERR_INVALID, "vly_vit_encode_gather needs at least one encoder layer (select_layer != 0)");
  }
  return VLY_OK;
}

// ---- fused all-gather plumbing ----
struct PeerFlags { int* p[8]; };
__global__ void gather_signal_kernel(PeerFlags pf, int world, int rank, int epoch) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();                                   // this GPU's peer stores (previous kernels) are performed
    *reinterpret_cast<volatile int*>(pf.p[threadIdx.x] + rank) = epoch;
  }
}
__global__ void gather_wait_kernel(volatile int* flags, int world, int epoch, int* timeout_flag) {
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (flags[threadIdx.x] < epoch) {
      if (clock64() - t0 > 20000000000LL) {                               // ~10 s: a peer died
        *reinterpret_cast<volatile int*>(timeout_flag) = 1;                // pinned host memory: the host sees it (gather_check_timeout)
        break;
      }
    }
    __threadfence_system();
  }
}

// A peer that never signals (dead / wedged rank) makes gather_wait_kernel give up after ~10 s and raise the context's pinned
// timeout flag.  The gather buffer then holds partially written or previous-epoch rows, so every later gather call -- and
// vly_gather_status, which callers poll after their next synchronisation -- fails loudly instead of decoding stale features.
static int gather_check_timeout(vly_ctx* c, const char* who) {
  if (c->g_timeout && *reinterpret_cast<volatile int*>(c->g_timeout) != 0)
    return fail(VLY_ERR_STATE, "%s: a peer rank never signalled its frame features (fused all-gather timed out); the gather "
                "buffer is stale -- the process group must be torn down", who);
  return VLY_OK;
}

extern "C" int vly_gather_status(vly_ctx* c, int* timed_out) {
  if (!c || !timed_out) return fail(VLY_ERR_INVALID, "vly_gather_status: null argument");
  *timed_out = (c->g_timeout && *reinterpret_cast<volatile int*>(c->g_timeout) != 0) ? 1 : 0;
  return *timed_out ? gather_check_timeout(c, "vly_gather_status") : VLY_OK;
}

extern "C" int vly_gather_create(vly_ctx* c, int64_t rows_total, void** local_buf, void* handle_out) {
  if (!c || rows_total <= 0 || !local_buf || !handle_out) return fail(VLY_ERR_INVALID, "vly_gather_create: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  if (c->g_buf) return fail(VLY_ERR_STATE, "vly_gather_create: a gather buffer already exists for this context");
  const size_t data = (((size_t)rows_total * c->cfg.vit_hidden * 2) + 255) & ~size_t(255);
  void* base;
  CK(cudaMalloc(&base, data + 256));
  CK(cudaMemset(base, 0, data + 256));
  c->g_buf = (bf16*)base;
  c->g_rows = rows_total;
  c->g_flags = (int*)((char*)base + data);
  if (!c->g_timeout) {
    CK(cudaHostAlloc((void**)&c->g_timeout, sizeof(int), cudaHostAllocMapped));
    *c->g_timeout = 0;
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  memcpy(handle_out, &h, 64);
  *local_buf = base;
  return VLY_OK;
}

extern "C" int vly_gather_open_peers(vly_ctx* c, const void* handles, int world, int rank) {
  if (!c || !handles || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(VLY_ERR_INVALID, "vly_gather_open_peers: bad argument (world <= 8)");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->g_buf) return fail(VLY_ERR_STATE, "vly_gather_open_peers: call vly_gather_create first");
  CK(cudaSetDevice(c->cfg.device));
  const size_t data = (((size_t)c->g_rows * c->cfg.vit_hidden * 2) + 255) & ~size_t(255);
  for (int q = 0; q < world; ++q) {
    void* base = nullptr;
    if (q == rank) base = c->g_buf;
    else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + (size_t)q * 64, 64);
      CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->g_peer_buf[q] = (bf16*)base;
    c->g_peer_flags[q] = (int*)((char*)base + data);
  }
  c->g_world = world;
  c->g_rank = rank;
  return VLY_OK;
}

// Tell every rank that this rank has finished reading the current epoch's gather buffer (enqueue after the consumer kernels).
extern "C" int vly_gather_release(vly_ctx* c, void* stream) {
  if (!c) return fail(VLY_ERR_INVALID, "null");
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->g_world == 0) return fail(VLY_ERR_STATE, "vly_gather_release: no gather group");
  TRY(gather_check_timeout(c, "vly_gather_release"));
  CK(cudaSetDevice(c->cfg.device));
  PeerFlags pf;
  for (int q = 0; q < 8; ++q) pf.p[q] = c->g_peer_flags[q] ? c->g_peer_flags[q] + 8 : nullptr;
  gather_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, c->g_world, c->g_rank, c->g_epoch);
  c->launches++;
  CKL();
  return VLY_OK;
}

extern "C" int vly_vit_encode_gather(vly_ctx* c, const void* pixels, int pixel_dtype, int F, int frame_offset, int select_layer, void* stream) {
  return vly_vit_encode_gather_strided(c, pixels, pixel_dtype, F, frame_offset, 1, select_layer, stream);
}

extern "C" int vly_vit_encode_gather_strided(vly_ctx* c, const void* pixels, int pixel_dtype, int F, int frame_offset, int frame_stride,
                                             int select_layer, void* stream) {
  if (!c || F < 0 || frame_offset < 0 || frame_stride < 1 || (F > 0 && !pixels)) return fail(VLY_ERR_INVALID, "vly_vit_encode_gather: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->finalized || !c->has_vit) return fail(VLY_ERR_STATE, "vly_vit_encode_gather: vision weights not loaded/finalised");
  if (c->g_world == 0) return fail(VLY_ERR_STATE, "vly_vit_encode_gather: call vly_gather_create / vly_gather_open_peers first");
  const int tokens = (c->cfg.vit_image / c->cfg.vit_patch) * (c->cfg.vit_image / c->cfg.vit_patch) + 1;
  if (F > 0 && ((long long)frame_offset + (long long)(F - 1) * frame_stride + 1) * tokens > c->g_rows)
    return fail(VLY_ERR_INVALID, "vly_vit_encode_gather: frames %d + i*%d, i < %d exceed the gather buffer", frame_offset, frame_stride, F);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  TRY(gather_check_timeout(c, "vly_vit_encode_gather"));
  int* timeout_flag = c->g_timeout;
  if (c->g_epoch > 0) {   // every rank must have finished READING the previous epoch before anyone overwrites its buffer
    gather_wait_kernel<<<1, 32, 0, st>>>(c->g_flags + 8, c->g_world, c->g_epoch, timeout_flag);
    c->launches++;
    CKL();
  }
  if (F > 0) {
    TRY(ensure(c->w_xlocal, (size_t)F * tokens * c->cfg.vit_hidden * 2));       // local residual stream (scratch)
    TRY(vit_encode_impl(c, pixels, pixel_dtype, F, select_layer, c->w_xlocal.p, stream, true, frame_offset, frame_stride));
  }
  // flag exchange: every rank tells every rank "my rows are in your buffer", then waits for all of them
  const int epoch = ++c->g_epoch;
  PeerFlags pf;
  for (int q = 0; q < 8; ++q) pf.p[q] = c->g_peer_flags[q];
  gather_signal_kernel<<<1, 32, 0, st>>>(pf, c->g_world, c->g_rank, epoch);
  gather_wait_kernel<<<1, 32, 0, st>>>(c->g_flags, c->g_world, epoch, timeout_flag);
  c->launches += 2;
  CKL();
  return VLY_OK;
}

extern "C" int vly_project(vly_ctx* c, const void* feats, int64_t rows, void* out, void* stream) {
  if (!c || !feats || !out || rows <= 0) return fail(VLY_ERR_INVALID, "vly_project: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->finalized || !c->proj_w) return fail(VLY_ERR_STATE, "vly_project: mm_projector not loaded");
  CK(cudaSetDevice(c->cfg.device));
  GemmParams p = {};
  p.M = (int)rows; p.N = c->cfg.hidden_size; p.K = c->cfg.vit_hidden;
  p.out = out; p.ldo = c->cfg.hidden_size; p.bias = c->proj_b;
  return launch_gemm<EPI_BIAS>(c, pick_bn(p.N), (const bf16*)feats, p.K, c->proj_w, p.K, p, (cudaStream_t)stream);
}

// 'max' and 'temporal_transformer' act on the PROJECTED tokens (they do not commute with the projector): project all
// T*257 rows of one video at a time, then pool.  valley_model.py:208-209, :123-133.
static int pool_project_after(vly_ctx* c, const bf16* feats, int n_videos, int T, bf16* vis_rows, cudaStream_t st) {
  const vly_config& g = c->cfg;
  const int D = g.vit_hidden, H = g.hidden_size, tokens = (g.vit_image / g.vit_patch) * (g.vit_image / g.vit_patch) + 1;
  const int NP = tokens - 1, rows_out = NP + T, nhead = 8;
  const vly_ctx::DeltaW& w = c->delta;
  const bool tr = g.patch_pooling_method == VLY_POOL_TEMPORAL_TRANSFORMER;
  if (tr) {
    if (!w.in_w) return fail(VLY_ERR_STATE, "vly_pool_project: transformer_delta_encoder weights not loaded");
    if (T > 64 || T > w.max_pos) return fail(VLY_ERR_INVALID, "vly_pool_project: temporal transformer supports at most %d frames (got %d)", w.max_pos < 64 ? w.max_pos : 64, T);
    if ((H / nhead) % 64) return fail(VLY_ERR_INVALID, "vly_pool_project: hidden_size / 8 must be a multiple of 64");
  }
  TRY(ensure(c->w_pall, (size_t)T * tokens * H * 2));
  if (tr) {
    TRY(ensure(c->w_xp, (size_t)T * NP * H * 2));
    TRY(ensure(c->w_dkv, (size_t)T * NP * 2 * H * 2));
    TRY(ensure(c->w_dq, (size_t)NP * H * 2));
    TRY(ensure(c->w_datt, (size_t)NP * H * 2));
    TRY(ensure(c->w_dx1, (size_t)NP * H * 2));
    TRY(ensure(c->w_df1, (size_t)NP * w.ffn * 2));
    TRY(ensure(c->w_dx2, (size_t)NP * H * 2));
  }
  bf16* P = (bf16*)c->w_pall.p;
  for (int v = 0; v < n_videos; ++v) {
    bf16* vis = vis_rows + (size_t)v * rows_out * H;
    {  // mm_projector over every token of the video (valley_model.py:187-190)
      GemmParams p = {};
      p.M = T * tokens; p.N = H; p.K = D; p.out = P; p.ldo = H; p.bias = c->proj_b;
      TRY(launch_gemm<EPI_BIAS>(c, pick_bn_m(c, H, p.M), feats + (size_t)v * T * tokens * D, D, c->proj_w, D, p, st));
    }
    if (!tr) {
      temporal_max_kernel<<<148 * 2, 256, 0, st>>>(P, vis, T, tokens, H);
      c->launches++;
      CKL();
      continue;
    }
    bf16 *Xp = (bf16*)c->w_xp.p, *KV = (bf16*)c->w_dkv.p, *Q = (bf16*)c->w_dq.p, *att = (bf16*)c->w_datt.p;
    bf16 *X1 = (bf16*)c->w_dx1.p, *F1 = (bf16*)c->w_df1.p, *X2 = (bf16*)c->w_dx2.p;
    const bf16* Xlast = Xp + (size_t)(T - 1) * NP * H;          // rows of the last frame: the only queries that are used (:130)
    delta_add_pos_kernel<<<148 * 2, 256, 0, st>>>(P, w.pos, Xp, T, tokens, H);
    c->launches++;
    GemmParams p = {};
    p.M = T * NP; p.N = 2 * H; p.K = H; p.out = KV; p.ldo = 2 * H; p.bias = w.in_b + H;          // k | v rows of in_proj
    TRY(launch_gemm<EPI_BIAS>(c, pick_bn_m(c, p.N, p.M), Xp, H, w.in_w + (size_t)H * H, H, p, st));
    p = {};
    p.M = NP; p.N = H; p.K = H; p.out = Q; p.ldo = H; p.bias = w.in_b;
    TRY(launch_gemm<EPI_BIAS>(c, pick_bn_m(c, p.N, p.M), Xlast, H, w.in_w, H, p, st));
    delta_attention_kernel<<<NP, nhead * 32, 0, st>>>(Q, KV, att, T, NP, H, nhead);
    c->launches++;
    p = {};
    p.M = NP; p.N = H; p.K = H; p.out = X1; p.ldo = H; p.bias = w.out_b; p.residual = Xlast; p.ldr = H;
    TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, pick_bn_m(c, p.N, p.M), att, H, w.out_w, H, p, st));
    layernorm_rows_kernel<<<NP, 256, 0, st>>>(X1, w.n1_g, w.n1_b, X1, H, 1e-5f);
    c->launches++;
    p = {};
    p.M = NP; p.N = w.ffn; p.K = H; p.out = F1; p.ldo = w.ffn; p.bias = w.l1_b;
    TRY(launch_gemm<EPI_BIAS>(c, pick_bn_m(c, p.N, p.M), X1, H, w.l1_w, H, p, st));
    relu_inplace_kernel<<<148, 256, 0, st>>>(F1, (long long)NP * w.ffn / 8);
    c->launches++;
    p = {};
    p.M = NP; p.N = H; p.K = w.ffn; p.out = X2; p.ldo = H; p.bias = w.l2_b; p.residual = X1; p.ldr = H;
    TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, pick_bn_m(c, p.N, p.M), F1, w.ffn, w.l2_w, w.ffn, p, st));
    layernorm_rows_kernel<<<NP, 256, 0, st>>>(X2, w.n2_g, w.n2_b, X2, H, 1e-5f);
    delta_finish_kernel<<<148 * 2, 256, 0, st>>>(P, X2, vis, T, tokens, H);
    c->launches += 2;
    CKL();
  }
  return VLY_OK;
}

extern "C" int vly_pool_project(vly_ctx* c, const void* feats, int n_videos, int T, void* vis_rows, void* stream) {
  if (!c || !feats || !vis_rows || n_videos <= 0 || T <= 0) return fail(VLY_ERR_INVALID, "vly_pool_project: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->finalized || !c->proj_w) return fail(VLY_ERR_STATE, "vly_pool_project: mm_projector not loaded");
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const vly_config& g = c->cfg;
  const int D = g.vit_hidden, tokens = (g.vit_image / g.vit_patch) * (g.vit_image / g.vit_patch) + 1;
  const int rows = tokens - 1 + T;
  if (g.patch_pooling_method == VLY_POOL_MAX || g.patch_pooling_method == VLY_POOL_TEMPORAL_TRANSFORMER)
    return pool_project_after(c, (const bf16*)feats, n_videos, T, (bf16*)vis_rows, st);
  TRY(ensure(c->w_pool, (size_t)n_videos * rows * D * 2));
  if (g.patch_pooling_method == VLY_POOL_TEMPORAL_IMPORTANCE) {
    // softmax_t(w . flatten(proj(x_t))) weights: scores straight from the ViT features through the folded U (pooling_kernels.cuh)
    if (!c->pool_U) return fail(VLY_ERR_STATE, "vly_pool_project: model.pooling_layer.weight not loaded");
    TRY(ensure(c->w_score, (size_t)n_videos * T * 4));
    importance_score_kernel<<<dim3(T, n_videos), 256, 0, st>>>((const bf16*)feats, c->pool_U, (float*)c->w_score.p, T, tokens, D);
    weighted_pool_kernel<<<148 * 4, 256, 0, st>>>((const bf16*)feats, (const float*)c->w_score.p, (bf16*)c->w_pool.p, n_videos, T, tokens, D);
    c->launches++;
  } else {
    temporal_pool_kernel<<<148 * 4, 256, 0, st>>>((const bf16*)feats, (bf16*)c->w_pool.p, n_videos, T, tokens, D);
  }
  c->launches++;
  CKL();
  GemmParams p = {};
  p.M = n_videos * rows; p.N = g.hidden_size; p.K = D;
  p.out = vis_rows; p.ldo = g.hidden_size; p.bias = c->proj_b;
  return launch_gemm<EPI_BIAS>(c, pick_bn(p.N), (bf16*)c->w_pool.p, D, c->proj_w, D, p, st);
}

extern "C" int vly_embed_splice(vly_ctx* c, const int64_t* ids, const int32_t* src_map, const int32_t* img_idx, const void* vis_rows,
                                int rows_per_img, int B, int S, void* out, void* stream) {
  if (!c || !ids || !out || B <= 0 || S <= 0) return fail(VLY_ERR_INVALID, "vly_embed_splice: bad argument");
  if (src_map && (!img_idx || !vis_rows)) return fail(VLY_ERR_INVALID, "vly_embed_splice: src_map given without img_idx / vis_rows");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->finalized || !c->has_llm) return fail(VLY_ERR_STATE, "vly_embed_splice: LLM weights not loaded");
  CK(cudaSetDevice(c->cfg.device));
  embed_splice_kernel<<<B * S, 128, 0, (cudaStream_t)stream>>>((const long long*)ids, src_map, img_idx, c->embed, (const bf16*)vis_rows,
                                                               rows_per_img, (bf16*)out, nullptr, 0, S, c->cfg.hidden_size, c->cfg.vocab_size);
  c->launches++;
  CKL();
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// KV cache
// ------------------------------------------------------------------------------------------------
// Shared-memory budget of the persistent decode kernel (decode_mega.cuh): activation block + barriers / handoff slots, the
// rest is the weight ring.
static void mega_smem_layout(const vly_config& g, int bmax, size_t* x_bytes, size_t* misc) {
  const bool tc = bmax > 1;
  const int kmax = g.intermediate_size > g.hidden_size ? g.intermediate_size : g.hidden_size;
  const size_t xs_stride_b = tc ? (size_t)(((kmax * 2 + 127) & ~127) + MegaCfg::PAD_TC) : (size_t)kmax * 2;
  *x_bytes = ((size_t)bmax * xs_stride_b + 127) & ~size_t(127);
  const int nv = (tc ? MegaCfg::ROWS_TC : MegaCfg::ROWS) * bmax;
  *misc = (2 * MegaCfg::MAX_STAGES + 2 * MegaCfg::RED_SLOTS + 2) * 8 + (size_t)MegaCfg::RED_SLOTS * 16 * nv * 4 + (3 + 16) * bmax * 4 + 256;
}
static const size_t kMegaSmem = 226 * 1024;

// Ring geometry of one weight phase of the persistent decode kernel: rows per work unit and columns per stage.  Measured with
// tools/ringbw.cu (the producer's access pattern without consumers, profiles/ringbw_r02.log):
//   * what streams fastest is ~110-125 KB of bulk copies outstanding per SM (7.3-7.4 TB/s); 64 KB: 6.2-7.1, 190 KB: 6.7-7.0;
//   * a bulk copy should be >= 4 KB: 8 rows x 2.5 KB stream at 5.9 TB/s where 4 rows x 5 KB reach 7.3;
//   * the consumers need a few hundred cycles to hand a landed stage back, which takes that stage out of flight: a ring of 3
//     large stages loses more to this than one of 5 smaller stages, so the target is a stage of 1/5 of the ring budget
//     (but 16-48 KB), cut so that K divides into EQUAL stages (no short tail stage);
//   * rows: the candidate (max, max / 2) that satisfies the copy-size rule, then the one whose work units balance better over
//     the SMs -- the phase lasts as long as its most loaded CTA: N = 5120 is 5 rounds of 8 rows (40) but 9 rounds of 4 (36).
// VLY_MEGA_STAGE_KB / VLY_MEGA_ROWS override (A/B measurements).
static void pick_phase_geometry(int bmax, size_t ring_budget, int N, int K, int num_sms, int* rows_out, int* kc_out) {
  const bool tc = bmax > 1;
  const int max_rows = tc ? MegaCfg::ROWS_TC : MegaCfg::ROWS;
  const int pad = tc ? MegaCfg::PAD_TC : 0, gran = tc ? 64 : 8;
  static const int env_kb = getenv("VLY_MEGA_STAGE_KB") ? atoi(getenv("VLY_MEGA_STAGE_KB")) : 0;
  static const int env_rows = getenv("VLY_MEGA_ROWS") ? atoi(getenv("VLY_MEGA_ROWS")) : 0;
  // measured (same-box A/B of the real step, profiles/decode_ab_r02.txt): CUDA-core path -- 4 slots of 32-41 KB beat 3 (more
  // bytes in flight) and 6 (the grid barriers queue behind the extra prefetch traffic); tensor-core path -- 8 rows x ~2048 columns
  size_t target = env_kb > 0 ? (size_t)env_kb * 1024 : (tc ? 33 * 1024 + 512 : 41 * 1024);
  if (target > ring_budget / 2) target = ring_budget / 2;
  if (target < 8 * 1024) target = 8 * 1024;
  auto kc_for = [&](int rows) {
    const int cap = (int)((target / rows - pad) / 2);
    const int ns = cdiv(K, cap > gran ? cap : gran);
    int kc = cdiv(cdiv(K, ns), gran) * gran;
    return kc > K ? cdiv(K, gran) * gran : kc;
  };
  int rows = max_rows;
  if (env_rows > 0) rows = env_rows > max_rows ? max_rows : env_rows;
  else {
    if (kc_for(rows) * 2 < 3072 && K * 2 >= 3072) rows = max_rows / 2;            // keep every bulk copy >= 3 KB
    if (rows == max_rows && !tc) {        // (tensor-core path: half-filled HMMAs cost more than the imbalance they would remove)
      const long long load_full = (long long)cdiv(cdiv(N, max_rows), num_sms) * max_rows;
      const long long load_half = (long long)cdiv(cdiv(N, max_rows / 2), num_sms) * (max_rows / 2);
      if (load_half * 100 < load_full * 97) rows = max_rows / 2;                  // at least 3 % shorter critical path
    }
  }
  *rows_out = rows;
  *kc_out = kc_for(rows);
}

extern "C" int vly_kv_create(vly_ctx* c, int batch, int max_seq, vly_kv** out) {
  if (!c || !out || batch <= 0 || max_seq <= 0) return fail(VLY_ERR_INVALID, "vly_kv_create: bad argument");
  if (!c->finalized || !c->has_llm) return fail(VLY_ERR_STATE, "vly_kv_create: LLM weights not finalised");
  if (max_seq > c->cfg.max_position_embeddings) return fail(VLY_ERR_INVALID, "vly_kv_create: max_seq %d > max_position_embeddings %d", max_seq, c->cfg.max_position_embeddings);
  CK(cudaSetDevice(c->cfg.device));
  const vly_config& g = c->cfg;
  vly_kv* kv = new vly_kv();
  kv->ctx = c;
  kv->B = batch;
  kv->Smax = (max_seq + 127) / 128 * 128;   // whole 128-key TMA tiles
  const int H = g.hidden_size, nH = g.num_attention_heads, I = g.intermediate_size, V = g.vocab_size, L = g.num_hidden_layers;
  const size_t cache_elems = (size_t)L * kv->layer_stride();
  CK(cudaMalloc((void**)&kv->cache, cache_elems * 2));
  CK(cudaMemset(kv->cache, 0, cache_elems * 2));   // padded keys must be finite: P(=0) * V(pad) must stay 0
  // split-KV factor: enough CTAs to cover the GPU about twice
  if (decode_mode() == 0) {
    int ns = (2 * c->num_sms + batch * nH - 1) / (batch * nH);
    kv->nsplit = ns < 1 ? 1 : (ns > 16 ? 16 : ns);
  } else {
    kv->nsplit = kv->Smax / MegaCfg::ATTN_KEYS_MIN; // capacity of the split dimension: the persistent kernel uses 16/32-key items, the
                                                // per-op kernel fixed 64-key splits (it only touches the first Smax / 64 slots)
  }
  kv->gemv_grid = 2 * c->num_sms;
  CK(cudaMalloc((void**)&kv->d_len, 8));
  CK(cudaHostAlloc((void**)&kv->h_len, sizeof(int), cudaHostAllocDefault));
  CK(cudaEventCreateWithFlags(&kv->len_event, cudaEventDisableTiming));
  kv->d_step = kv->d_len + 1;
  CK(cudaMemset(kv->d_len, 0, 8));
  // (8 rows tall, rows >= batch stay zero: the tcgen05 decode consumer stages these buffers as 8-row tensor-TMA boxes)
  const size_t act_rows = batch < 8 ? 8 : batch;
  CK(cudaMalloc((void**)&kv->x, act_rows * H * 2));
  CK(cudaMalloc((void**)&kv->q, act_rows * H * 2));
  CK(cudaMalloc((void**)&kv->attn, act_rows * H * 2));
  CK(cudaMalloc((void**)&kv->hb, act_rows * I * 2));
  CK(cudaMemset(kv->x, 0, act_rows * H * 2));
  CK(cudaMemset(kv->attn, 0, act_rows * H * 2));
  CK(cudaMemset(kv->hb, 0, act_rows * I * 2));
  CK(cudaMalloc((void**)&kv->part_o, (size_t)batch * nH * kv->nsplit * 128 * 4));
  CK(cudaMalloc((void**)&kv->part_ml, (size_t)batch * nH * kv->nsplit * sizeof(float2)));
  CK(cudaMalloc((void**)&kv->counters, ((size_t)batch * nH + 4) * 4));
  CK(cudaMemset(kv->counters, 0, ((size_t)batch * nH + 4) * 4));
  kv->grid_counter = kv->counters + (size_t)batch * nH + 1;      // [+2] = launch epoch of the grid barrier
  CK(cudaMalloc((void**)&kv->part_val, (size_t)batch * kv->gemv_grid * 4));
  CK(cudaMalloc((void**)&kv->part_idx, (size_t)batch * kv->gemv_grid * 4));
  CK(cudaMalloc((void**)&kv->logits, (size_t)batch * V * 4));
  CK(cudaMalloc((void**)&kv->cur_tokens, (size_t)batch * 8));
  CK(cudaMalloc((void**)&kv->gen_tokens, (size_t)batch * kv->Smax * 8));
  CK(cudaMalloc((void**)&kv->dbg, 1024 * 8 * 8));
  CK(cudaMalloc((void**)&kv->d_sample, sizeof(SampleState)));
  {
    SampleState s0 = {};
    s0.inv_temp = 1.f; s0.eos = -1; s0.pad = 0; s0.stop2 = -1;
    CK(cudaMemcpy(kv->d_sample, &s0, sizeof(s0), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc((void**)&kv->key_bits, (size_t)batch * kv->mask_words() * 4));
  CK(cudaMemset(kv->key_bits, 0xff, (size_t)batch * kv->mask_words() * 4));
  // ---- tcgen05 consumer for B = 2..4 (decode_umma.cuh): needs H a multiple of 512, I of 64, and the activation block + ring in 227 KB ----
  {
    static const int env_umma = getenv("VLY_DECODE_UMMA") ? atoi(getenv("VLY_DECODE_UMMA")) : 1;
    static const int env_xc = getenv("VLY_UMMA_XC") ? atoi(getenv("VLY_UMMA_XC")) : 5120;      // (tests shrink it to force sub-phases)
    const int bmax = batch <= 1 ? 1 : (batch <= 2 ? 2 : 4);
    const int xcap = env_xc < H ? H : env_xc;
    // (I not a multiple of 512 -- Llama-2-7B's 11008 -- works through zero-filled out-of-bounds panels but measured slower than the
    //  mma.sync consumer there: 3.82 vs 3.58 ms/step at B = 4, three activation re-stagings for down_proj; VLY_DECODE_UMMA=2 forces it)
    kv->umma = env_umma && bmax > 1 && batch <= 4 && decode_mode() == 2 && (H % 512 == 0) && (I % 512 == 0 || (env_umma >= 2 && I % 64 == 0)) &&
               H <= 5120 && xcap % 512 == 0 &&
               cdiv(cdiv(H, UmmaCfg::ROWS), c->num_sms) <= UmmaCfg::ACC_SLOTS;
    if (kv->umma) {
      // stage width: the largest multiple of 512 columns <= 2560 that divides K
      auto kc_of = [](int K) {
        for (int kc = 2560; kc >= 512; kc -= 512)
          if (K % kc == 0) return kc;
        return 512;
      };
      std::vector<PhaseDesc> ph;
      std::vector<CUtensorMap> maps;
      int err = VLY_OK;
      auto add_map = [&](const bf16* W, int N, int K, int kc) -> int {       // -> index
        CUtensorMap m;
        cuuint64_t dims[3] = {64, (cuuint64_t)N, (cuuint64_t)(K / 64)};
        cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
        cuuint32_t box[3] = {64, (cuuint32_t)UmmaCfg::ROWS, (cuuint32_t)(kc / 64)};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = c->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(W), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) err = fail(VLY_ERR_CUDA, "cuTensorMapEncodeTiled(decode weights, N=%d K=%d kc=%d) failed: %d", N, K, kc, (int)r);
        maps.push_back(m);
        return (int)maps.size() - 1;
      };
      static const int env_oob = getenv("VLY_UMMA_XOOB") ? atoi(getenv("VLY_UMMA_XOOB")) : 1;
      auto add_xmap = [&](const bf16* X, int K, int x_cols) -> int {
        CUtensorMap m;
        // rows >= batch of the 8-row box lie outside the tensor: the TMA unit fills them with zeros without reading memory
        cuuint64_t dims[3] = {64, (cuuint64_t)(env_oob ? batch : 8), (cuuint64_t)(K / 64)};
        cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
        cuuint32_t box[3] = {64, 8, (cuuint32_t)(x_cols / 64)};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = c->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(X), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) err = fail(VLY_ERR_CUDA, "cuTensorMapEncodeTiled(decode activations, K=%d box %d) failed: %d", K, x_cols, (int)r);
        maps.push_back(m);
        return (int)maps.size() - 1;
      };
      std::vector<int> map_of;       // phase -> map index (-1: none)
      std::vector<int> xmap_of;      // phase -> activation map index (-1: none)
      int x_cols_max = 0, stage_max = 0, n_sync = 1;
      // one weight matrix = one or more (sub-)phases: pieces of <= xcap columns (each staged once), a piece = full stages of kc
      // columns plus, if kc does not divide it, ONE shorter tail stage per unit that re-uses the staged block.  K is walked in whole
      // 512-column MMA groups: when 512 does not divide it (Llama-2-7B's down_proj, K = 11008) the last group's missing 64-column
      // panels lie outside BOTH tensor maps (dims use the real K) and the TMA unit zero-fills them without reading memory
      auto add_matrix = [&](PhaseDesc d, const bf16* W) {
        const int K_real = d.K, K = cdiv(d.K, 512) * 512;
        d.ldx = K_real; d.rows = UmmaCfg::ROWS;
        int k0 = 0;
        while (k0 < K) {
          const int piece = (K - k0) < xcap ? (K - k0) : xcap;
          int kc = kc_of(piece);
          if (kc < 1536 && kc != piece) kc = piece < 2560 ? (piece / 512) * 512 : 2560;      // only small divisors: full stages + one tail stage
          const int main_cols = (piece / kc) * kc, tail_cols = piece - main_cols;
          const bool last_piece = (k0 + piece == K);
          PhaseDesc q = d;
          q.k_off = k0; q.K = main_cols; q.kc = kc; q.x_cols = piece; q.x_panel0 = 0;
          q.flags = (k0 == 0 ? PHF_FIRST : 0);
          if (tail_cols == 0) q.flags |= last_piece ? PHF_LAST : PHF_LOCAL_SYNC;
          else q.flags |= PHF_NO_SYNC;
          ph.push_back(q);
          map_of.push_back(add_map(W, d.N, K_real, kc));
          xmap_of.push_back(add_xmap(d.x_in, K_real, piece));
          if (tail_cols > 0) {
            PhaseDesc t = d;
            t.k_off = k0 + main_cols; t.K = tail_cols; t.kc = tail_cols; t.x_cols = 0; t.x_panel0 = main_cols / 64;
            t.flags = last_piece ? PHF_LAST : PHF_LOCAL_SYNC;
            ph.push_back(t);
            map_of.push_back(add_map(W, d.N, K_real, tail_cols));
            xmap_of.push_back(-1);
          }
          if (piece > x_cols_max) x_cols_max = piece;
          if (UmmaCfg::ROWS * kc * 2 > stage_max) stage_max = UmmaCfg::ROWS * kc * 2;
          if (last_piece) ++n_sync;
          k0 += piece;
        }
      };
      for (int l = 0; l < L; ++l) {
        const LlamaLayerW& w = c->layers[l];
        PhaseDesc d = {};
        d.layer = l; d.kcache = kv->k_layer(l); d.vcache = kv->v_layer(l);
        d.type = PH_QKV; d.N = 3 * H; d.K = H; d.W = w.wqkv; d.x_in = kv->x; d.out = kv->q; add_matrix(d, w.wqkv);
        d.type = PH_ATTN; d.N = 0; d.K = 0; d.W = nullptr; d.x_in = nullptr; d.out = kv->attn; ph.push_back(d); map_of.push_back(-1); xmap_of.push_back(-1); ++n_sync;
        d.type = PH_OPROJ; d.N = H; d.K = H; d.W = w.wo; d.x_in = kv->attn; d.out = kv->x; add_matrix(d, w.wo);
        d.type = PH_GATEUP; d.N = 2 * I; d.K = H; d.W = w.wgu; d.x_in = kv->x; d.out = kv->hb; add_matrix(d, w.wgu);
        d.type = PH_DOWN; d.N = H; d.K = I; d.W = w.wdown; d.x_in = kv->hb; d.out = kv->x; add_matrix(d, w.wdown);
      }
      {
        PhaseDesc d = {};
        d.type = PH_LOGITS; d.N = V; d.K = H; d.W = c->lm_head; d.x_in = kv->x; d.out = nullptr; add_matrix(d, c->lm_head);
      }
      if (err != VLY_OK) return err;
      static const int env_if = getenv("VLY_MEGA_INFLIGHT_KB") ? atoi(getenv("VLY_MEGA_INFLIGHT_KB")) : 100;
      for (PhaseDesc& q : ph) {
        if (q.type == PH_ATTN) continue;
        const int sb = UmmaCfg::ROWS * q.kc * 2;
        q.inflight = (env_if * 1024 + sb / 2) / sb;
        if (q.inflight < 2) q.inflight = 2;
      }
      CK(cudaMalloc(&kv->d_tmaps, maps.size() * sizeof(CUtensorMap)));
      CK(cudaMemcpy(kv->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      for (size_t i = 0; i < ph.size(); ++i)
      {
        ph[i].tmap = map_of[i] >= 0 ? (const void*)((const CUtensorMap*)kv->d_tmaps + map_of[i]) : nullptr;
        ph[i].xmap = xmap_of[i] >= 0 ? (const void*)((const CUtensorMap*)kv->d_tmaps + xmap_of[i]) : nullptr;
      }
      kv->umma_x_cols = x_cols_max;
      kv->stage_bytes = stage_max;
      kv->n_grid_syncs = n_sync;
      kv->n_phases = (int)ph.size();
      if (getenv("VLY_MEGA_DBG")) {
        fprintf(stderr, "[vly] decode (tcgen05 consumer): B=%d activation block %d columns (%d KB), stage %d B, %d phases (%d grid barriers);", batch,
                x_cols_max, x_cols_max / 64, stage_max, kv->n_phases, n_sync);
        for (int i = 0; i < 9 && i < (int)ph.size(); ++i)
          if (ph[i].type != PH_ATTN)
            fprintf(stderr, " [type%d N=%d k=%d+%d kc=%d stage_x=%d flags=%d]", ph[i].type, ph[i].N, ph[i].k_off, ph[i].K, ph[i].kc, ph[i].x_cols, ph[i].flags);
        fprintf(stderr, "\n");
      }
      CK(cudaMalloc((void**)&kv->d_phases, ph.size() * sizeof(PhaseDesc)));
      CK(cudaMemcpy(kv->d_phases, ph.data(), ph.size() * sizeof(PhaseDesc), cudaMemcpyHostToDevice));
    }
  }
  if (!kv->umma) {  // phase table of the persistent decode-step kernel: execution order of one step
    std::vector<PhaseDesc> ph;
    const int bmax = batch <= 1 ? 1 : (batch <= 2 ? 2 : 4);
    const int pad = bmax > 1 ? MegaCfg::PAD_TC : 0;
    size_t x_bytes, misc;
    mega_smem_layout(g, bmax, &x_bytes, &misc);
    const size_t ring_budget = kMegaSmem > x_bytes + misc ? kMegaSmem - x_bytes - misc : 0;
    kv->stage_bytes = 0;
    auto add = [&](PhaseDesc d) {
      d.rows = d.kc = 0;
      if (d.type != PH_ATTN) {
        pick_phase_geometry(bmax, ring_budget, d.N, d.K, c->num_sms, &d.rows, &d.kc);
        const int sb = d.rows * (d.kc * 2 + pad);
        if (sb > kv->stage_bytes) kv->stage_bytes = sb;
      }
      ph.push_back(d);
    };
    for (int l = 0; l < L; ++l) {
      const LlamaLayerW& w = c->layers[l];
      PhaseDesc d = {};
      d.layer = l; d.kcache = kv->k_layer(l); d.vcache = kv->v_layer(l);
      d.type = PH_QKV; d.N = 3 * H; d.K = H; d.W = w.wqkv; d.x_in = kv->x; d.out = kv->q; add(d);
      d.type = PH_ATTN; d.N = 0; d.K = 0; d.W = nullptr; d.x_in = nullptr; d.out = kv->attn; add(d);
      d.type = PH_OPROJ; d.N = H; d.K = H; d.W = w.wo; d.x_in = kv->attn; d.out = kv->x; add(d);
      d.type = PH_GATEUP; d.N = 2 * I; d.K = H; d.W = w.wgu; d.x_in = kv->x; d.out = kv->hb; add(d);
      d.type = PH_DOWN; d.N = H; d.K = I; d.W = w.wdown; d.x_in = kv->hb; d.out = kv->x; add(d);
    }
    PhaseDesc d = {};
    d.type = PH_LOGITS; d.N = V; d.K = H; d.W = c->lm_head; d.x_in = kv->x; d.out = nullptr; add(d);
    kv->stage_bytes = (kv->stage_bytes + 127) & ~127;
    // Stages of this phase's size kept in flight: ~100 KB of bulk copies outstanding per SM saturate HBM; every byte beyond
    // that only lengthens the queues the latency-critical traffic (grid barrier, activation staging, attention) waits in --
    // measured: 128 KB in flight made every barrier ~1 us slower at an unchanged streaming rate.  The ring may hold more slots
    // than are in flight: they absorb the consumers' hand-back latency.  VLY_MEGA_INFLIGHT_KB overrides.
    static const int env_if = getenv("VLY_MEGA_INFLIGHT_KB") ? atoi(getenv("VLY_MEGA_INFLIGHT_KB")) : 100;
    for (PhaseDesc& q : ph) {
      if (q.type == PH_ATTN) continue;
      const int sb = q.rows * (q.kc * 2 + pad);
      q.inflight = (env_if * 1024 + sb / 2) / sb;
      if (q.inflight < 2) q.inflight = 2;
    }
    if (getenv("VLY_MEGA_DBG")) {
      const int n_fit = (int)(ring_budget / kv->stage_bytes);
      fprintf(stderr, "[vly] decode ring: B=%d x=%zu B, ring budget %zu B, slot %d B, %d slots fit;", batch, x_bytes, ring_budget, kv->stage_bytes, n_fit);
      for (int i = 0; i < 5 && i < (int)ph.size(); ++i)
        if (ph[i].type != PH_ATTN) fprintf(stderr, " type%d N=%d K=%d rows=%d kc=%d inflight=%d;", ph[i].type, ph[i].N, ph[i].K, ph[i].rows, ph[i].kc, ph[i].inflight);
      fprintf(stderr, " logits rows=%d kc=%d\n", ph.back().rows, ph.back().kc);
    }
    kv->n_phases = (int)ph.size();
    kv->n_grid_syncs = kv->n_phases + 1;
    CK(cudaMalloc((void**)&kv->d_phases, ph.size() * sizeof(PhaseDesc)));
    CK(cudaMemcpy(kv->d_phases, ph.data(), ph.size() * sizeof(PhaseDesc), cudaMemcpyHostToDevice));
  }
  *out = kv;
  return VLY_OK;
}

extern "C" int vly_kv_decode_kernel(vly_kv* kv, char* name, int cap) {
  if (!kv || !name || cap <= 0) return fail(VLY_ERR_INVALID, "vly_kv_decode_kernel: bad argument");
  const int bmax = kv->B <= 1 ? 1 : (kv->B <= 2 ? 2 : 4);
  if (decode_mode() == 2 && kv->B <= 4) snprintf(name, (size_t)cap, "%s<%d>", kv->umma ? "decode_step_umma_kernel" : "decode_step_kernel", bmax);
  else snprintf(name, (size_t)cap, "%s", decode_mode() == 0 ? "per-op decode kernels (generation 1)" : "per-op TMA-ring decode kernels");
  return VLY_OK;
}

extern "C" void vly_kv_destroy(vly_kv* kv) {
  if (!kv) return;
  cudaSetDevice(kv->ctx->cfg.device);
  if (kv->graph) cudaGraphExecDestroy(kv->graph);
  if (kv->graph_n) cudaGraphExecDestroy(kv->graph_n);
  if (kv->h_len) cudaFreeHost(kv->h_len);
  if (kv->len_event) cudaEventDestroy(kv->len_event);
  void* ps[] = {kv->d_tmaps, kv->d_sample, kv->key_bits, kv->dbg, kv->d_phases, kv->cache, kv->d_len, kv->x, kv->q, kv->attn, kv->hb, kv->part_o, kv->part_ml, kv->counters, kv->part_val, kv->part_idx, kv->logits, kv->cur_tokens, kv->gen_tokens};
  for (void* p : ps)
    if (p) cudaFree(p);
  delete kv;
}

extern "C" int vly_kv_seq_len(vly_kv* kv, int* out) {
  if (!kv || !out) return fail(VLY_ERR_INVALID, "null");
  TRY(sync_len(kv));
  *out = kv->host_len;
  return VLY_OK;
}

extern "C" int vly_kv_reset(vly_kv* kv, void* stream) {
  if (!kv) return fail(VLY_ERR_INVALID, "null");
  CK(cudaSetDevice(kv->ctx->cfg.device));
  kv->host_len = 0;
  kv->len_dirty = false;
  CK(cudaMemsetAsync(kv->d_len, 0, 8, (cudaStream_t)stream));
  if (kv->masked) {
    CK(cudaMemsetAsync(kv->key_bits, 0xff, (size_t)kv->B * kv->mask_words() * 4, (cudaStream_t)stream));
    kv->masked = false;
  }
  return VLY_OK;
}

// one thread per 32-key word: bit i = mask[b, 32w+i] != 0; positions >= len stay attendable
__global__ void pack_key_mask_kernel(const uint8_t* __restrict__ mask, int len, int words, uint32_t* __restrict__ bits) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (w >= words) return;
  uint32_t v = 0;
  for (int i = 0; i < 32; ++i) {
    const int k = w * 32 + i;
    v |= uint32_t(k >= len || mask[(size_t)b * len + k] != 0) << i;
  }
  bits[(size_t)b * words + w] = v;
}

extern "C" int vly_kv_set_key_mask(vly_kv* kv, const uint8_t* mask_dev, int len, void* stream) {
  if (!kv || len < 0 || len > kv->Smax || (len > 0 && !mask_dev)) return fail(VLY_ERR_INVALID, "vly_kv_set_key_mask: bad argument (len %d, capacity %d)", len, kv ? kv->Smax : 0);
  std::lock_guard<std::mutex> lk(kv->ctx->mu);
  CK(cudaSetDevice(kv->ctx->cfg.device));
  const int words = kv->mask_words();
  if (len == 0) {
    CK(cudaMemsetAsync(kv->key_bits, 0xff, (size_t)kv->B * words * 4, (cudaStream_t)stream));
    kv->masked = false;
    return VLY_OK;
  }
  pack_key_mask_kernel<<<dim3(cdiv(words, 128), kv->B), 128, 0, (cudaStream_t)stream>>>(mask_dev, len, words, kv->key_bits);
  kv->ctx->launches++;
  CKL();
  kv->masked = true;
  return VLY_OK;
}

extern "C" int vly_kv_export(vly_ctx* c, vly_kv* kv, int layer, int which, void* out, void* stream) {
  if (!c || !kv || !out || layer < 0 || layer >= c->cfg.num_hidden_layers || (which != 0 && which != 1)) return fail(VLY_ERR_INVALID, "vly_kv_export: bad argument");
  TRY(sync_len(kv));
  if (kv->host_len == 0) return VLY_OK;
  CK(cudaSetDevice(c->cfg.device));
  dim3 grid(kv->host_len, kv->B * c->cfg.num_attention_heads);
  kv_export_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(which ? kv->v_layer(layer) : kv->k_layer(layer), (bf16*)out, kv->Smax, kv->host_len, which == 0);
  c->launches++;
  CKL();
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// decode-side launchers
// ------------------------------------------------------------------------------------------------
static size_t gemv_smem(int bmax, int K) { return (size_t)bmax * K * 2 + (size_t)(2 * 8 * 8 * bmax + 2 * 8 * bmax + 3 * bmax) * 4 + 64; }

template <int MODE>
static int launch_gemv(vly_ctx* c, GemvParams p, int grid, cudaStream_t st) {
  const int bmax = p.B <= 1 ? 1 : (p.B <= 2 ? 2 : 4);
  if (p.B > 4) return fail(VLY_ERR_INVALID, "gemv: batch %d > 4 per call (callers split the batch)", p.B);
  if (p.ldx == 0) p.ldx = p.K;
  const size_t smem = gemv_smem(bmax, p.K);
  const int units = (p.N + 7) / 8;
  if (grid > units) grid = units;
#define VLY_GEMV_CASE(BM)                                                                                              \
  {                                                                                                                    \
    TRY(ensure_smem_attr(c->cfg.device, gemv_kernel<BM, MODE>, smem));                                                 \
    gemv_kernel<BM, MODE><<<grid, 256, smem, st>>>(p);                                                                 \
  }
  if (bmax == 1) VLY_GEMV_CASE(1)
  else if (bmax == 2) VLY_GEMV_CASE(2)
  else VLY_GEMV_CASE(4)
#undef VLY_GEMV_CASE
  c->launches++;
  CKL();
  return VLY_OK;
}

// ---- generation-2 decode launchers (TMA ring + programmatic dependent launch) ----
template <int MODE>
static int launch_gemv_ring(vly_ctx* c, GemvParams p, bool pdl, cudaStream_t st) {
  const int bmax = p.B <= 1 ? 1 : (p.B <= 2 ? 2 : 4);
  if (p.B > 4) return fail(VLY_ERR_INVALID, "gemv: batch %d > 4 per call (callers split the batch)", p.B);
  if (p.ldx == 0) p.ldx = p.K;
  const size_t x_bytes = (((size_t)bmax * p.K * 2) + 15) & ~size_t(15);
  const size_t misc = 4096 + 128;
  int n_stages = (int)((113 * 1024 - (long long)x_bytes - (long long)misc) / RingCfg::STAGE_BYTES);   // two kernels co-resident per SM
  if (n_stages >= 3) n_stages = n_stages > 6 ? 6 : n_stages;
  else {
    n_stages = (int)((225 * 1024 - (long long)x_bytes - (long long)misc) / RingCfg::STAGE_BYTES);
    if (n_stages > RingCfg::MAX_STAGES) n_stages = RingCfg::MAX_STAGES;
    if (n_stages < 1) return fail(VLY_ERR_INVALID, "gemv: activation rows (B=%d, K=%d) do not fit in shared memory", p.B, p.K);
  }
  const size_t smem = (size_t)n_stages * RingCfg::STAGE_BYTES + x_bytes + misc;
  const int groups = (p.N + RingCfg::ROWS - 1) / RingCfg::ROWS;
  const int grid = groups < c->num_sms ? groups : c->num_sms;
  cudaError_t e;
#define VLY_RING_CASE(BM)                                                                                                   \
  {                                                                                                                         \
    /* maximum shared-memory carve-out so the NEXT kernel's CTA can co-reside (PDL overlap) */                            \
    TRY(ensure_smem_attr(c->cfg.device, gemv_ring_kernel<BM, MODE>, smem, true));                                           \
    e = launch_ex(gemv_ring_kernel<BM, MODE>, dim3(grid), dim3(RingCfg::THREADS), smem, st, pdl, p, n_stages);              \
  }
  if (bmax == 1) VLY_RING_CASE(1)
  else if (bmax == 2) VLY_RING_CASE(2)
  else VLY_RING_CASE(4)
#undef VLY_RING_CASE
  c->launches++;
  CK(e);
  return VLY_OK;
}

// Enqueue one decode step for batch rows [b0, b0+nb) of kv (nb <= 4).  Reads kv->cur_tokens, writes kv->cur_tokens.
static int enqueue_decode_step(vly_ctx* c, vly_kv* kv, int b0, int nb, bool bump, cudaStream_t st) {
  const vly_config& g = c->cfg;
  const int H = g.hidden_size, nH = g.num_attention_heads, I = g.intermediate_size, V = g.vocab_size;
  bf16 *x = kv->x + (size_t)b0 * H, *q = kv->q + (size_t)b0 * H, *attn = kv->attn + (size_t)b0 * H, *hb = kv->hb + (size_t)b0 * I;
  decode_embed_kernel<<<nb, 256, 0, st>>>(kv->cur_tokens + b0, c->embed, x, H, V);
  c->launches++;
  CKL();
  const size_t attn_smem = ((size_t)kv->Smax / kv->nsplit + 8) * 4;
  if (attn_smem > 40000) TRY(ensure_smem_attr(c->cfg.device, decode_attention_kernel, attn_smem));
  for (int l = 0; l < g.num_hidden_layers; ++l) {
    const LlamaLayerW& w = c->layers[l];
    bf16* kc = kv->k_layer(l) + (size_t)b0 * nH * kv->Smax * 128;
    bf16* vc = kv->v_layer(l) + (size_t)b0 * nH * kv->Smax * 128;
    {
      GemvParams p = {};
      p.N = 3 * H; p.K = H; p.B = nb; p.W = w.wqkv; p.x = x; p.eps = g.rms_norm_eps;
      p.out = q; p.rope = c->rope; p.seq_len = kv->d_len; p.H = H; p.nH = nH; p.Smax = kv->Smax; p.kcache = kc; p.vcache = vc;
      if (use_decode_v1()) TRY(launch_gemv<GEMV_QKV_ROPE>(c, p, kv->gemv_grid, st));
      else TRY(launch_gemv_ring<GEMV_QKV_ROPE>(c, p, true, st));
    }
    {
      DecAttnParams p = {};
      p.B = nb; p.nH = nH; p.H = H; p.Smax = kv->Smax; p.seq_len = kv->d_len;
      p.nsplit = use_decode_v1() ? kv->nsplit : kv->Smax / kDecSplitKeys;
      p.q = q; p.kcache = kc; p.vcache = vc;
      p.part_o = kv->part_o + (size_t)b0 * nH * kv->nsplit * 128;
      p.part_ml = kv->part_ml + (size_t)b0 * nH * kv->nsplit;
      p.counters = kv->counters + (size_t)b0 * nH;
      p.out = attn;
      p.scale_log2e = 0.08838834764831845f * 1.4426950408889634f;
      p.key_bits = kv->key_bits + (size_t)b0 * kv->mask_words(); p.mask_words = kv->mask_words();
      if (use_decode_v1()) {
        dim3 grid(nb * nH, kv->nsplit);
        decode_attention_kernel<<<grid, 128, attn_smem, st>>>(p);
        CKL();
      } else {
        TRY(ensure_smem_attr(c->cfg.device, decode_attention_v2_kernel, 0, true));
        dim3 grid(nb * nH, p.nsplit);
        CK(launch_ex(decode_attention_v2_kernel, grid, dim3(128), 0, st, true, p));
      }
      c->launches++;
    }
    {
      GemvParams p = {};
      p.N = H; p.K = H; p.B = nb; p.W = w.wo; p.x = attn; p.out = x; p.res = x;
      if (use_decode_v1()) TRY(launch_gemv<GEMV_RESIDUAL>(c, p, kv->gemv_grid, st));
      else TRY(launch_gemv_ring<GEMV_RESIDUAL>(c, p, true, st));
    }
    {
      GemvParams p = {};
      p.N = 2 * I; p.K = H; p.B = nb; p.W = w.wgu; p.x = x; p.eps = g.rms_norm_eps; p.out = hb;
      if (use_decode_v1()) TRY(launch_gemv<GEMV_SWIGLU>(c, p, kv->gemv_grid, st));
      else TRY(launch_gemv_ring<GEMV_SWIGLU>(c, p, true, st));
    }
    {
      GemvParams p = {};
      p.N = H; p.K = I; p.B = nb; p.W = w.wdown; p.x = hb; p.out = x; p.res = x;
      if (use_decode_v1()) TRY(launch_gemv<GEMV_RESIDUAL>(c, p, kv->gemv_grid, st));
      else TRY(launch_gemv_ring<GEMV_RESIDUAL>(c, p, true, st));
    }
  }
  {
    GemvParams p = {};
    p.N = V; p.K = H; p.B = nb; p.W = c->lm_head; p.x = x; p.eps = g.rms_norm_eps;
    p.logits = kv->logits + (size_t)b0 * V;
    p.part_val = kv->part_val + (size_t)b0 * kv->gemv_grid;
    p.part_idx = kv->part_idx + (size_t)b0 * kv->gemv_grid;
    p.counter = kv->counters + (size_t)kv->B * nH;
    p.next_tokens = kv->cur_tokens + b0;
    p.out_tokens = kv->gen_tokens + (size_t)b0 * kv->Smax;
    p.out_stride = kv->Smax;
    p.step = kv->d_step;
    p.seq_len_rw = kv->d_len;
    p.bump = bump ? 1 : 0;   // only the last batch group of a step advances the step / length counters
    if (use_decode_v1()) TRY(launch_gemv<GEMV_LOGITS>(c, p, kv->gemv_grid, st));
    else TRY(launch_gemv_ring<GEMV_LOGITS>(c, p, true, st));
  }
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// prefill
// ------------------------------------------------------------------------------------------------
static int launch_prefill_attention(vly_ctx* c, vly_kv* kv, const bf16* qbuf, int B, int S, int past, int layer, bf16* out, cudaStream_t st) {
  const vly_config& g = c->cfg;
  const int H = g.hidden_size, nH = g.num_attention_heads;
  TRY(ensure_smem_attr(c->cfg.device, llama_prefill_attention_kernel, PrefillAttnCfg::SMEM_BYTES));
  CUtensorMap tq, tk, tv;
  TRY(make_tmap_2d(c, &tq, qbuf, H, (uint64_t)B * S, (uint64_t)H * 2, 64, 128));
  TRY(make_tmap_3d(c, &tk, kv->k_layer(layer), 128, kv->Smax, (uint64_t)B * nH, 256, (uint64_t)kv->Smax * 256, 64, 128));
  TRY(make_tmap_3d(c, &tv, kv->v_layer(layer), 128, kv->Smax, (uint64_t)B * nH, 256, (uint64_t)kv->Smax * 256, 64, 128));
  PrefillAttnParams p;
  p.B = B; p.S = S; p.past = past; p.nH = nH; p.H = H; p.Smax = kv->Smax; p.ctx = out;
  p.scale_log2e = 0.08838834764831845f * 1.4426950408889634f;
  p.key_bits = kv->masked ? kv->key_bits : nullptr; p.mask_words = kv->mask_words();
  const int n_qt = cdiv(S, 128);
  CK(launch_ex(llama_prefill_attention_kernel, dim3(B * nH * n_qt), dim3(PrefillAttnCfg::THREADS), PrefillAttnCfg::SMEM_BYTES, st, true, tq, tk, tv, p));
  c->launches++;
  return VLY_OK;
}

extern "C" int vly_llama_prefill(vly_ctx* c, vly_kv* kv, const void* inputs_embeds, int B, int S, int logits_mode, void* logits_dev,
                                 int64_t* next_tokens_dev, void* stream) {
  if (!c || !kv || !inputs_embeds || B <= 0 || S <= 0) return fail(VLY_ERR_INVALID, "vly_llama_prefill: bad argument");
  if (kv->ctx != c || B != kv->B) return fail(VLY_ERR_INVALID, "vly_llama_prefill: batch %d does not match the kv cache (%d)", B, kv->B);
  if (logits_mode < 0 || logits_mode > 2 || (logits_mode && !logits_dev)) return fail(VLY_ERR_INVALID, "vly_llama_prefill: bad logits mode");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->finalized || !c->has_llm) return fail(VLY_ERR_STATE, "vly_llama_prefill: LLM weights not finalised");
  TRY(sync_len(kv));
  const int past = kv->host_len;
  if (past + S > kv->Smax) return fail(VLY_ERR_INVALID, "vly_llama_prefill: %d cached + %d new tokens exceed the cache capacity %d", past, S, kv->Smax);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const vly_config& g = c->cfg;
  const int H = g.hidden_size, nH = g.num_attention_heads, I = g.intermediate_size, V = g.vocab_size;
  const int M = B * S;
  const int bn_h = pick_bn_m(c, H, M), nt = cdiv(H, bn_h);
  TRY(ensure(c->w_x, (size_t)M * H * 2));
  TRY(ensure(c->w_q, (size_t)M * H * 2));
  TRY(ensure(c->w_attn, (size_t)M * H * 2));
  TRY(ensure(c->w_hb, (size_t)M * I * 2));
  TRY(ensure(c->w_pstats, (size_t)M * nt * sizeof(float2)));
  bf16 *x = (bf16*)c->w_x.p, *qb = (bf16*)c->w_q.p, *attn = (bf16*)c->w_attn.p, *hb = (bf16*)c->w_hb.p;
  float2* stats = (float2*)c->w_pstats.p;
  copy_rows_stats_kernel<<<M, 128, 0, st>>>((const bf16*)inputs_embeds, x, stats, nt, H);
  c->launches++;
  CKL();
  for (int l = 0; l < g.num_hidden_layers; ++l) {
    const LlamaLayerW& w = c->layers[l];
    {
      GemmParams p = {};
      p.M = M; p.N = 3 * H; p.K = H; p.out = qb; p.ldo = H;
      p.stats_in = stats; p.stats_in_nt = nt; p.inv_dim = 1.f / H; p.eps = g.rms_norm_eps;
      p.rope = c->rope; p.S = S; p.past = past; p.H = H; p.nH = nH; p.Smax = kv->Smax;
      p.kcache = kv->k_layer(l); p.vcache = kv->v_layer(l);
      TRY(launch_gemm<EPI_RMS_QKV_ROPE>(c, pick_bn_m(c, 3 * H, M), x, H, w.wqkv, H, p, st));
    }
    TRY(launch_prefill_attention(c, kv, qb, B, S, past, l, attn, st));
    {
      GemmParams p = {};
      p.M = M; p.N = H; p.K = H; p.out = x; p.ldo = H; p.residual = x; p.ldr = H; p.stats_out = stats;
      TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, bn_h, attn, H, w.wo, H, p, st));
    }
    {
      GemmParams p = {};
      p.M = M; p.N = 2 * I; p.K = H; p.out = hb; p.ldo = I;
      p.stats_in = stats; p.stats_in_nt = nt; p.inv_dim = 1.f / H; p.eps = g.rms_norm_eps;
      TRY(launch_gemm<EPI_RMS_SWIGLU>(c, pick_bn_m(c, 2 * I, M), x, H, w.wgu, H, p, st));
    }
    {
      GemmParams p = {};
      p.M = M; p.N = H; p.K = I; p.out = x; p.ldo = H; p.residual = x; p.ldr = H; p.stats_out = stats;
      TRY(launch_gemm<EPI_BIAS_RES_STATS>(c, bn_h, hb, I, w.wdown, I, p, st));
    }
  }
  if (logits_mode == 2) {  // lm_head over every position (valley_model.py:304-305)
    GemmParams p = {};
    p.M = M; p.N = V; p.K = H; p.out = logits_dev; p.ldo = V;
    p.stats_in = stats; p.stats_in_nt = nt; p.inv_dim = 1.f / H; p.eps = g.rms_norm_eps;
    TRY(launch_gemm<EPI_RMS_F32>(c, 256, x, H, c->lm_head, H, p, st));
  }
  // last position only: final RMSNorm (folded) + lm_head GEMV + greedy argmax (model_worker.py:389-391)
  for (int b0 = 0; b0 < B; b0 += 4) {
    const int nb = (B - b0) < 4 ? (B - b0) : 4;
    GemvParams p = {};
    p.N = V; p.K = H; p.B = nb; p.W = c->lm_head; p.x = x + ((size_t)b0 * S + (S - 1)) * H; p.ldx = (long long)S * H;
    p.eps = g.rms_norm_eps;
    p.logits = kv->logits + (size_t)b0 * V;
    p.part_val = kv->part_val + (size_t)b0 * kv->gemv_grid;
    p.part_idx = kv->part_idx + (size_t)b0 * kv->gemv_grid;
    p.counter = kv->counters + (size_t)kv->B * nH;
    p.next_tokens = kv->cur_tokens + b0;
    p.out_tokens = nullptr;
    p.step = kv->d_step; p.seq_len_rw = kv->d_len; p.bump = 0;
    if (use_decode_v1()) TRY(launch_gemv<GEMV_LOGITS>(c, p, kv->gemv_grid, st));
    else TRY(launch_gemv_ring<GEMV_LOGITS>(c, p, false, st));
  }
  if (logits_mode == 1) CK(cudaMemcpyAsync(logits_dev, kv->logits, (size_t)B * V * 4, cudaMemcpyDeviceToDevice, st));
  if (next_tokens_dev) CK(cudaMemcpyAsync(next_tokens_dev, kv->cur_tokens, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
  kv->host_len = past + S;
  set_int_kernel<<<1, 1, 0, st>>>(kv->d_len, kv->host_len);
  c->launches++;
  CKL();
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
static long long* g_mega_dbg = nullptr;
extern "C" int vly_debug_mega_counters(long long* host_out, int n) {
  if (!g_mega_dbg) return -1;
  return cudaMemcpy(host_out, g_mega_dbg, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
static int launch_decode_mega(vly_ctx* c, vly_kv* kv, cudaStream_t st) {
  const vly_config& g = c->cfg;
  const int B = kv->B, bmax = B <= 1 ? 1 : (B <= 2 ? 2 : 4);
  StepParams p = {};
  p.phases = kv->d_phases; p.n_phases = kv->n_phases;
  p.B = B; p.H = g.hidden_size; p.nH = g.num_attention_heads; p.Smax = kv->Smax; p.V = g.vocab_size;
  p.Kmax = g.intermediate_size > g.hidden_size ? g.intermediate_size : g.hidden_size;
  p.eps = g.rms_norm_eps; p.scale_log2e = 0.08838834764831845f * 1.4426950408889634f;
  p.rope = c->rope; p.seq_len = kv->d_len; p.step = kv->d_step; p.embed = c->embed; p.tokens_in = kv->cur_tokens;
  p.x = kv->x; p.q = kv->q; p.attn = kv->attn;
  p.part_o = kv->part_o; p.part_ml = kv->part_ml; p.attn_counters = kv->counters; p.nsplit = kv->nsplit;
  p.key_bits = kv->key_bits; p.mask_words = kv->mask_words();
  p.logits = kv->logits; p.part_val = kv->part_val; p.part_idx = kv->part_idx;
  p.next_tokens = kv->cur_tokens; p.out_tokens = kv->gen_tokens; p.out_stride = kv->Smax;
  p.grid_counter = kv->grid_counter;
  p.grid_epoch = kv->grid_counter + 1;
  p.n_grid_syncs = kv->n_grid_syncs;
  {
    static const int env_ik = getenv("VLY_ATTN_IKEYS") ? atoi(getenv("VLY_ATTN_IKEYS")) : 0;
    p.attn_ikeys = (env_ik >= 16 && env_ik <= 256 && env_ik % 16 == 0) ? env_ik : 0;
  }
  p.sample = kv->d_sample;
  {
    static const bool want = getenv("VLY_MEGA_DBG") != nullptr;
    p.dbg = want ? kv->dbg : nullptr;
    g_mega_dbg = kv->dbg;
  }
  if (kv->umma) {
    // tcgen05 consumer: swizzled activation block + ring of whole-box stages + barriers / handoff slots
    p.Kmax = kv->umma_x_cols;
    const size_t x_bytes = (size_t)(kv->umma_x_cols / 64) * 1024, stage_b = (size_t)kv->stage_bytes;
    const size_t misc = ((1 + UmmaCfg::ISSUERS) * UmmaCfg::MAX_STAGES + 2 * UmmaCfg::ACC_SLOTS + 2 * UmmaCfg::RED_SLOTS + 2) * 8 + 16 +
                        (size_t)UmmaCfg::RED_SLOTS * 4 * UmmaCfg::ROWS * bmax * 4 + (3 + 16) * bmax * 4 + 256;
    int n_stages = (int)(((long long)227 * 1024 - 1024 - (long long)x_bytes - (long long)misc) / (long long)stage_b);
    if (n_stages > UmmaCfg::MAX_STAGES) n_stages = UmmaCfg::MAX_STAGES;
    static const int want = getenv("VLY_MEGA_STAGES") ? atoi(getenv("VLY_MEGA_STAGES")) : 0;
    if (want > 0 && n_stages > want) n_stages = want;
    if (n_stages < 2) return fail(VLY_ERR_INVALID, "decode (tcgen05 consumer): no room for the weight ring");
    static const int inflight = getenv("VLY_MEGA_INFLIGHT") ? atoi(getenv("VLY_MEGA_INFLIGHT")) : 0;
    p.n_stages = n_stages;
    p.stage_bytes = (int)stage_b;
    p.n_inflight = (inflight > 0 && inflight < n_stages) ? inflight : n_stages;
    const size_t smem = x_bytes + (size_t)n_stages * stage_b + misc;
    void* args[] = {&p};
    cudaError_t e;
    if (bmax == 2) {
      TRY(ensure_smem_attr(c->cfg.device, decode_step_umma_kernel<2>, smem));
      e = cudaLaunchCooperativeKernel((void*)decode_step_umma_kernel<2>, dim3(c->num_sms), dim3(MegaCfg::THREADS), args, smem, st);
    } else {
      TRY(ensure_smem_attr(c->cfg.device, decode_step_umma_kernel<4>, smem));
      e = cudaLaunchCooperativeKernel((void*)decode_step_umma_kernel<4>, dim3(c->num_sms), dim3(MegaCfg::THREADS), args, smem, st);
    }
    c->launches++;
    CK(e);
    return VLY_OK;
  }
  size_t x_bytes, misc;
  mega_smem_layout(g, bmax, &x_bytes, &misc);
  const size_t stage_b = (size_t)kv->stage_bytes;
  int n_stages = (int)(((long long)kMegaSmem - (long long)x_bytes - (long long)misc) / (long long)stage_b);
  if (n_stages > MegaCfg::MAX_STAGES) n_stages = MegaCfg::MAX_STAGES;
  {
    // depth of the ring and number of stages kept in flight: see pick_phase_geometry.  VLY_MEGA_STAGES / VLY_MEGA_INFLIGHT override.
    static const int want = getenv("VLY_MEGA_STAGES") ? atoi(getenv("VLY_MEGA_STAGES")) : 0;
    const int cap = want > 0 ? want : 4;            // deeper rings made every grid barrier slower (see pick_phase_geometry)
    if (n_stages > cap) n_stages = cap;
    static const int inflight = getenv("VLY_MEGA_INFLIGHT") ? atoi(getenv("VLY_MEGA_INFLIGHT")) : 0;
    p.n_inflight = inflight > 0 ? inflight : n_stages;                        // (PhaseDesc::inflight is the per-phase value)
    if (p.n_inflight > n_stages) p.n_inflight = n_stages;
  }
  if (n_stages < 2) return fail(VLY_ERR_INVALID, "decode: activations (B=%d, K=%d) leave no room for the weight ring", B, p.Kmax);
  p.n_stages = n_stages;
  p.stage_bytes = (int)stage_b;
  const size_t smem = (size_t)n_stages * stage_b + x_bytes + misc;
  void* args[] = {&p};
  cudaError_t e;
#define VLY_MEGA_CASE(BM)                                                                                               \
  {                                                                                                                     \
    TRY(ensure_smem_attr(c->cfg.device, decode_step_kernel<BM>, smem));                                                 \
    e = cudaLaunchCooperativeKernel((void*)decode_step_kernel<BM>, dim3(c->num_sms), dim3(MegaCfg::THREADS), args, smem, st); \
  }
  if (bmax == 1) VLY_MEGA_CASE(1)
  else if (bmax == 2) VLY_MEGA_CASE(2)
  else VLY_MEGA_CASE(4)
#undef VLY_MEGA_CASE
  c->launches++;
  CK(e);
  return VLY_OK;
}

static int enqueue_full_step(vly_ctx* c, vly_kv* kv, cudaStream_t st) {
  if (decode_mode() == 2 && kv->B <= 4) return launch_decode_mega(c, kv, st);
  for (int b0 = 0; b0 < kv->B; b0 += 4) {
    const int nb = (kv->B - b0) < 4 ? (kv->B - b0) : 4;
    TRY(enqueue_decode_step(c, kv, b0, nb, b0 + 4 >= kv->B, st));
  }
  // per-op paths: temperature sampling / eos bookkeeping as one more launch over the step's logits (returns at once when greedy)
  if (kv->B <= kMaxSampleRows) {
    sample_rows_kernel<<<1, 1024, 0, st>>>(kv->logits, kv->B, c->cfg.vocab_size, kv->d_sample, kv->d_len, kv->d_step, kv->cur_tokens,
                                           kv->gen_tokens, kv->Smax, 0);
    c->launches++;
    CKL();
  }
  return VLY_OK;
}

// ---- token selection state ----
__global__ void set_sample_state_kernel(SampleState* s, float inv_temp, int enabled, uint32_t k0, uint32_t k1, long long eos, long long pad,
                                        long long stop2, int reset_done) {
  if (threadIdx.x == 0) {
    s->inv_temp = inv_temp; s->enabled = enabled; s->seed_lo = k0; s->seed_hi = k1; s->eos = eos; s->pad = pad; s->stop2 = stop2;
    if (reset_done) { s->all_done = 0; s->steps_valid = 0; }
  }
  if (reset_done && threadIdx.x < kMaxSampleRows) s->done[threadIdx.x] = 0;
}

// sampling == nullptr: plain greedy, no stop token (skipped when the device state already says so)
static int set_sampling(vly_ctx* c, vly_kv* kv, const vly_sampling* sp, bool reset_done, cudaStream_t st) {
  if (!sp) {
    if (!kv->sample_dirty) return VLY_OK;
    set_sample_state_kernel<<<1, 64, 0, st>>>(kv->d_sample, 1.f, 0, 0, 0, -1, 0, -1, 1);
    kv->sample_dirty = false;
  } else {
    if (kv->B > kMaxSampleRows) return fail(VLY_ERR_INVALID, "sampling / eos bookkeeping supports at most %d sequences per cache", kMaxSampleRows);
    const bool on = sp->temperature >= 1e-4f;         // model_worker.py:390: below that the reference takes the arg-max
    set_sample_state_kernel<<<1, 64, 0, st>>>(kv->d_sample, on ? 1.f / sp->temperature : 1.f, on ? 1 : 0, (uint32_t)sp->seed,
                                              (uint32_t)(sp->seed >> 32), sp->eos_token_id < 0 ? -1 : sp->eos_token_id, sp->pad_token_id,
                                              sp->stop_token_id < 0 ? -1 : sp->stop_token_id, reset_done ? 1 : 0);
    kv->sample_dirty = true;
  }
  c->launches++;
  CKL();
  return VLY_OK;
}

constexpr int kGraphSteps = 8;
static int capture_steps(vly_ctx* c, vly_kv* kv, int n, cudaGraphExec_t* out) {
  const int64_t before = c->launches;
  CK(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
  int r = VLY_OK;
  for (int i = 0; i < n && r == VLY_OK; ++i) r = enqueue_full_step(c, kv, c->cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->cap_stream, &graph);
  kv->graph_nodes = (int)((c->launches - before) / n);
  c->launches = before;
  if (r != VLY_OK) {
    if (graph) cudaGraphDestroy(graph);
    return r;
  }
  if (e != cudaSuccess) return fail(VLY_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  const cudaError_t e2 = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) return fail(VLY_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e2));
  return VLY_OK;
}
static int build_graph(vly_ctx* c, vly_kv* kv) {
  if (kv->graph) return VLY_OK;
  TRY(capture_steps(c, kv, 1, &kv->graph));
  static const bool multi = getenv("VLY_GRAPH_STEPS1") == nullptr;
  if (multi) TRY(capture_steps(c, kv, kGraphSteps, &kv->graph_n));
  return VLY_OK;
}

extern "C" int vly_sample_logits(vly_ctx* c, vly_kv* kv, const float* logits, const vly_sampling* sp, int64_t* tokens_out, void* stream) {
  if (!c || !kv || !logits || !sp || !tokens_out || kv->ctx != c) return fail(VLY_ERR_INVALID, "vly_sample_logits: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  TRY(set_sampling(c, kv, sp, true, st));
  sample_rows_kernel<<<1, 1024, 0, st>>>(logits, kv->B, c->cfg.vocab_size, kv->d_sample, kv->d_len, kv->d_step, (long long*)tokens_out, nullptr, 0, 1);
  c->launches++;
  CKL();
  return VLY_OK;
}

static int generate_impl(vly_ctx* c, vly_kv* kv, const int64_t* first_tokens, int n_steps, int64_t* out_tokens, const vly_sampling* sp,
                         int* steps_done_dev, void* stream);

extern "C" int vly_generate_greedy(vly_ctx* c, vly_kv* kv, const int64_t* first_tokens, int n_steps, int64_t* out_tokens, void* stream) {
  return generate_impl(c, kv, first_tokens, n_steps, out_tokens, nullptr, nullptr, stream);
}

extern "C" int vly_generate(vly_ctx* c, vly_kv* kv, const int64_t* first_tokens, int n_steps, int64_t* out_tokens, const vly_sampling* sp,
                            int* steps_done_dev, void* stream) {
  return generate_impl(c, kv, first_tokens, n_steps, out_tokens, sp, steps_done_dev, stream);
}

static int generate_impl(vly_ctx* c, vly_kv* kv, const int64_t* first_tokens, int n_steps, int64_t* out_tokens, const vly_sampling* sp,
                         int* steps_done_dev, void* stream) {
  if (!c || !kv || !first_tokens || n_steps <= 0 || kv->ctx != c) return fail(VLY_ERR_INVALID, "vly_generate: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  TRY(sync_len(kv));
  if (kv->host_len + n_steps > kv->Smax) return fail(VLY_ERR_INVALID, "vly_generate: %d cached + %d steps exceed the cache capacity %d", kv->host_len, n_steps, kv->Smax);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  static const bool no_graph = getenv("VLY_NO_GRAPH") != nullptr;   // profiling aid: eager launches instead of graph replay
  if (!no_graph) TRY(build_graph(c, kv));
  // with a sampling struct the eos flags raised by vly_sample_logits (the first token) are kept; greedy starts clean
  TRY(set_sampling(c, kv, sp, false, st));
  CK(cudaMemcpyAsync(kv->cur_tokens, first_tokens, (size_t)kv->B * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(kv->d_step, 0, 4, st));
  if (steps_done_dev) CK(cudaMemsetAsync(&kv->d_sample->steps_valid, 0, 4, st));
  for (int i = 0; i < n_steps;) {
    if (no_graph) { TRY(enqueue_full_step(c, kv, st)); ++i; }
    else if (kv->graph_n && n_steps - i >= kGraphSteps) { CK(cudaGraphLaunch(kv->graph_n, st)); i += kGraphSteps; }
    else { CK(cudaGraphLaunch(kv->graph, st)); ++i; }
  }
  if (!no_graph) c->launches += (int64_t)n_steps * kv->graph_nodes;
  if (out_tokens)
    CK(cudaMemcpy2DAsync(out_tokens, (size_t)n_steps * 8, kv->gen_tokens, (size_t)kv->Smax * 8, (size_t)n_steps * 8, kv->B,
                         cudaMemcpyDeviceToDevice, st));
  if (steps_done_dev)      // (zeroed below, before the first step)
    CK(cudaMemcpyAsync(steps_done_dev, &kv->d_sample->steps_valid, 4, cudaMemcpyDeviceToDevice, st));
  kv->host_len += n_steps;
  if (sp && (sp->eos_token_id >= 0 || sp->stop_token_id >= 0)) {     // the loop may have stopped early: the device holds the true length
    CK(cudaMemcpyAsync(kv->h_len, kv->d_len, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(kv->len_event, st));
    kv->len_dirty = true;
  }
  return VLY_OK;
}

extern "C" int vly_llama_decode(vly_ctx* c, vly_kv* kv, const int64_t* tokens, int64_t* next_tokens, void* logits_dev, void* stream) {
  if (!c || !kv || !tokens || kv->ctx != c) return fail(VLY_ERR_INVALID, "vly_llama_decode: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  TRY(sync_len(kv));
  if (kv->host_len + 1 > kv->Smax) return fail(VLY_ERR_INVALID, "vly_llama_decode: cache full (%d)", kv->Smax);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  TRY(build_graph(c, kv));
  TRY(set_sampling(c, kv, nullptr, true, st));
  CK(cudaMemcpyAsync(kv->cur_tokens, tokens, (size_t)kv->B * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(kv->d_step, 0, 4, st));
  CK(cudaGraphLaunch(kv->graph, st));
  c->launches += kv->graph_nodes;
  if (next_tokens) CK(cudaMemcpyAsync(next_tokens, kv->cur_tokens, (size_t)kv->B * 8, cudaMemcpyDeviceToDevice, st));
  if (logits_dev) CK(cudaMemcpyAsync(logits_dev, kv->logits, (size_t)kv->B * c->cfg.vocab_size * 4, cudaMemcpyDeviceToDevice, st));
  kv->host_len += 1;
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// shifted cross-entropy over [B,S,V] logits (valley_model.py:308-318)
// ------------------------------------------------------------------------------------------------
extern "C" int vly_cross_entropy(vly_ctx* c, const float* logits, const int64_t* labels, int B, int S, int64_t ignore_index, float* loss_out,
                                 void* stream) {
  if (!c || !logits || !labels || !loss_out || B <= 0 || S < 2) return fail(VLY_ERR_INVALID, "vly_cross_entropy: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * (S - 1);
  TRY(ensure(c->w_score, (size_t)rows * 8));
  float* nll = (float*)c->w_score.p;
  int* cnt = (int*)(nll + rows);
  ce_rows_kernel<<<rows, 256, 0, st>>>(logits, (const long long*)labels, S, c->cfg.vocab_size, ignore_index, nll, cnt);
  ce_mean_kernel<<<1, 1024, 0, st>>>(nll, cnt, rows, loss_out);
  c->launches += 2;
  CKL();
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// frame preprocessing (load_video's Resize(256) -> CenterCrop(224) -> /255 -> CLIP mean/std), SURVEY 8 f-2
// ------------------------------------------------------------------------------------------------
extern "C" int vly_preprocess_frames(vly_ctx* c, const uint8_t* frames, int T, int H, int W, int out_dtype, void* out, void* stream) {
  if (!c || !frames || !out || T <= 0 || H <= 0 || W <= 0 || out_dtype < VLY_F32 || out_dtype > VLY_F16)
    return fail(VLY_ERR_INVALID, "vly_preprocess_frames: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  PreprocParams& p = c->pre;
  if (c->pre_H != H || c->pre_W != W) {      // new geometry: rebuild the fixed-point tables on the host (exact, Pillow's arithmetic)
    int nh, nw, cy, cx, kh = 0, kv = 0;
    TRY(vly_preprocess_plan(H, W, &nh, &nw, &cy, &cx));
    TRY(vly_resample_coeffs(W, nw, &kh, nullptr, nullptr, nullptr));
    TRY(vly_resample_coeffs(H, nh, &kv, nullptr, nullptr, nullptr));
    std::vector<int32_t> tab((size_t)nw * (2 + kh) + (size_t)nh * (2 + kv));
    int32_t* xmin = tab.data(); int32_t* xcnt = xmin + nw; int32_t* xk = xcnt + nw;
    int32_t* ymin = xk + (size_t)nw * kh; int32_t* ycnt = ymin + nh; int32_t* yk = ycnt + nh;
    TRY(vly_resample_coeffs(W, nw, &kh, xmin, xcnt, xk));
    TRY(vly_resample_coeffs(H, nh, &kv, ymin, ycnt, yk));
    int row0 = H, row1 = 0;
    for (int y = cy; y < cy + 224; ++y) {
      row0 = std::min(row0, ymin[y]);
      row1 = std::max(row1, ymin[y] + ycnt[y]);
    }
    TRY(ensure(c->w_tables, tab.size() * 4));
    CK(cudaStreamSynchronize(st));           // the previous geometry's tables may still be in use on this stream
    CK(cudaMemcpy(c->w_tables.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    const int32_t* d = (const int32_t*)c->w_tables.p;
    p.H = H; p.W = W; p.row0 = row0; p.rows = row1 - row0; p.crop_x = cx; p.crop_y = cy; p.ksize_h = kh; p.ksize_v = kv;
    p.xmin_h = d; p.cnt_h = d + nw; p.kk_h = d + 2 * (size_t)nw;
    p.ymin_v = p.kk_h + (size_t)nw * kh; p.cnt_v = p.ymin_v + nh; p.kk_v = p.cnt_v + nh;
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, sd[3] = {0.26862954f, 0.26130258f, 0.27577711f};   // data_util.py:272-273
    for (int i = 0; i < 3; ++i) { p.mean[i] = mean[i]; p.std[i] = sd[i]; }
    c->pre_H = H; c->pre_W = W;
  }
  TRY(ensure(c->w_strip, (size_t)T * p.rows * 224 * 3));
  p.frames = frames; p.T = T; p.strip = (uint8_t*)c->w_strip.p; p.out = out; p.out_dtype = out_dtype;
  preprocess_horizontal_kernel<<<dim3(p.rows, T), 224, 0, st>>>(p);
  CKL();
  if (out_dtype == VLY_F32) preprocess_vertical_kernel<float><<<dim3(224, T), 224, 0, st>>>(p);
  else if (out_dtype == VLY_F16) preprocess_vertical_kernel<__half><<<dim3(224, T), 224, 0, st>>>(p);
  else preprocess_vertical_kernel<__nv_bfloat16><<<dim3(224, T), 224, 0, st>>>(p);
  CKL();
  c->launches += 2;
  return VLY_OK;
}

// ------------------------------------------------------------------------------------------------
// per-kernel test hooks
// ------------------------------------------------------------------------------------------------
extern "C" int vly_test_gemm(vly_ctx* c, const void* a, const void* w, int M, int N, int K, int epi, const float* bias, const void* residual,
                             void* out, int block_n, void* stream) {
  if (!c || !a || !w || !out || M <= 0 || N <= 0 || K <= 0 || (K % 8) || (block_n != 128 && block_n != 256))
    return fail(VLY_ERR_INVALID, "vly_test_gemm: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = N; p.bias = bias;
  if (epi == EPI_BIAS) return launch_gemm<EPI_BIAS>(c, block_n, (const bf16*)a, K, (const bf16*)w, K, p, (cudaStream_t)stream);
  if (epi == EPI_BIAS_RES_STATS) {
    if (!residual || (N % 32)) return fail(VLY_ERR_INVALID, "vly_test_gemm: residual epilogue needs a residual and N %% 32 == 0");
    p.residual = (const bf16*)residual; p.ldr = N;
    return launch_gemm<EPI_BIAS_RES_STATS>(c, block_n, (const bf16*)a, K, (const bf16*)w, K, p, (cudaStream_t)stream);
  }
  return fail(VLY_ERR_INVALID, "vly_test_gemm: unsupported epilogue %d", epi);
}

extern "C" int vly_test_vit_attention(vly_ctx* c, const void* qkv, int F, void* out, void* stream) {
  if (!c || !qkv || !out || F <= 0) return fail(VLY_ERR_INVALID, "vly_test_vit_attention: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  CK(cudaSetDevice(c->cfg.device));
  return launch_vit_attention(c, (const bf16*)qkv, F, (bf16*)out, (cudaStream_t)stream);
}
