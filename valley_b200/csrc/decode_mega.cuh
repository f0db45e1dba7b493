// One persistent cooperative kernel per decode step ("mega-kernel").
//
// Decode at B <= 4 is pure weight streaming (13.2 GB / step at 7B).  Launching 5 kernels per layer leaves the HBM idle
// at every kernel boundary (launch gap, activation staging, first-load latency, tail imbalance: ~5 us x 160 per step).
// Here ONE kernel runs the whole step on one CTA per SM:
//   warp 0 (1 thread) : producer -- walks the step's weight matrices in execution order (QKV, o_proj, gate/up, down of
//                       every layer, then lm_head) and streams this CTA's rows through a shared-memory ring with 1-D bulk
//                       (TMA) copies: 4 rows x 4096 columns = 32 KB per stage, mbarrier complete_tx.  It never waits for a
//                       phase boundary -- weights do not depend on activations -- so while the consumers sit in a grid
//                       barrier the ring fills with the NEXT phase's weights and HBM keeps streaming.
//   warps 1..16       : compute -- per phase: stage the activation rows in shared memory (RMSNorm statistics where a
//                       norm is folded), multiply the ring stages (fp32 accumulate), warp-reduce, and hand the 16 per-warp
//                       partials of a 4-row unit to the finalize warp through a 4-slot mbarrier handoff -- there is NO
//                       blocking barrier inside the streaming loop.  The attention phase runs on four 128-thread teams
//                       (split-KV, last-arriver merge).  Phases are separated by a grid-wide barrier.
//   warp 17           : finalize -- sums the partials and runs the fused epilogue (RoPE + KV append, SwiGLU, residual,
//                       logits + running arg-max); its global operands are prefetched while the unit is being computed.
// Cross-CTA activations are read with ld.global.cg (L2) -- the L1 of an SM is not coherent with other SMs' writes.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
#include "decode_kernels.cuh"
#include "sampling.cuh"

namespace vly {

enum PhaseType : int { PH_QKV = 0, PH_ATTN = 1, PH_OPROJ = 2, PH_GATEUP = 3, PH_DOWN = 4, PH_LOGITS = 5 };

struct PhaseDesc {
  int type, N, K, layer;
  const __nv_bfloat16* W;       // [N, K] (nullptr for PH_ATTN)
  const __nv_bfloat16* x_in;    // activation rows [B, K]
  __nv_bfloat16* out;           // QKV: q [B,H]; OPROJ/DOWN: x [B,H] (in place, also the residual); GATEUP: hb [B,I]
  __nv_bfloat16* kcache;        // [B, nH, Smax, 128] of this layer
  __nv_bfloat16* vcache;
};

struct StepParams {
  const PhaseDesc* phases;
  int n_phases;
  int B, H, nH, Smax, V, Kmax;
  float eps, scale_log2e;
  const float2* rope;
  int* seq_len;                 // tokens in the cache before this step (== position of the new token); bumped at the end
  int* step;                    // column of out_tokens; bumped at the end
  const __nv_bfloat16* embed;
  const long long* tokens_in;   // [B]
  __nv_bfloat16* x;             // [B, H] residual stream
  const __nv_bfloat16* q;       // [B, H]
  __nv_bfloat16* attn;          // [B, H]
  float* part_o;                // [B*nH, nsplit, 128]
  float2* part_ml;              // [B*nH, nsplit]
  unsigned int* attn_counters;  // [B*nH]
  int nsplit;
  const uint32_t* key_bits;     // [B, mask_words] attention_mask, one bit per cache position (1 = attend)
  int mask_words;
  float* logits;                // [B, V]
  float* part_val;              // [B, grid]
  int* part_idx;
  long long* next_tokens;       // [B]
  long long* out_tokens;        // [B, out_stride]
  int out_stride;
  SampleState* sample;          // token selection state (greedy / temperature sampling, eos bookkeeping); sampling.cuh
  unsigned int* grid_counter;   // monotonically increasing arrival counter of the grid barrier (never reset: no memset node per step)
  unsigned int* grid_epoch;     // launches that ran to completion; barrier k of a launch waits for (epoch * n_barriers + k) * gridDim
  int n_stages;
  int n_inflight;               // bulk copies outstanding per SM are capped at this many stages (the ring may be deeper)
  long long* dbg;               // optional [gridDim][32] cycle counters: [0..3] sync, stage-x, weight loop, attention totals;
                                // [8 + 3*type + {0,1,2}] = stage-x, loop, trailing grid sync of every phase of that PhaseType
};

struct MegaCfg {
  static constexpr int ROWS = 4, KC = 4096;
  static constexpr int STAGE_BYTES = ROWS * KC * 2;   // 32 KB
  // BMAX > 1 (tensor-core consumers): a stage is 8 weight rows x 2048 columns; rows of a stage / of the activation block are
  // 64 bytes apart modulo 128, so the 16-byte fragment loads of a quarter-warp never share a bank
  static constexpr int ROWS_TC = 8, KC_TC = 2048;
  static constexpr int ROW_STRIDE_TC = KC_TC * 2 + 64;
  static constexpr int STAGE_BYTES_TC = ROWS_TC * ROW_STRIDE_TC;
  static constexpr int CONSUMERS = 512, THREADS = 576;   // producer warp + 16 compute warps + 1 finalize warp
  static constexpr int RED_SLOTS = 4;
  static constexpr int MAX_STAGES = 6;
  static constexpr int ATTN_SCRATCH = 4 * (64 + 8 * 128 + 8) * 4;   // per team: scores[64] + redg[8][128] + wr
};

VLY_DEVINL uint4 ldcg_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
VLY_DEVINL float ldcg_bf16(const __nv_bfloat16* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];\n" : "=h"(v) : "l"(p) : "memory");
  return __uint_as_float((uint32_t)v << 16);
}
// D[16x8] += A[16x16] * B[16x8], bf16 inputs, fp32 accumulate (legacy tensor path: plenty for an HBM-bound consumer)
VLY_DEVINL void mma_m16n8k16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

VLY_DEVINL uint32_t ld_acquire_u32(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// consumers-only grid barrier (512 threads per CTA take part; the producer warp streams on)
VLY_DEVINL void grid_sync_consumers(unsigned int* counter, unsigned int target, int ct) {
  asm volatile("bar.sync 2, 544;" ::: "memory");          // every write of this CTA happens-before thread 0's release
  if (ct == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    while ((int)(ld_acquire_u32(counter) - target) < 0) {     // wrap-safe
    }
  }
  asm volatile("bar.sync 2, 544;" ::: "memory");
}

template <int BMAX>
__global__ void __launch_bounds__(576, 1) decode_step_kernel(const StepParams p) {
  using M = MegaCfg;
  constexpr bool kTC = BMAX > 1;                                                         // tensor-core consumers
  constexpr int ROWS = kTC ? M::ROWS_TC : M::ROWS;                                       // weight rows per work unit
  constexpr int KC = kTC ? M::KC_TC : M::KC;                                             // columns per ring stage
  constexpr int NV = ROWS * BMAX;
  constexpr int ROW_STRIDE = kTC ? M::ROW_STRIDE_TC : M::KC * 2;                         // bytes between rows of a ring stage
  constexpr int STAGE_B = kTC ? M::STAGE_BYTES_TC : M::STAGE_BYTES;
  extern __shared__ uint8_t msm_raw[];
  uint8_t* msm = msm_raw + ((128u - (smem_u32(msm_raw) & 127u)) & 127u);
  uint8_t* ring = msm;                                                                   // [n_stages][4][ROW_STRIDE]
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ring + (size_t)p.n_stages * STAGE_B);   // [BMAX][xs_stride]
  // activation rows: stride == Kmax for the CUDA-core path; == 64 bytes modulo 128 for the tensor-core path
  const int xs_stride = kTC ? (((p.Kmax * 2 + 127) & ~127) + 64) / 2 : p.Kmax;
  // the attention teams' scratch ALIASES the activation block: x is re-staged at the start of every weight phase and is
  // dead during the attention phase (this keeps a third ring stage at 13B, B = 4 where x alone is 110 KB)
  size_t xs_bytes = ((size_t)BMAX * xs_stride * 2 + 127) & ~size_t(127);
  if (xs_bytes < (size_t)M::ATTN_SCRATCH) xs_bytes = M::ATTN_SCRATCH;
  uint8_t* tail = reinterpret_cast<uint8_t*>(xs) + xs_bytes;
  float* scratch = reinterpret_cast<float*>(xs);                                         // attention teams
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + M::MAX_STAGES;
  uint64_t* red_full = empty_bar + M::MAX_STAGES;                                        // [RED_SLOTS]
  uint64_t* red_empty = red_full + M::RED_SLOTS;                                         // [RED_SLOTS]
  float* red = reinterpret_cast<float*>(red_empty + M::RED_SLOTS);                       // [RED_SLOTS][16][NV]
  float* rstd_s = red + M::RED_SLOTS * 16 * NV;                                          // [BMAX]
  float* bestv = rstd_s + BMAX;                                                          // [BMAX]
  int* besti = reinterpret_cast<int*>(bestv + BMAX);                                     // [BMAX]
  float* wred = reinterpret_cast<float*>(besti + BMAX);                                  // [16][BMAX]
  int* team_flag = reinterpret_cast<int*>(wred + 16 * BMAX);                             // [4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.sample->all_done) return;       // every sequence has produced its stop token: the remaining replays are no-ops
  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 16);
    }
    for (int i = 0; i < M::RED_SLOTS; ++i) {
      mbar_init(&red_full[i], 16);
      mbar_init(&red_empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ======================================= producer =======================================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      // The ring is deeper than the number of copies kept in flight: ~96 KB outstanding per SM is what streams fastest
      // (tools/membw.cu), but while the consumers sit in a grid barrier / stage activations the extra slots keep HBM busy.
      int wst = 0, issued = 0;
      uint32_t wph = 0;
      for (int pi = 0; pi < p.n_phases; ++pi) {
        const PhaseDesc& d = p.phases[pi];
        if (d.type == PH_ATTN) continue;
        const int n_groups = (d.N + ROWS - 1) / ROWS;
        const int n_slices = (d.K + KC - 1) / KC;
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
          const int n0 = g * ROWS;
          const int rows = min(ROWS, d.N - n0);
          for (int s = 0; s < n_slices; ++s) {
            const int kc = min(KC, d.K - s * KC);
            mbar_wait(&empty_bar[st], ph ^ 1);
            if (issued >= p.n_inflight) {                 // copy #(issued - n_inflight) must have landed
              mbar_wait(&full_bar[wst], wph);
              if (++wst == p.n_stages) { wst = 0; wph ^= 1; }
            }
            ++issued;
            mbar_expect_tx(&full_bar[st], (uint32_t)rows * kc * 2);
            uint8_t* dst = ring + (size_t)st * STAGE_B;
            const __nv_bfloat16* src = d.W + (size_t)n0 * d.K + (size_t)s * KC;
            for (int r = 0; r < rows; ++r) bulk_load_1d(dst + r * ROW_STRIDE, src + (size_t)r * d.K, (uint32_t)kc * 2, &full_bar[st]);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    return;
  }

  // ================================ compute warps (1..16) and finalize warp (17) ================================
  const bool is_fin = (warp == 17);
  const int ct = tid - 32;            // compute thread 0..511 (finalize warp: 512..543)
  const int cw = warp - 1;            // compute warp 0..15
  const int pos = *p.seq_len;
  const bool samp_on = p.sample->enabled != 0;
  const float samp_it = p.sample->inv_temp;
  const uint32_t samp_k0 = p.sample->seed_lo, samp_k1 = p.sample->seed_hi;
  unsigned int sync_no = 0;
  // written by block 0 at the very end of the previous completed launch (stream order): the same value in every CTA
  const unsigned int sync_base = *p.grid_epoch * (unsigned int)(p.n_phases + 1) * gridDim.x;
  long long t_sync = 0, t_stage = 0, t_loop = 0, t_attn = 0, t0 = clock64();
  long long* dbg_o = (p.dbg != nullptr && ct == 0) ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;
  if (dbg_o != nullptr)
    for (int i = 8; i < 32; ++i) dbg_o[i] = 0;
  // ---- phase -1: x = embed[token] (decode input) ----
  {
    if (!is_fin) {
      const int chunks = p.B * (p.H >> 3);
      for (int i = blockIdx.x * M::CONSUMERS + ct; i < chunks; i += gridDim.x * M::CONSUMERS) {
        const int b = i / (p.H >> 3), c = i % (p.H >> 3);
        long long id = p.tokens_in[b];
        id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
        *reinterpret_cast<uint4*>(p.x + (size_t)b * p.H + c * 8) = *reinterpret_cast<const uint4*>(p.embed + (size_t)id * p.H + c * 8);
      }
    } else if (lane < BMAX) {
      bestv[lane] = -INFINITY;
      besti[lane] = 0;
    }
    grid_sync_consumers(p.grid_counter, sync_base + (++sync_no) * gridDim.x, ct);
    t_sync += clock64() - t0;
  }

  int st = 0;
  uint32_t ph = 0;
  unsigned int unit_no = 0;           // running work-unit counter of this CTA: selects the handoff slot
  for (int pi = 0; pi < p.n_phases; ++pi) {
    const PhaseDesc d = p.phases[pi];
    t0 = clock64();
    if (d.type == PH_ATTN) {
      if (!is_fin) {
        // ------------------------------ attention: 4 teams of 128 threads ------------------------------
        const int tm = cw >> 2, tt = ct & 127, tw = cw & 3;
        float* sc = scratch + tm * (64 + 8 * 128 + 8);
        float* redg = sc + 64;
        const int len = pos + 1;
        const int n_act = (len + 63) >> 6;
        const int items = p.B * p.nH * n_act;
        const int hl = lane & 15;
        // items are dealt team-major: every CTA's team 0 first, so the K/V rows of one layer are requested by as many SMs as
        // possible (an SM pulls ~50-100 GB/s; four 32 KB items on one SM were the critical path of the phase)
        for (int it = tm * gridDim.x + blockIdx.x; it < items; it += gridDim.x * 4) {
          const int split = it % n_act, bh = it / n_act;
          const int b = bh / p.nH, h = bh % p.nH;
          const int k0 = split * 64, nk = min(len, k0 + 64) - k0;
          const __nv_bfloat16* kb = d.kcache + ((size_t)bh * p.Smax) * 128;
          const __nv_bfloat16* vb = d.vcache + ((size_t)bh * p.Smax) * 128;
          float qf[8];
          {
            const uint4 w = ldcg_v4(p.q + (size_t)b * p.H + h * 128 + hl * 8);
            qf[0] = bf16_lo(w.x); qf[1] = bf16_hi(w.x); qf[2] = bf16_lo(w.y); qf[3] = bf16_hi(w.y);
            qf[4] = bf16_lo(w.z); qf[5] = bf16_hi(w.z); qf[6] = bf16_lo(w.w); qf[7] = bf16_hi(w.w);
          }
          // every K and V row this thread needs is requested up front (one L2 round trip instead of 16 serialised ones):
          // scores: half-warp per key, keys tw*2 + (lane>>4) + 8j;  P.V: 16 threads per key, keys (tt>>4) + 8j
          uint4 kw[8], vw[8];
          const int ki0 = tw * 2 + (lane >> 4), vi0 = tt >> 4, dl = tt & 15;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ik = ki0 + 8 * j, iv = vi0 + 8 * j;
            kw[j] = (ik < nk) ? ldcg_v4(kb + (size_t)(k0 + ik) * 128 + hl * 8) : make_uint4(0, 0, 0, 0);
            vw[j] = (iv < nk) ? ldcg_v4(vb + (size_t)(k0 + iv) * 128 + dl * 8) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = ki0 + 8 * j;
            const uint4 w = kw[j];
            float dd = qf[0] * bf16_lo(w.x) + qf[1] * bf16_hi(w.x) + qf[2] * bf16_lo(w.y) + qf[3] * bf16_hi(w.y) +
                       qf[4] * bf16_lo(w.z) + qf[5] * bf16_hi(w.z) + qf[6] * bf16_lo(w.w) + qf[7] * bf16_hi(w.w);
            dd += __shfl_xor_sync(0xffffffffu, dd, 8);
            dd += __shfl_xor_sync(0xffffffffu, dd, 4);
            dd += __shfl_xor_sync(0xffffffffu, dd, 2);
            dd += __shfl_xor_sync(0xffffffffu, dd, 1);
            if (i < nk && hl == 0) sc[i] = dd * p.scale_log2e;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          // attention_mask: one bit per cache position, two words per 64-key split (all ones unless the caller masked keys)
          const uint2 kbits = __ldg(reinterpret_cast<const uint2*>(p.key_bits + (size_t)b * p.mask_words + (k0 >> 5)));
          const float s0 = (lane < nk && ((kbits.x >> lane) & 1u)) ? sc[lane] : -INFINITY;
          const float s1 = (lane + 32 < nk && ((kbits.y >> lane) & 1u)) ? sc[lane + 32] : -INFINITY;
          const float mx = warp_max(fmaxf(s0, s1));
          const float e0 = s0 > -INFINITY ? fast_exp2(s0 - mx) : 0.f, e1 = s1 > -INFINITY ? fast_exp2(s1 - mx) : 0.f;
          const float l = warp_sum(e0 + e1);
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          if (tw == 0) {
            if (lane < nk) sc[lane] = e0;
            if (lane + 32 < nk) sc[lane + 32] = e1;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          {
            const int g = vi0;
            float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int i = g + 8 * j;
              const float pw = (i < nk) ? sc[i] : 0.f;
              const uint4 w = vw[j];
              o[0] = fmaf(pw, bf16_lo(w.x), o[0]); o[1] = fmaf(pw, bf16_hi(w.x), o[1]);
              o[2] = fmaf(pw, bf16_lo(w.y), o[2]); o[3] = fmaf(pw, bf16_hi(w.y), o[3]);
              o[4] = fmaf(pw, bf16_lo(w.z), o[4]); o[5] = fmaf(pw, bf16_hi(w.z), o[5]);
              o[6] = fmaf(pw, bf16_lo(w.w), o[6]); o[7] = fmaf(pw, bf16_hi(w.w), o[7]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) redg[g * 128 + dl * 8 + e] = o[e];
          }
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          float ot = 0.f;
#pragma unroll
          for (int g = 0; g < 8; ++g) ot += redg[g * 128 + tt];
          p.part_o[((size_t)bh * p.nsplit + split) * 128 + tt] = ot;
          if (tt == 0) p.part_ml[(size_t)bh * p.nsplit + split] = make_float2(mx, l);
          __threadfence();
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          if (tt == 0) team_flag[tm] = (atomicAdd(p.attn_counters + bh, 1u) == (unsigned)n_act - 1);
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");
          if (team_flag[tm]) {
            __threadfence();
            float L = 0.f, acc = 0.f;
            if (n_act <= 32) {
              // the merge sits on the critical path of the phase: one L2 round trip for the (max, sum) pairs -- lane s holds
              // split s -- and batches of eight independent loads for the partial outputs, instead of 2 n_act dependent ones
              float ms = -INFINITY, ls = 0.f;
              if (lane < n_act) {
                const float2 ml = __ldcg(&p.part_ml[(size_t)bh * p.nsplit + lane]);
                ms = ml.x;
                ls = ml.y;
              }
              const float Mx = warp_max(ms);
              const float wgt = ls > 0.f ? fast_exp2(ms - Mx) : 0.f;          // a fully masked split has m = -inf, l = 0
              L = warp_sum(ls * wgt);
              for (int s0 = 0; s0 < n_act; s0 += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  v[j] = (s0 + j < n_act) ? __ldcg(p.part_o + ((size_t)bh * p.nsplit + s0 + j) * 128 + tt) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(v[j], __shfl_sync(0xffffffffu, wgt, (s0 + j) & 31), acc);
              }
            } else {
              float Mx = -INFINITY;
              for (int s = 0; s < n_act; ++s) Mx = fmaxf(Mx, __ldcg(&p.part_ml[(size_t)bh * p.nsplit + s].x));
              for (int s = 0; s < n_act; ++s) {
                const float2 ml = __ldcg(&p.part_ml[(size_t)bh * p.nsplit + s]);
                const float w = ml.y > 0.f ? fast_exp2(ml.x - Mx) : 0.f;
                L += ml.y * w;
                acc += __ldcg(p.part_o + ((size_t)bh * p.nsplit + s) * 128 + tt) * w;
              }
            }
            p.attn[(size_t)b * p.H + h * 128 + tt] = __float2bfloat16_rn(L > 0.f ? acc / L : 0.f);
            if (tt == 0) p.attn_counters[bh] = 0;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(3 + tm) : "memory");   // scratch reuse by the next item
        }
      }
      t_attn += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * PH_ATTN + 1] += clock64() - t0;
    } else {
      // ------------------------------ weight phase ------------------------------
      const bool norm = (d.type == PH_QKV || d.type == PH_GATEUP || d.type == PH_LOGITS);
      if (!is_fin) {
        float sq[BMAX];
#pragma unroll
        for (int b = 0; b < BMAX; ++b) sq[b] = 0.f;
        const int chunks = d.K >> 3;
        for (int c = ct; c < chunks; c += M::CONSUMERS) {
#pragma unroll
          for (int b = 0; b < BMAX; ++b) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (b < p.B) w = ldcg_v4(d.x_in + (size_t)b * d.K + c * 8);
            *reinterpret_cast<uint4*>(xs + (size_t)b * xs_stride + c * 8) = w;
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = bf16_lo(ww[i]), bb = bf16_hi(ww[i]);
              sq[b] += a * a + bb * bb;
            }
          }
        }
        if (norm) {
#pragma unroll
          for (int b = 0; b < BMAX; ++b) {
            const float v = warp_sum(sq[b]);
            if (lane == 0) wred[cw * BMAX + b] = v;
          }
        }
      }
      asm volatile("bar.sync 2, 544;" ::: "memory");
      if (norm && !is_fin && ct < BMAX) {
        float t = 0.f;
        for (int w = 0; w < 16; ++w) t += wred[w * BMAX + ct];
        rstd_s[ct] = rsqrtf(t / d.K + p.eps);
      }
      asm volatile("bar.sync 2, 544;" ::: "memory");
      t_stage += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * d.type] += clock64() - t0;
      t0 = clock64();
      const int n_groups = (d.N + ROWS - 1) / ROWS;
      const int n_slices = (d.K + KC - 1) / KC;
      if (!is_fin) {
        // ===== compute warps: ring stage x activation rows -> per-warp partials -> handoff slot =====
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++unit_no) {
          const int n0 = g * ROWS;
          const int rows = min(ROWS, d.N - n0);
          const int slot = unit_no & (M::RED_SLOTS - 1);
          const uint32_t round = (unit_no / M::RED_SLOTS) & 1;
          if constexpr (kTC) {
            // ---- tensor-core consumer (B = 2..4): mma.sync m16n8k16 with A = activation rows (batch on M, rows >= B zero)
            // and B = 8 weight rows (N).  Lane (g = lane/4, t = lane%4) loads 16 contiguous bytes x[g][k..k+7] and
            // W[g][k..k+7]; both operands use the same k permutation, so two MMAs consume them.
            const int gq = lane >> 2, tq = lane & 3;
            float dacc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < n_slices; ++s) {
              const int kc = min(KC, d.K - s * KC);
              mbar_wait(&full_bar[st], ph);
              const uint8_t* wrow = ring + (size_t)st * STAGE_B + gq * ROW_STRIDE;
              const __nv_bfloat16* xrow = xs + (size_t)gq * xs_stride + (size_t)s * KC;
              const bool w_ok = gq < rows, x_ok = gq < p.B;
#pragma unroll
              for (int j = 0; j < KC / (16 * 32); ++j) {
                const int k = cw * (KC / 16) + j * 32;       // this warp's 32-wide k block (warp-uniform bound check)
                if (k < kc) {
                  uint4 wb = make_uint4(0, 0, 0, 0), xa = make_uint4(0, 0, 0, 0);
                  if (w_ok) wb = *reinterpret_cast<const uint4*>(wrow + (k + tq * 8) * 2);
                  if (x_ok) xa = *reinterpret_cast<const uint4*>(xrow + k + tq * 8);
                  mma_m16n8k16_bf16(dacc, xa.x, 0u, xa.y, 0u, wb.x, wb.y);
                  mma_m16n8k16_bf16(dacc, xa.z, 0u, xa.w, 0u, wb.z, wb.w);
                }
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty_bar[st]);
              if (++st == p.n_stages) { st = 0; ph ^= 1; }
            }
            // lane (g < B, t) holds D[batch g][weight rows 2t, 2t+1] summed over this warp's k range
            mbar_wait(&red_empty[slot], round ^ 1);
            if (gq < BMAX) {
              float* rp = red + (slot * 16 + cw) * NV;
              rp[(2 * tq) * BMAX + gq] = dacc[0];
              rp[(2 * tq + 1) * BMAX + gq] = dacc[1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_full[slot]);
            continue;
          }
          float acc[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = 0.f;
          for (int s = 0; s < n_slices; ++s) {
            const int kc = min(KC, d.K - s * KC);
            mbar_wait(&full_bar[st], ph);
            if (ct * 8 < kc) {
              const uint8_t* src = ring + (size_t)st * STAGE_B + ct * 16;
              float xf[BMAX][8];
#pragma unroll
              for (int b = 0; b < BMAX; ++b) {
                const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)b * xs_stride + (size_t)s * KC + ct * 8);
                xf[b][0] = bf16_lo(xv.x); xf[b][1] = bf16_hi(xv.x); xf[b][2] = bf16_lo(xv.y); xf[b][3] = bf16_hi(xv.y);
                xf[b][4] = bf16_lo(xv.z); xf[b][5] = bf16_hi(xv.z); xf[b][6] = bf16_lo(xv.w); xf[b][7] = bf16_hi(xv.w);
              }
#pragma unroll
              for (int r = 0; r < ROWS; ++r) {
                if (r < rows) {
                  const uint4 wv = *reinterpret_cast<const uint4*>(src + r * ROW_STRIDE);
                  const float wf[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y),
                                       bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
#pragma unroll
                  for (int b = 0; b < BMAX; ++b)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[r * BMAX + b] = fmaf(wf[e], xf[b][e], acc[r * BMAX + b]);
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
          }
          warp_reduce_scatter<NV>(acc, lane);
          mbar_wait(&red_empty[slot], round ^ 1);           // the finalize warp has drained this slot (4 units ago)
          if ((lane & (32 / NV - 1)) == 0) red[(slot * 16 + cw) * NV + lane / (32 / NV)] = acc[0];
          __syncwarp();
          if (lane == 0) mbar_arrive(&red_full[slot]);
        }
      } else {
        // ===== finalize warp: sum the 16 partials of each unit, fused epilogue =====
        constexpr unsigned kMask = (NV == 32) ? 0xffffffffu : ((1u << NV) - 1u);
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++unit_no) {
          const int n0 = g * ROWS;
          const int slot = unit_no & (M::RED_SLOTS - 1);
          const uint32_t round = (unit_no / M::RED_SLOTS) & 1;
          const int r = lane / BMAX, b = lane % BMAX, n = n0 + r;
          const bool ok = lane < NV && b < p.B && n < d.N;
          // operands of the epilogue are fetched while the compute warps are still busy with this unit
          float pre0 = 0.f, pre1 = 0.f;
          if (ok) {
            if (d.type == PH_OPROJ || d.type == PH_DOWN) pre0 = ldcg_bf16(d.out + (size_t)b * d.N + n);
            else if (d.type == PH_QKV && n < 2 * p.H) {
              const float2 cs = __ldg(p.rope + (size_t)pos * 64 + ((n & 127) >> 1));
              pre0 = cs.x;
              pre1 = cs.y;
            }
          }
          mbar_wait(&red_full[slot], round);
          float t = 0.f;
          if (lane < NV) {
#pragma unroll
            for (int w = 0; w < 16; ++w) t += red[(slot * 16 + w) * NV + lane];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&red_empty[slot]);
          if (lane < NV) {
            if (d.type == PH_OPROJ || d.type == PH_DOWN) {
              if (ok) d.out[(size_t)b * d.N + n] = __float2bfloat16_rn(t + pre0);
            } else if (d.type == PH_LOGITS) {
              const float y = t * rstd_s[b];
              if (ok && p.logits != nullptr) p.logits[(size_t)b * d.N + n] = y;
              // greedy: the logit itself; sampling: logit / T + Gumbel noise (arg-max == multinomial(softmax(logits / T)))
              float bv = ok ? (samp_on ? sample_score(y, samp_it, samp_k0, samp_k1, n, b, pos) : y) : -INFINITY;
              int bi = n;
#pragma unroll
              for (int o = BMAX; o < NV; o <<= 1) {
                const float ov = __shfl_xor_sync(kMask, bv, o);
                const int oi = __shfl_xor_sync(kMask, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
              }
              if (lane < BMAX && lane < p.B && bv > bestv[lane]) {
                bestv[lane] = bv;
                besti[lane] = bi;
              }
            } else {
              const float mine = t * rstd_s[b];
              const float other = __shfl_xor_sync(kMask, mine, BMAX);
              if ((r & 1) == 0 && ok && n + 1 < d.N) {
                float x0 = mine, x1 = other;
                if (d.type == PH_GATEUP) {
                  const float gte = bf16_round(x0), up = bf16_round(x1);      // HF:modeling_llama.py:182-184 rounds both
                  d.out[(size_t)b * (d.N >> 1) + (n >> 1)] = __float2bfloat16_rn(bf16_round(gte / (1.f + __expf(-gte))) * up);
                } else {
                  const int which = n / p.H, nh = n - which * p.H, head = nh >> 7, cidx = nh & 127;
                  if (which < 2) {
                    const float a = x0 * pre0 - x1 * pre1, c2 = x1 * pre0 + x0 * pre1;   // (cos, sin) prefetched
                    x0 = a;
                    x1 = c2;
                  }
                  __nv_bfloat16* dst;
                  if (which == 0) dst = d.out + (size_t)b * p.H + nh;
                  else dst = ((which == 1) ? d.kcache : d.vcache) + (((size_t)b * p.nH + head) * p.Smax + pos) * 128 + cidx;
                  *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(x0, x1);
                }
              }
            }
          }
        }
      }
      t_loop += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * d.type + 1] += clock64() - t0;
    }
    t0 = clock64();
    if (pi == p.n_phases - 1 && is_fin && lane < p.B) {
      p.part_val[(size_t)lane * gridDim.x + blockIdx.x] = bestv[lane];
      p.part_idx[(size_t)lane * gridDim.x + blockIdx.x] = besti[lane];
    }
    grid_sync_consumers(p.grid_counter, sync_base + (++sync_no) * gridDim.x, ct);
    t_sync += clock64() - t0;
    if (dbg_o != nullptr) dbg_o[8 + 3 * d.type + 2] += clock64() - t0;
  }
  if (dbg_o != nullptr) {
    dbg_o[0] = t_sync; dbg_o[1] = t_stage; dbg_o[2] = t_loop; dbg_o[3] = t_attn; dbg_o[4] = 0;
  }

  // ---- greedy arg-max over the per-CTA partials (lowest index on ties, like torch.argmax); advance the counters ----
  if (blockIdx.x == 0 && !is_fin) {
    if (cw < p.B) {
      const int b = cw;
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int g = lane; g < (int)gridDim.x; g += 32) {
        const float v = __ldcg(p.part_val + (size_t)b * gridDim.x + g);
        const int i = __ldcg(p.part_idx + (size_t)b * gridDim.x + g);
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        const long long tok = sample_finish_row(p.sample, b, bi);
        p.next_tokens[b] = tok;
        if (p.out_tokens != nullptr) p.out_tokens[(size_t)b * p.out_stride + *p.step] = tok;
      }
    }
    asm volatile("bar.sync 7, 512;" ::: "memory");
    if (ct == 0) {
      *p.step += 1;
      *p.seq_len += 1;
      *p.grid_epoch += 1;          // every CTA has passed the last barrier of this launch (they read the epoch at their start)
      p.sample->steps_valid += 1;
      if (p.sample->eos >= 0 || p.sample->stop2 >= 0) {
        int all = 1;
        for (int b = 0; b < p.B; ++b) all &= p.sample->done[b];
        p.sample->all_done = all;
      }
    }
  }
}

}  // namespace vly
