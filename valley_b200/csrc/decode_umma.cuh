// decode_step_umma_kernel -- the persistent decode step for B = 2..4 with a tcgen05 consumer.
//
// Why: at B > 1 the CUDA cores are issue-bound, and the legacy mma.sync pipe of this part sustains only about one m16n8k16 per 40
// cycles per SM sub-partition (tools/ringbw.cu): with the 4 activation rows padded to M = 16 the HMMA consumers of
// decode_step_kernel<4>, not HBM, bounded the step (5.7 TB/s in the weight loop, 0.64 of the copy bandwidth overall).
//
// How (validated stand-alone in tools/umma_probe.cu -- bit-level agreement with a CPU reference at 7.1 TB/s):
//   * a work unit is still 8 weight rows x K.  The producer brings a stage (8 rows x kc columns) in with ONE 3-D tensor TMA copy
//     -- tensor {64 (k inside a panel), N (rows), K / 64 (panels)}, box {64, 8, kc / 64}, SWIZZLE_128B -- which lands as
//     [panel][8 rows][128 B]: the canonical K-major UMMA operand layout.  No per-row bulk copies, no row padding.
//   * the activation rows are staged ONCE per phase into the same layout: [K / 64 panels][8 rows][128 B] (rows >= B are never
//     written: whatever they hold only reaches accumulator columns nobody reads).
//   * "diagonal-block" MMA: tcgen05.mma M = 64, N = 64, K = 16 with 8-row-group stride SBO = 1 KB on BOTH operands, so the 64 rows of
//     A are (panel c, weight row r) of 8 consecutive panels of the stage and the 64 rows of B are (panel c', batch row b) of the 8
//     matching activation panels.  D[(c, r)][(c', b)] accumulates W[r][panel c] . x[b][panel c']; the dot product that is wanted
//     is the sum over c of the diagonal blocks c = c'.  7/8 of the MACs are thrown away -- the tensor pipe has them to spare --
//     and one instruction consumes 8 rows x 512 columns x (16 / 64) = 2 KB of weights instead of the 256 B an N = 8 tile gives
//     (which is issue-bound at 2.7 TB/s).  Four MMAs (the 4 k steps of a panel) retire 8 KB of the stage.
//   * one elected thread issues; tcgen05.commit hands the ring slot back (no consumer warps touch the weights at all) and, after
//     the unit's last stage, publishes the accumulator (64 TMEM columns, 8 slots).
//   * four reader warps, one per TMEM lane quadrant (M = 64 puts D rows 16 q .. 16 q + 15 on lanes 32 q .. 32 q + 15), pull the two
//     diagonal blocks of their quadrant with one tcgen05.ld, add them and pass 8 x B partial sums to the epilogue warp through the
//     same 4-slot shared-memory handoff the other kernel uses; the fused epilogue (RoPE + KV append, SwiGLU, residual, logits +
//     arg-max / sampling) is the shared mega_unit_epilogue.
//   * a matrix whose K exceeds the activation block (down_proj: K = 13824) is walked in sub-phases of <= 5120 columns: the
//     accumulators of the CTA's (<= 8) units stay resident in TMEM across the sub-phases, the activation block is re-staged in
//     between behind a CTA-local barrier (no grid barrier), and only the last sub-phase runs the epilogue.
// Requires hidden_size and intermediate_size to be multiples of 512 (LLaMA-13B: 5120 / 13824); otherwise the host keeps
// decode_step_kernel<BMAX>.  The attention phase, the grid barrier, the sampling state and the step bookkeeping are shared code.
#pragma once
#include "decode_mega.cuh"

namespace vly {

struct UmmaCfg {
  static constexpr int ROWS = 8;                 // weight rows per work unit == the 8-row group of the UMMA operand layout
  static constexpr int ACC_SLOTS = 8;            // 64 TMEM columns each
  static constexpr int RED_SLOTS = 4;
  static constexpr int MAX_STAGES = 8;
  static constexpr int ISSUERS = 2;              // warps 1..ISSUERS (lane 0 each): issuer i takes the units u with u % ISSUERS == i.  One
                                                 // thread sustains ~1 MMA per 105 cycles (5.6 TB/s); two keep the pipe busy (~80 cycles / MMA)
  static constexpr int FIRST_READER_WARP = 13;   // warps 13..16 = TMEM lane quadrants 1, 2, 3, 0
};

template <int BMAX>
__global__ void __launch_bounds__(576, 1) decode_step_umma_kernel(const StepParams p) {
  using U = UmmaCfg;
  constexpr int NV = U::ROWS * BMAX;
  extern __shared__ __align__(1024) uint8_t usm[];
  if (smem_u32(usm) & 1023u) __trap();                                                   // SWIZZLE_128B atoms are 1 KB aligned
  uint8_t* xsw = usm;                                                                    // [Kmax / 64 panels][8 rows][128 B]
  const size_t x_bytes = (size_t)(p.Kmax >> 6) * 1024;
  uint8_t* ring = usm + x_bytes;                                                         // [n_stages][stage_bytes]
  uint8_t* tail = ring + (size_t)p.n_stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);                                // [ISSUERS][MAX_STAGES]: a fill is announced on the
  uint64_t* empty_bar = full_bar + U::ISSUERS * U::MAX_STAGES;                           // barrier of the issuer that will consume it
  uint64_t* acc_full = empty_bar + U::MAX_STAGES;                                        // [ACC_SLOTS]
  uint64_t* acc_empty = acc_full + U::ACC_SLOTS;                                         // [ACC_SLOTS]
  uint64_t* red_full = acc_empty + U::ACC_SLOTS;                                         // [RED_SLOTS]
  uint64_t* red_empty = red_full + U::RED_SLOTS;                                         // [RED_SLOTS]
  uint64_t* sub_done = red_empty + U::RED_SLOTS;                                         // the sub-phase's MMAs have all retired
  uint64_t* x_bar = sub_done + 1;                                                        // the activation block has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);                                  // [RED_SLOTS][4 quadrants][NV]
  float* rstd_s = red + U::RED_SLOTS * 4 * NV;                                           // [BMAX]
  float* bestv = rstd_s + BMAX;                                                          // [BMAX]
  int* besti = reinterpret_cast<int*>(bestv + BMAX);                                     // [BMAX]
  float* wred = reinterpret_cast<float*>(besti + BMAX);                                  // [16][BMAX]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.sample->all_done) return;       // every sequence has produced its stop token: the remaining replays are no-ops
  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      for (int j = 0; j < U::ISSUERS; ++j) mbar_init(&full_bar[j * U::MAX_STAGES + i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < U::ACC_SLOTS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < U::RED_SLOTS; ++i) {
      mbar_init(&red_full[i], 4);
      mbar_init(&red_empty[i], 1);
    }
    mbar_init(sub_done, U::ISSUERS);
    mbar_init(x_bar, 1);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================================= producer =======================================
    if (lane == 0) {
      // Work units alternate between the issuers; a fill is announced on the full barrier of the issuer that owns its unit, so every
      // barrier has ONE waiter that sees ALL of its phases in order (an mbarrier parity wait cannot tell "two phases back" from
      // "done": with one barrier per slot shared by both issuers, the one that skipped a phase could run ahead of the data or stall
      // forever).  The ring slots themselves are shared and handed out in order.
      int st = 0;
      uint32_t ph = 0;
      unsigned int unit_base = 0;
      for (int pi = 0; pi < p.n_phases; ++pi) {
        const PhaseDesc d = p.phases[pi];
        if (d.type == PH_ATTN) continue;
        const int n_groups = (d.N + U::ROWS - 1) / U::ROWS;
        const int n_slices = d.K / d.kc;
        const uint32_t stage_tx = (uint32_t)U::ROWS * d.kc * 2;     // a box is always written in full (rows past N arrive as zeros)
        const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(d.tmap);
        int u = 0;
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++u) {
          uint64_t* fb = full_bar + ((unit_base + u) % U::ISSUERS) * U::MAX_STAGES;
          for (int s = 0; s < n_slices; ++s) {
            mbar_wait(&empty_bar[st], ph ^ 1);
            mbar_expect_tx(&fb[st], stage_tx);
            tma_load_3d(ring + (size_t)st * p.stage_bytes, tm, &fb[st], 0, g * U::ROWS, (d.k_off + s * d.kc) >> 6);
            if (++st == p.n_stages) { st = 0; ph ^= 1; }
          }
        }
        if (d.flags & PHF_LAST) unit_base += (unsigned int)u;
      }
    }
  } else {
    // ================================ compute warps (1..16) and epilogue warp (17) ================================
    const bool is_fin = (warp == 17);
    const int ct = tid - 32;            // compute thread 0..511 (epilogue warp: 512..543)
    const int cw = warp - 1;            // compute warp 0..15
    const bool is_issuer = (warp >= 1 && warp <= U::ISSUERS);
    const bool is_reader = (warp >= U::FIRST_READER_WARP && warp < U::FIRST_READER_WARP + 4);
    const int quad = warp & 3;          // TMEM lane quadrant a reader warp may access
    const int pos = *p.seq_len;
    const bool samp_on = p.sample->enabled != 0;
    const float samp_it = p.sample->inv_temp;
    const uint32_t samp_k0 = p.sample->seed_lo, samp_k1 = p.sample->seed_hi;
    unsigned int sync_no = 0;
    const unsigned int sync_base = *p.grid_epoch * (unsigned int)p.n_grid_syncs * gridDim.x;
    long long t_sync = 0, t_stage = 0, t_loop = 0, t_attn = 0, t0 = clock64();
    long long* dbg_o = (p.dbg != nullptr && ct == 0) ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;
    if (dbg_o != nullptr)
      for (int i = 8; i < 32; ++i) dbg_o[i] = 0;
    // ---- phase -1: x = embed[token] (decode input) ----
    {
      if (!is_fin) {
        const int chunks = p.B * (p.H >> 3);
        for (int i = blockIdx.x * MegaCfg::CONSUMERS + ct; i < chunks; i += gridDim.x * MegaCfg::CONSUMERS) {
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

    unsigned int ring_idx = 0;          // stages the ring has carried before the current phase (the same in every thread)
    uint32_t fills = 0;                 // issuer: bit s = parity of how often ring slot s has carried one of ITS stages
    uint32_t sd_ph = 0, xph = 0;
    unsigned int unit_base = 0;         // accumulator / handoff slot of the first unit of the current matrix (fixed across its sub-phases)
    for (int pi = 0; pi < p.n_phases; ++pi) {
      const PhaseDesc d = p.phases[pi];
      t0 = clock64();
      if (d.type == PH_ATTN) {
        if (!is_fin) mega_attention_phase(p, d, cw, lane, pos);
        t_attn += clock64() - t0;
        if (dbg_o != nullptr) dbg_o[8 + 3 * PH_ATTN + 1] += clock64() - t0;
        t0 = clock64();
        grid_sync_consumers(p.grid_counter, sync_base + (++sync_no) * gridDim.x, ct);
        t_sync += clock64() - t0;
        if (dbg_o != nullptr) dbg_o[8 + 3 * PH_ATTN + 2] += clock64() - t0;
        continue;
      }
      // ------------------------------ weight (sub-)phase ------------------------------
      const bool norm = (d.type == PH_QKV || d.type == PH_GATEUP || d.type == PH_LOGITS);
      const bool first_sub = (d.flags & PHF_FIRST) != 0, last_sub = (d.flags & PHF_LAST) != 0;
      if (d.x_cols > 0) {
        // stage the activation rows [B, x_cols] (columns k_off .. of x_in) into the swizzled K-major operand layout: ONE 3-D tensor
        // TMA copy (the source buffer is 8 rows tall, rows >= B are zero), issued by one thread; everybody waits on the mbarrier
        if (!is_fin) {
          if (ct == 0) {
            fence_proxy_async_global();     // the rows were written by other CTAs (generic proxy) before the grid barrier
            mbar_expect_tx(x_bar, (uint32_t)U::ROWS * d.x_cols * 2);
            tma_load_3d(xsw, reinterpret_cast<const CUtensorMap*>(d.xmap), x_bar, 0, 0, d.k_off >> 6);
          }
          mbar_wait(x_bar, xph);
          // RMSNorm statistics: the row scale is applied in the epilogue, so the MMA issuers (warps 1, 2) start on the raw block at
          // once; warps 3..16 sum the squares and meet the epilogue warp on their own named barrier
          if (norm && warp > U::ISSUERS) {
            float sq[BMAX];
#pragma unroll
            for (int b = 0; b < BMAX; ++b) sq[b] = 0.f;
            const int chunks = d.x_cols >> 3;
            constexpr int NT = (16 - U::ISSUERS) * 32;
            for (int c = ct - U::ISSUERS * 32; c < chunks; c += NT) {
              const uint8_t* src = xsw + (size_t)(c >> 3) * 1024;
#pragma unroll
              for (int b = 0; b < BMAX; ++b) {
                if (b < p.B) {
                  const uint4 w = *reinterpret_cast<const uint4*>(src + b * 128 + (((c & 7) ^ b) << 4));
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float a = bf16_lo(ww[i]), bb = bf16_hi(ww[i]);
                    sq[b] += a * a + bb * bb;
                  }
                }
              }
            }
#pragma unroll
            for (int b = 0; b < BMAX; ++b) {
              const float v = warp_sum(sq[b]);
              if (lane == 0) wred[cw * BMAX + b] = v;
            }
          }
        }
        xph ^= 1;
        if (norm && warp > U::ISSUERS) {
          constexpr int NB = (17 - U::ISSUERS) * 32;          // warps 3..17
          asm volatile("bar.sync 3, %0;" ::"n"(NB) : "memory");
          if (is_fin && lane < BMAX) {
            float t = 0.f;
            for (int w = U::ISSUERS; w < 16; ++w) t += wred[w * BMAX + lane];
            rstd_s[lane] = rsqrtf(t / d.K + p.eps);
          }
          if (is_fin) __syncwarp();
        }
      }
      t_stage += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * d.type] += clock64() - t0;
      t0 = clock64();
      const int n_groups = (d.N + U::ROWS - 1) / U::ROWS;
      const int n_slices = d.K / d.kc;
      const int my_units = (n_groups > (int)blockIdx.x) ? (n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      if (is_issuer) {
        // ===== the MMA issuer =====
        if (lane == 0) {
          constexpr uint32_t idesc = make_idesc_bf16(64, 64);
          tc_fence_after();
          const uint32_t x_addr = smem_u32(xsw) + (uint32_t)d.x_panel0 * 1024;
          const int pgroups = d.kc >> 9;                                    // 8-panel (512-column) groups per stage
          uint64_t* fb = full_bar + (warp - 1) * U::MAX_STAGES;         // this issuer's full barriers
          for (int u = 0; u < my_units; ++u) {
            const unsigned int un = unit_base + u;
            if ((int)(un % U::ISSUERS) != warp - 1) continue;               // the other issuer's unit
            const int slot = un % U::ACC_SLOTS;
            if (first_sub) {                                                // the slot's previous accumulator has been read
              mbar_wait(&acc_empty[slot], ((un / U::ACC_SLOTS) & 1) ^ 1);
              tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + slot * 64;
            uint32_t acc = first_sub ? 0u : 1u;                             // 0: the unit's very first MMA overwrites the accumulator
            for (int s = 0; s < n_slices; ++s) {
              const unsigned int idx = ring_idx + (unsigned int)(u * n_slices + s);
              const int st = (int)(idx % (unsigned int)p.n_stages);
              mbar_wait(&fb[st], (fills >> st) & 1u);                          // parity = number of MY earlier fills of this slot
              fills ^= 1u << st;
              tc_fence_after();
              // descriptors are built once per stage; inside, only their 16-byte-unit address field advances (a panel group is
              // 8 KB = 512 units, a k step 32 B = 2 units)
              const uint64_t wd = make_smem_desc_sw128(smem_u32(ring) + (uint32_t)st * p.stage_bytes, 16, 1024);
              const uint64_t xd = make_smem_desc_sw128(x_addr + (uint32_t)(s * (d.kc >> 6)) * 1024, 16, 1024);
              for (int pg = 0; pg < pgroups; ++pg) {
                const uint64_t o = (uint64_t)pg * 512;
                tc_mma_bf16(d_tmem, wd + o, xd + o, idesc, acc);
                tc_mma_bf16(d_tmem, wd + o + 2, xd + o + 2, idesc, 1u);
                tc_mma_bf16(d_tmem, wd + o + 4, xd + o + 4, idesc, 1u);
                tc_mma_bf16(d_tmem, wd + o + 6, xd + o + 6, idesc, 1u);
                acc = 1u;
              }
              tc_commit(&empty_bar[st]);                                    // the slot is free once these MMAs have read it
            }
            if (last_sub) tc_commit(&acc_full[slot]);
          }
          if (d.flags & PHF_LOCAL_SYNC) tc_commit(sub_done);                // the activation block may be overwritten once this fires
        }
        __syncwarp();
      } else if (is_reader && last_sub) {
        // ===== TMEM readers: the two diagonal blocks of this lane quadrant -> 8 x B partial sums -> handoff slot =====
        for (int u = 0; u < my_units; ++u) {
          const unsigned int un = unit_base + u;
          const int slot = un % U::ACC_SLOTS, rs = un % U::RED_SLOTS;
          mbar_wait(&acc_full[slot], (un / U::ACC_SLOTS) & 1);
          tc_fence_after();
          uint32_t v[16];
          tmem_ld_32x16(tmem_base + (uint32_t(quad * 32) << 16) + slot * 64 + quad * 16, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);
          // lane l < 16 holds D row (panel c = 2 quad + (l >> 3), weight row r = l & 7); its diagonal block = columns (l >> 3) * 8 + b
          float tb[BMAX];
#pragma unroll
          for (int b = 0; b < BMAX; ++b) {
            tb[b] = (lane & 8) ? __uint_as_float(v[8 + b]) : __uint_as_float(v[b]);
            tb[b] += __shfl_xor_sync(0xffffffffu, tb[b], 8);                // the two panels of this quadrant
          }
          mbar_wait(&red_empty[rs], ((un / U::RED_SLOTS) & 1) ^ 1);
          if (lane < 8) {
            float* rp = red + (rs * 4 + quad) * NV + lane * BMAX;
#pragma unroll
            for (int b = 0; b < BMAX; ++b) rp[b] = tb[b];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&red_full[rs]);
        }
      } else if (is_fin && last_sub) {
        // ===== epilogue warp: sum the 4 quadrant partials of each unit, fused epilogue =====
        for (int u = 0; u < my_units; ++u) {
          const unsigned int un = unit_base + u;
          const int rs = un % U::RED_SLOTS;
          const int n0 = ((int)blockIdx.x + u * (int)gridDim.x) * U::ROWS;
          const int r = lane / BMAX, b = lane % BMAX, n = n0 + r;
          const bool ok = lane < NV && b < p.B && n < d.N;
          float pre0 = 0.f, pre1 = 0.f;
          mega_epilogue_prefetch(p, d, ok, b, n, pos, pre0, pre1);
          mbar_wait(&red_full[rs], (un / U::RED_SLOTS) & 1);
          float t = 0.f;
          if (lane < NV) {
#pragma unroll
            for (int q = 0; q < 4; ++q) t += red[(rs * 4 + q) * NV + lane];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&red_empty[rs]);
          if (lane < NV)
            mega_unit_epilogue<BMAX, NV>(p, d, lane, t, ok, r, b, n, pre0, pre1, pos, rstd_s, bestv, besti, samp_on, samp_it, samp_k0, samp_k1);
        }
      }
      if (last_sub) unit_base += (unsigned int)my_units;
      ring_idx += (unsigned int)(my_units * n_slices);
      t_loop += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * d.type + 1] += clock64() - t0;
      t0 = clock64();
      if (d.flags & PHF_NO_SYNC) {
        // the next sub-phase continues on the same staged activation block (the short tail stage of a piece): nothing to wait for
      } else if (d.flags & PHF_LOCAL_SYNC) {
        // the next sub-phase re-stages the activation block: every MMA of this one must have retired (CTA-local, no grid barrier)
        mbar_wait(sub_done, sd_ph);
        sd_ph ^= 1;
        asm volatile("bar.sync 2, 544;" ::: "memory");
      } else {
        if (pi == p.n_phases - 1 && is_fin && lane < p.B) {
          p.part_val[(size_t)lane * gridDim.x + blockIdx.x] = bestv[lane];
          p.part_idx[(size_t)lane * gridDim.x + blockIdx.x] = besti[lane];
        }
        grid_sync_consumers(p.grid_counter, sync_base + (++sync_no) * gridDim.x, ct);
      }
      t_sync += clock64() - t0;
      if (dbg_o != nullptr) dbg_o[8 + 3 * d.type + 2] += clock64() - t0;
    }
    if (dbg_o != nullptr) {
      dbg_o[0] = t_sync; dbg_o[1] = t_stage; dbg_o[2] = t_loop; dbg_o[3] = t_attn; dbg_o[4] = 0;
    }
    if (blockIdx.x == 0 && !is_fin) mega_finish_step(p, cw, lane, ct);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vly
