// SIMT kernels: layout / gather kernels around the tensor-core path, and the HBM-bound
// single-token decode path (weight-streaming GEMVs with fused RMSNorm / RoPE / KV append /
// SwiGLU / residual / argmax epilogues, split-KV attention).
#pragma once
#include "common.cuh"

namespace vly {

// ============================================================================================
// ViT front end
// ============================================================================================
// pixels [F,3,IMG,IMG] (fp32 / fp16 / bf16) -> patch matrix [F*G*G, KPAD] bf16, k = c*P*P + ky*P + kx
// (the flattening of conv weight [D,3,P,P]; HF:modeling_clip.py:208-210 casts pixels to the weight dtype).
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ px, __nv_bfloat16* __restrict__ out, int F, int IMG, int P, int KPAD) {
  const int G = IMG / P, KK = 3 * P * P;
  const int chunks = KPAD / 8;
  const long long total = (long long)F * G * G * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks);
    const long long row = i / chunks;
    const int pxi = int(row % G), pyi = int((row / G) % G), f = int(row / (G * G));
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      if (k < KK) {
        const int c = k / (P * P), rem = k % (P * P), ky = rem / P, kx = rem % P;
        v[e] = float(px[(((long long)f * 3 + c) * IMG + (pyi * P + ky)) * IMG + (pxi * P + kx)]);
      } else {
        v[e] = 0.f;
      }
    }
    *reinterpret_cast<uint4*>(out + row * KPAD + ch * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

VLY_DEVINL float block_sum_128(float v, float* red) {  // blockDim.x == 128
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// CLS concat + position embedding + pre-LayerNorm (HF:modeling_clip.py:212-219, :677), one CTA per token row.
// Writes hidden_states[0] and its row statistics (for layer 0's LN1 fold).  D == 1024, 128 threads x 8.
__global__ void __launch_bounds__(128) vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch_out,  // [F*NP, D]
                                                           const float* __restrict__ cls, const float* __restrict__ pos,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           __nv_bfloat16* __restrict__ x, float2* __restrict__ stats,
                                                           int stats_nt, int tokens, int D, float eps) {
  __shared__ float red[4];
  const int row = blockIdx.x, f = row / tokens, t = row % tokens;
  const int c0 = threadIdx.x * 8;
  float e[8];
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = bf16_round(cls[c0 + i] + pos[c0 + i]);
  } else {
    const uint4 pv = *reinterpret_cast<const uint4*>(patch_out + ((size_t)f * (tokens - 1) + (t - 1)) * D + c0);
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[2 * i] = bf16_round(bf16_lo(w[i]) + pos[(size_t)t * D + c0 + 2 * i]);
      e[2 * i + 1] = bf16_round(bf16_hi(w[i]) + pos[(size_t)t * D + c0 + 2 * i + 1]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += e[i];
  const float mean = block_sum_128(s, red) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += (e[i] - mean) * (e[i] - mean);
  const float rstd = rsqrtf(block_sum_128(ss, red) / D + eps);
  uint32_t o[4];
  float so = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = (e[2 * i] - mean) * rstd * gamma[c0 + 2 * i] + beta[c0 + 2 * i];
    const float b = (e[2 * i + 1] - mean) * rstd * gamma[c0 + 2 * i + 1] + beta[c0 + 2 * i + 1];
    o[i] = pack_bf16x2(a, b);
    const float ar = bf16_lo(o[i]), br = bf16_hi(o[i]);
    so += ar + br;
    sq += ar * ar + br * br;
  }
  *reinterpret_cast<uint4*>(x + (size_t)row * D + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  so = block_sum_128(so, red);
  sq = block_sum_128(sq, red);
  if (threadIdx.x < stats_nt) stats[(size_t)row * stats_nt + threadIdx.x] = threadIdx.x == 0 ? make_float2(so, sq) : make_float2(0.f, 0.f);
}

// ============================================================================================
// temporal pool (valley_model.py:207, :215) -- pool BEFORE projecting (the projector is linear)
// feats [NV*T, tokens, D] -> vis_in [NV, (tokens-1)+T, D]: rows 0..tokens-2 = mean over T of patch rows,
// rows tokens-1.. = CLS row of each frame.
// ============================================================================================
__global__ void temporal_pool_kernel(const __nv_bfloat16* __restrict__ feats, __nv_bfloat16* __restrict__ out, int NV, int T,
                                     int tokens, int D) {
  const int rows_out = tokens - 1 + T;
  const int chunks = D / 8;
  const long long total = (long long)NV * rows_out * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks);
    const long long ro = i / chunks;
    const int r = int(ro % rows_out), v = int(ro / rows_out);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < tokens - 1) {
      for (int t = 0; t < T; ++t) {
        const uint4 w = *reinterpret_cast<const uint4*>(feats + (((size_t)v * T + t) * tokens + (r + 1)) * D + ch * 8);
        acc[0] += bf16_lo(w.x); acc[1] += bf16_hi(w.x); acc[2] += bf16_lo(w.y); acc[3] += bf16_hi(w.y);
        acc[4] += bf16_lo(w.z); acc[5] += bf16_hi(w.z); acc[6] += bf16_lo(w.w); acc[7] += bf16_hi(w.w);
      }
      const float inv = 1.f / T;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= inv;
    } else {
      const int t = r - (tokens - 1);
      const uint4 w = *reinterpret_cast<const uint4*>(feats + (((size_t)v * T + t) * tokens) * D + ch * 8);
      acc[0] = bf16_lo(w.x); acc[1] = bf16_hi(w.x); acc[2] = bf16_lo(w.y); acc[3] = bf16_hi(w.y);
      acc[4] = bf16_lo(w.z); acc[5] = bf16_hi(w.z); acc[6] = bf16_lo(w.w); acc[7] = bf16_hi(w.w);
    }
    *reinterpret_cast<uint4*>(out + ro * D + ch * 8) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                                  pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

// ============================================================================================
// embedding gather + visual splice (valley_model.py:160, :223-247): one CTA per sequence position.
// src_map[b,s] = -1 -> token embedding row; >= 0 -> row (img_idx[b]*rows_per_img + src) of the projected visual rows.
// Also emits the row sum-of-squares for the first RMSNorm fold.
// ============================================================================================
__global__ void __launch_bounds__(128) embed_splice_kernel(const long long* __restrict__ ids, const int* __restrict__ src_map,
                                                           const int* __restrict__ img_idx,
                                                           const __nv_bfloat16* __restrict__ embed,
                                                           const __nv_bfloat16* __restrict__ vis_rows, int rows_per_img,
                                                           __nv_bfloat16* __restrict__ out, float2* __restrict__ stats,
                                                           int stats_nt, int S, int H, int vocab) {
  __shared__ float red[4];
  const int row = blockIdx.x, b = row / S;
  const int src = src_map ? src_map[row] : -1;
  const __nv_bfloat16* sp;
  if (src < 0) {
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    sp = embed + (size_t)id * H;
  } else {
    sp = vis_rows + ((size_t)img_idx[b] * rows_per_img + src) * H;
  }
  float s = 0.f, sq = 0.f;
  for (int c = threadIdx.x * 8; c < H; c += 128 * 8) {
    const uint4 w = *reinterpret_cast<const uint4*>(sp + c);
    *reinterpret_cast<uint4*>(out + (size_t)row * H + c) = w;
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = bf16_lo(ww[i]), bb = bf16_hi(ww[i]);
      s += a + bb;
      sq += a * a + bb * bb;
    }
  }
  s = block_sum_128(s, red);
  sq = block_sum_128(sq, red);
  if (stats != nullptr && threadIdx.x < stats_nt)
    stats[(size_t)row * stats_nt + threadIdx.x] = threadIdx.x == 0 ? make_float2(s, sq) : make_float2(0.f, 0.f);
}

// decode: x[b,:] = embed[token[b],:]
__global__ void decode_embed_kernel(const long long* __restrict__ tokens, const __nv_bfloat16* __restrict__ embed,
                                    __nv_bfloat16* __restrict__ x, int H, int vocab) {
  const int b = blockIdx.x;
  long long id = tokens[b];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(x + (size_t)b * H + c) = *reinterpret_cast<const uint4*>(embed + (size_t)id * H + c);
}

// ============================================================================================
// Decode GEMV family:  y[b, n] = sum_k x[b, k] * W[n, k],  B <= BMAX (1..4), W streamed once from HBM.
// CTA = 256 threads; a work unit is 8 consecutive weight rows; every thread owns 16-byte K chunks
// (chunk c -> thread c % 256), so a warp reads 512 contiguous bytes of each row.  x lives in shared
// memory (bf16), read once per chunk and reused for the 8 rows.
// ============================================================================================
enum GemvMode : int {
  GEMV_QKV_ROPE = 0,   // RMSNorm fold + RoPE + KV-cache append (HF:modeling_llama.py:262-270)
  GEMV_RESIDUAL = 1,   // y + residual -> bf16 (o_proj, down_proj)
  GEMV_SWIGLU = 2,     // RMSNorm fold + silu(g)*u with interleaved (g,u) rows
  GEMV_LOGITS = 3,     // RMSNorm fold + fp32 logits (+ fused greedy argmax, model_worker.py:390-391)
};

struct GemvParams {
  int N, K, B;
  const __nv_bfloat16* W;
  const __nv_bfloat16* x;       // [B, K], row stride ldx elements
  long long ldx;
  float eps;
  __nv_bfloat16* out;           // QKV: q [B,H]; RESIDUAL: [B,N]; SWIGLU: [B,N/2]
  const __nv_bfloat16* res;     // RESIDUAL: [B,N]
  const float2* rope;           // [max_pos, 64]
  const int* seq_len;           // device scalar: tokens already in the cache (== position of the new token)
  int H, nH, Smax;
  __nv_bfloat16* kcache;        // [B, nH, Smax, 128] (this layer)
  __nv_bfloat16* vcache;
  float* logits;                // [B, N] or nullptr
  float* part_val;              // [B, grid]
  int* part_idx;                // [B, grid]
  unsigned int* counter;
  long long* next_tokens;       // [B]
  long long* out_tokens;        // [B, out_stride] or nullptr
  int out_stride;
  int* step;                    // device scalar: decode step index (column of out_tokens)
  int* seq_len_rw;              // incremented by the last CTA of the logits kernel (end of step)
  int bump;                     // 1: this launch closes the step (advance *step and *seq_len_rw)
};

template <int BMAX, int MODE>
__global__ void __launch_bounds__(256, 2) gemv_kernel(const GemvParams p) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gsm);                           // [BMAX, K]
  float* red = reinterpret_cast<float*>(gsm + (size_t)BMAX * p.K * 2);                  // [2][8 warps][8][BMAX]
  float* fin = red + 2 * 8 * 8 * BMAX;                                                  // [2][8][BMAX]
  float* rstd_s = fin + 2 * 8 * BMAX;                                                   // [BMAX]
  float* bestv = rstd_s + BMAX;                                                         // [BMAX]
  int* besti = reinterpret_cast<int*>(bestv + BMAX);                                    // [BMAX]
  __shared__ float wred[8][BMAX];
  __shared__ int is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = p.K >> 3;

  // ---- stage x in shared memory; RMSNorm statistics where the mode folds a norm ----
  {
    float sq[BMAX];
#pragma unroll
    for (int b = 0; b < BMAX; ++b) sq[b] = 0.f;
    for (int c = tid; c < chunks; c += 256) {
#pragma unroll
      for (int b = 0; b < BMAX; ++b) {
        uint4 w = make_uint4(0, 0, 0, 0);
        if (b < p.B) w = *reinterpret_cast<const uint4*>(p.x + (size_t)b * p.ldx + c * 8);
        *reinterpret_cast<uint4*>(xs + (size_t)b * p.K + c * 8) = w;
        if constexpr (MODE != GEMV_RESIDUAL) {
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = bf16_lo(ww[i]), bb = bf16_hi(ww[i]);
            sq[b] += a * a + bb * bb;
          }
        }
      }
    }
    if constexpr (MODE != GEMV_RESIDUAL) {
#pragma unroll
      for (int b = 0; b < BMAX; ++b) {
        const float v = warp_sum(sq[b]);
        if (lane == 0) wred[warp][b] = v;
      }
      __syncthreads();
      if (tid < BMAX) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += wred[w][tid];
        rstd_s[tid] = rsqrtf(t / p.K + p.eps);
      }
    }
    if (tid < BMAX) {
      bestv[tid] = -INFINITY;
      besti[tid] = 0;
    }
    __syncthreads();
  }

  int pos = 0;
  if constexpr (MODE == GEMV_QKV_ROPE) pos = *p.seq_len;

  const int units = (p.N + 7) >> 3;
  int par = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, par ^= 1) {
    const int n0 = u * 8;
    float acc[8][BMAX];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int b = 0; b < BMAX; ++b) acc[r][b] = 0.f;

    const __nv_bfloat16* wbase = p.W + (size_t)n0 * p.K;
    const int rows_ok = min(8, p.N - n0);
#pragma unroll 2
    for (int c = tid; c < chunks; c += 256) {
      uint4 w[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        w[r] = (r < rows_ok) ? ldg_nc_v4(wbase + (size_t)r * p.K + c * 8) : make_uint4(0, 0, 0, 0);
      float xf[BMAX][8];
#pragma unroll
      for (int b = 0; b < BMAX; ++b) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)b * p.K + c * 8);
        xf[b][0] = bf16_lo(xv.x); xf[b][1] = bf16_hi(xv.x); xf[b][2] = bf16_lo(xv.y); xf[b][3] = bf16_hi(xv.y);
        xf[b][4] = bf16_lo(xv.z); xf[b][5] = bf16_hi(xv.z); xf[b][6] = bf16_lo(xv.w); xf[b][7] = bf16_hi(xv.w);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float wf[8] = {bf16_lo(w[r].x), bf16_hi(w[r].x), bf16_lo(w[r].y), bf16_hi(w[r].y),
                             bf16_lo(w[r].z), bf16_hi(w[r].z), bf16_lo(w[r].w), bf16_hi(w[r].w)};
#pragma unroll
        for (int b = 0; b < BMAX; ++b)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r][b] = fmaf(wf[e], xf[b][e], acc[r][b]);
      }
    }
    // ---- reduce over the 256 threads ----
    float* redp = red + par * (8 * 8 * BMAX);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int b = 0; b < BMAX; ++b) {
        const float v = warp_sum(acc[r][b]);
        if (lane == 0) redp[(warp * 8 + r) * BMAX + b] = v;
      }
    __syncthreads();
    float* finp = fin + par * (8 * BMAX);
    if (tid < 8 * BMAX) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += redp[w * 8 * BMAX + tid];   // tid == r*BMAX + b
      finp[tid] = t;
    }
    __syncthreads();

    // ---- fused epilogues ----
    if constexpr (MODE == GEMV_RESIDUAL) {
      if (tid < 8 * BMAX) {
        const int r = tid / BMAX, b = tid % BMAX, n = n0 + r;
        if (b < p.B && n < p.N) {
          const float y = finp[tid] + __bfloat162float(p.res[(size_t)b * p.N + n]);
          p.out[(size_t)b * p.N + n] = __float2bfloat16_rn(y);
        }
      }
    } else if constexpr (MODE == GEMV_SWIGLU) {
      if (tid < 4 * BMAX) {
        const int j = tid / BMAX, b = tid % BMAX, n = n0 + 2 * j;
        if (b < p.B && n + 1 < p.N) {
          const float rs = rstd_s[b];
          // HF rounds gate and up to bf16 before silu*mul (modeling_llama.py:182-184)
          const float g = bf16_round(finp[(2 * j) * BMAX + b] * rs), uu = bf16_round(finp[(2 * j + 1) * BMAX + b] * rs);
          p.out[(size_t)b * (p.N >> 1) + (n >> 1)] = __float2bfloat16_rn(bf16_round(g / (1.f + __expf(-g))) * uu);
        }
      }
    } else if constexpr (MODE == GEMV_QKV_ROPE) {
      if (tid < 4 * BMAX) {
        const int j = tid / BMAX, b = tid % BMAX, n = n0 + 2 * j;
        if (b < p.B && n + 1 < p.N) {
          const float rs = rstd_s[b];
          float x0 = finp[(2 * j) * BMAX + b] * rs, x1 = finp[(2 * j + 1) * BMAX + b] * rs;
          const int which = n / p.H, nh = n - which * p.H, head = nh >> 7, cidx = nh & 127;
          if (which < 2) {
            const float2 cs = p.rope[(size_t)pos * 64 + (cidx >> 1)];
            const float a = x0 * cs.x - x1 * cs.y, c2 = x1 * cs.x + x0 * cs.y;
            x0 = a;
            x1 = c2;
          }
          __nv_bfloat16* dst;
          if (which == 0) dst = p.out + (size_t)b * p.H + nh;
          else dst = ((which == 1) ? p.kcache : p.vcache) + (((size_t)b * p.nH + head) * p.Smax + pos) * 128 + cidx;
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(x0, x1);
        }
      }
    } else {  // GEMV_LOGITS
      if (tid < 8 * BMAX) {
        const int r = tid / BMAX, b = tid % BMAX, n = n0 + r;
        if (b < p.B && n < p.N) {
          const float y = finp[tid] * rstd_s[b];
          finp[tid] = y;
          if (p.logits != nullptr) p.logits[(size_t)b * p.N + n] = y;
        }
      }
      __syncwarp();
      // rows of a unit are finalised by 8*BMAX <= 32 threads of warp 0; one thread per b scans them in index order
      if (tid < BMAX && tid < p.B) {
        for (int r = 0; r < 8 && n0 + r < p.N; ++r) {
          const float y = finp[r * BMAX + tid];
          if (y > bestv[tid]) {
            bestv[tid] = y;
            besti[tid] = n0 + r;
          }
        }
      }
    }
  }

  if constexpr (MODE == GEMV_LOGITS) {
    __syncthreads();
    if (tid < p.B) {
      p.part_val[(size_t)tid * gridDim.x + blockIdx.x] = bestv[tid];
      p.part_idx[(size_t)tid * gridDim.x + blockIdx.x] = besti[tid];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      if (warp < p.B) {   // one warp per batch row
        const int b = warp;
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
          p.next_tokens[b] = bi;
          if (p.out_tokens != nullptr) p.out_tokens[(size_t)b * p.out_stride + *p.step] = bi;
        }
      }
      __syncthreads();
      if (tid == 0) {
        *p.counter = 0;
        if (p.bump) {
          *p.step += 1;
          *p.seq_len_rw += 1;
        }
      }
    }
  }
}

// ============================================================================================
// Decode attention: one new query per (b, head) against the cache (HF:modeling_llama.py:199-222 with S_q = 1).
// grid (B*nH, nsplit), 128 threads.  Split-KV partials are merged by the last CTA of each (b, head).
// ============================================================================================
struct DecAttnParams {
  int B, nH, H, Smax, nsplit;
  const int* seq_len;               // tokens in the cache BEFORE this step; the new K/V were just appended at that index
  const __nv_bfloat16* q;           // [B, H]  (interleaved RoPE order, matches the cache's K)
  const __nv_bfloat16* kcache;      // [B, nH, Smax, 128]
  const __nv_bfloat16* vcache;
  float* part_o;                    // [B*nH, nsplit, 128]
  float2* part_ml;                  // [B*nH, nsplit]
  unsigned int* counters;           // [B*nH]
  __nv_bfloat16* out;               // [B, H]
  float scale_log2e;
  const uint32_t* key_bits;         // [B, mask_words] bit k of row b = key k may be attended (attention_mask); never null
  int mask_words;
};

// HF's 2-D attention_mask as one bit per cache position (vly_kv_set_key_mask); positions nobody masked are 1.
__device__ __forceinline__ bool key_attendable(const uint32_t* __restrict__ bits, int k) {
  return (__ldg(bits + (k >> 5)) >> (k & 31)) & 1u;
}

__global__ void __launch_bounds__(128) decode_attention_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) float dsm[];
  float* sc = dsm;                                  // [per]
  __shared__ float redg[8][128];
  __shared__ float wr[4];
  __shared__ int is_last;
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / p.nH, h = bh % p.nH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int len = *p.seq_len + 1;
  const int per = (len + p.nsplit - 1) / p.nsplit;
  const int k0 = split * per, k1 = min(len, k0 + per);
  const int nk = max(0, k1 - k0);
  const __nv_bfloat16* kb = p.kcache + ((size_t)bh * p.Smax) * 128;
  const __nv_bfloat16* vb = p.vcache + ((size_t)bh * p.Smax) * 128;

  // q chunk of this lane (8 dims)
  const int hl = lane & 15;
  float qf[8];
  {
    const uint4 w = *reinterpret_cast<const uint4*>(p.q + (size_t)b * p.H + h * 128 + hl * 8);
    qf[0] = bf16_lo(w.x); qf[1] = bf16_hi(w.x); qf[2] = bf16_lo(w.y); qf[3] = bf16_hi(w.y);
    qf[4] = bf16_lo(w.z); qf[5] = bf16_hi(w.z); qf[6] = bf16_lo(w.w); qf[7] = bf16_hi(w.w);
  }
  // ---- scores ----
  for (int i0 = warp * 2; i0 < nk; i0 += 8) {   // warp-uniform trip count (two keys per warp iteration)
    const int i = i0 + (lane >> 4);
    const bool ok = i < nk;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (ok) w = ldg_nc_v4(kb + (size_t)(k0 + i) * 128 + hl * 8);
    float d = qf[0] * bf16_lo(w.x) + qf[1] * bf16_hi(w.x) + qf[2] * bf16_lo(w.y) + qf[3] * bf16_hi(w.y) +
              qf[4] * bf16_lo(w.z) + qf[5] * bf16_hi(w.z) + qf[6] * bf16_lo(w.w) + qf[7] * bf16_hi(w.w);
    d += __shfl_xor_sync(0xffffffffu, d, 8);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    if (ok && hl == 0) sc[i] = key_attendable(p.key_bits + (size_t)b * p.mask_words, k0 + i) ? d * p.scale_log2e : -INFINITY;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int i = tid; i < nk; i += 128) m = fmaxf(m, sc[i]);
  m = warp_max(m);
  if (lane == 0) wr[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(wr[0], wr[1]), fmaxf(wr[2], wr[3]));
  __syncthreads();
  float l = 0.f;
  for (int i = tid; i < nk; i += 128) {
    const float e = sc[i] > -INFINITY ? fast_exp2(sc[i] - m) : 0.f;   // masked key (m itself may be -inf)
    sc[i] = e;
    l += e;
  }
  l = warp_sum(l);
  if (lane == 0) wr[warp] = l;
  __syncthreads();
  l = wr[0] + wr[1] + wr[2] + wr[3];
  // ---- P V : 16 threads cover one value row (8 dims each), 8 keys in flight ----
  {
    const int g = tid >> 4, dl = tid & 15;
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = g; i < nk; i += 8) {
      const float pw = sc[i];
      const uint4 w = ldg_nc_v4(vb + (size_t)(k0 + i) * 128 + dl * 8);
      o[0] = fmaf(pw, bf16_lo(w.x), o[0]); o[1] = fmaf(pw, bf16_hi(w.x), o[1]);
      o[2] = fmaf(pw, bf16_lo(w.y), o[2]); o[3] = fmaf(pw, bf16_hi(w.y), o[3]);
      o[4] = fmaf(pw, bf16_lo(w.z), o[4]); o[5] = fmaf(pw, bf16_hi(w.z), o[5]);
      o[6] = fmaf(pw, bf16_lo(w.w), o[6]); o[7] = fmaf(pw, bf16_hi(w.w), o[7]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) redg[g][dl * 8 + e] = o[e];
  }
  __syncthreads();
  float ot = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) ot += redg[g][tid];

  p.part_o[((size_t)bh * p.nsplit + split) * 128 + tid] = ot;
  if (tid == 0) p.part_ml[(size_t)bh * p.nsplit + split] = make_float2(m, l);
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(p.counters + bh, 1u) == (unsigned)p.nsplit - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    float M = -INFINITY;
    for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(&p.part_ml[(size_t)bh * p.nsplit + s].x));
    float L = 0.f, acc = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float ms = __ldcg(&p.part_ml[(size_t)bh * p.nsplit + s].x);
      const float ls = __ldcg(&p.part_ml[(size_t)bh * p.nsplit + s].y);
      if (ls > 0.f) {
        const float w = fast_exp2(ms - M);
        L += ls * w;
        acc += __ldcg(p.part_o + ((size_t)bh * p.nsplit + s) * 128 + tid) * w;
      }
    }
    p.out[(size_t)b * p.H + h * 128 + tid] = __float2bfloat16_rn(L > 0.f ? acc / L : 0.f);
    if (tid == 0) p.counters[bh] = 0;
  }
}

// ============================================================================================
// Shifted cross-entropy (valley_model.py:308-318): row r = (b, s) with s < S-1 scores logits[b, s, :] against labels[b, s+1];
// nll = logsumexp - logit[label]; labels == ignore_index are skipped; loss = mean over the counted rows.
// grid (B*(S-1)), 256 threads -> nll[r] (0 when ignored), cnt[r]; ce_mean_kernel folds them in a fixed order (deterministic).
// ============================================================================================
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int S, int V,
                                                      long long ignore_index, float* __restrict__ nll, int* __restrict__ cnt) {
  __shared__ float red[8];
  __shared__ float bc;
  const int r = blockIdx.x, b = r / (S - 1), s = r % (S - 1);
  const long long lab = labels[(size_t)b * S + s + 1];
  if (lab == ignore_index || lab < 0 || lab >= V) {
    if (threadIdx.x == 0) { nll[r] = 0.f; cnt[r] = 0; }
    return;
  }
  const float* x = logits + ((size_t)b * S + s) * V;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) m = fmaxf(m, x[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bc = t;
  }
  __syncthreads();
  m = bc;
  float sum = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) sum += expf(x[i] - m);
  sum = warp_sum(sum);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    nll[r] = m + logf(t) - x[lab];
    cnt[r] = 1;
  }
}

__global__ void __launch_bounds__(1024) ce_mean_kernel(const float* __restrict__ nll, const int* __restrict__ cnt, int rows, float* __restrict__ loss) {
  __shared__ double sv[32];
  __shared__ int sc[32];
  double a = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < rows; i += 1024) { a += nll[i]; c += cnt[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = a; sc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int n = 0;
    for (int i = 0; i < 32; ++i) { t += sv[i]; n += sc[i]; }
    *loss = n > 0 ? (float)(t / n) : __int_as_float(0x7fc00000);     // no counted label: nan, like torch
  }
}

// Generic dtype conversion to bf16 / fp32 staging (weights upload).
template <typename T>
__global__ void convert_to_bf16_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(float(in[i]));
}
template <typename T>
__global__ void convert_to_f32_bf16rounded_kernel(const T* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = bf16_round(float(in[i]));
}

}  // namespace vly
