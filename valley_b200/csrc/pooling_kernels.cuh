// Temporal pooling variants of the visual tokens (valley/model/valley_model.py:205-213).
//   mean                 : simt_kernels.cuh temporal_pool_kernel (pool first, then project -- the projector is linear)
//   temporal_importance  : softmax_t(w . flatten(proj(x_t)) + b) weighted sum (:113-121).  The weights sum to one, so the
//                          weighted sum ALSO commutes with the projector, and the score needs no projected features either:
//                          w . flatten(W x + b_proj) = sum_p (W^T w_p) . x_{t,p} + const, the constant cancels in the softmax.
//                          U = W^T w ([256,1024] fp32) is folded once at weight-finalise time.
//   max                  : element-wise max over frames of the PROJECTED patch rows (:208-209) -- does not commute
//   temporal_transformer : one post-LN TransformerEncoderLayer over the T frames of every patch position, last frame's
//                          output + temporal mean (:123-133).  GEMMs run on gemm_tc_kernel; the pieces here are the glue:
//                          position add, the (tiny) T-key attention of the last query, LayerNorm, relu, the final add.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"

namespace vly {

// U[p][d] = sum_h wpool[p*H + h] * proj_w[h*D + d]     grid (256), block D/4 threads... generic: thread per d, loop h
__global__ void fold_importance_kernel(const __nv_bfloat16* __restrict__ wpool, const __nv_bfloat16* __restrict__ proj_w,
                                       float* __restrict__ U, int H, int D) {
  const int p = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc = fmaf(__bfloat162float(wpool[(size_t)p * H + h]), __bfloat162float(proj_w[(size_t)h * D + d]), acc);
    U[(size_t)p * D + d] = acc;
  }
}

// score[v*T + t] = sum_{p,d} U[p][d] * feats[v,t,1+p,d]      grid (T, NV), 256 threads
__global__ void __launch_bounds__(256) importance_score_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ U,
                                                               float* __restrict__ score, int T, int tokens, int D) {
  __shared__ float red[8];
  const int t = blockIdx.x, v = blockIdx.y;
  const __nv_bfloat16* x = feats + (((size_t)v * T + t) * tokens + 1) * D;
  const int chunks = (tokens - 1) * D / 8;
  float acc = 0.f;
  for (int i = threadIdx.x; i < chunks; i += 256) {
    const uint4 w = *reinterpret_cast<const uint4*>(x + (size_t)i * 8);
    const float4 u0 = *reinterpret_cast<const float4*>(U + (size_t)i * 8), u1 = *reinterpret_cast<const float4*>(U + (size_t)i * 8 + 4);
    acc += bf16_lo(w.x) * u0.x + bf16_hi(w.x) * u0.y + bf16_lo(w.y) * u0.z + bf16_hi(w.y) * u0.w +
           bf16_lo(w.z) * u1.x + bf16_hi(w.z) * u1.y + bf16_lo(w.w) * u1.z + bf16_hi(w.w) * u1.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    score[(size_t)v * T + t] = s;
  }
}

// feats [NV*T, tokens, D] -> out [NV, tokens-1+T, D]: rows < tokens-1 = sum_t softmax_t(score)[t] * patch row, rest = CLS rows
__global__ void weighted_pool_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ score,
                                     __nv_bfloat16* __restrict__ out, int NV, int T, int tokens, int D) {
  const int rows_out = tokens - 1 + T, chunks = D / 8;
  const long long total = (long long)NV * rows_out * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks);
    const long long ro = i / chunks;
    const int r = int(ro % rows_out), v = int(ro / rows_out);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < tokens - 1) {
      const float* sc = score + (size_t)v * T;
      float mx = -INFINITY, den = 0.f;
      for (int t = 0; t < T; ++t) mx = fmaxf(mx, sc[t]);
      for (int t = 0; t < T; ++t) den += __expf(sc[t] - mx);
      const float inv = 1.f / den;
      for (int t = 0; t < T; ++t) {
        const float wt = __expf(sc[t] - mx) * inv;
        const uint4 w = *reinterpret_cast<const uint4*>(feats + (((size_t)v * T + t) * tokens + (r + 1)) * D + ch * 8);
        acc[0] = fmaf(wt, bf16_lo(w.x), acc[0]); acc[1] = fmaf(wt, bf16_hi(w.x), acc[1]);
        acc[2] = fmaf(wt, bf16_lo(w.y), acc[2]); acc[3] = fmaf(wt, bf16_hi(w.y), acc[3]);
        acc[4] = fmaf(wt, bf16_lo(w.z), acc[4]); acc[5] = fmaf(wt, bf16_hi(w.z), acc[5]);
        acc[6] = fmaf(wt, bf16_lo(w.w), acc[6]); acc[7] = fmaf(wt, bf16_hi(w.w), acc[7]);
      }
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

// projected P [T, tokens, H] of ONE video -> vis [tokens-1+T, H]: rows < tokens-1 = max_t P[t, 1+r], rest = P[t, 0]
__global__ void temporal_max_kernel(const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ vis, int T, int tokens, int H) {
  const int rows_out = tokens - 1 + T, chunks = H / 8;
  const long long total = (long long)rows_out * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks), r = int(i / chunks);
    float a[8];
    if (r < tokens - 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = -INFINITY;
      for (int t = 0; t < T; ++t) {
        const uint4 w = *reinterpret_cast<const uint4*>(P + ((size_t)t * tokens + r + 1) * H + ch * 8);
        a[0] = fmaxf(a[0], bf16_lo(w.x)); a[1] = fmaxf(a[1], bf16_hi(w.x)); a[2] = fmaxf(a[2], bf16_lo(w.y)); a[3] = fmaxf(a[3], bf16_hi(w.y));
        a[4] = fmaxf(a[4], bf16_lo(w.z)); a[5] = fmaxf(a[5], bf16_hi(w.z)); a[6] = fmaxf(a[6], bf16_lo(w.w)); a[7] = fmaxf(a[7], bf16_hi(w.w));
      }
    } else {
      const uint4 w = *reinterpret_cast<const uint4*>(P + ((size_t)(r - (tokens - 1)) * tokens) * H + ch * 8);
      a[0] = bf16_lo(w.x); a[1] = bf16_hi(w.x); a[2] = bf16_lo(w.y); a[3] = bf16_hi(w.y);
      a[4] = bf16_lo(w.z); a[5] = bf16_hi(w.z); a[6] = bf16_lo(w.w); a[7] = bf16_hi(w.w);
    }
    *reinterpret_cast<uint4*>(vis + (size_t)r * H + ch * 8) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]),
                                                                         pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
  }
}

// ---- temporal transformer glue (one video) ----
// Xp[t*NP + p, :] = bf16(P[t*tokens + 1 + p, :] + pos[t, :])         (valley_model.py:127-129; frame-major rows so that the
// last frame's NP rows -- the only queries needed -- are one contiguous block)
__global__ void delta_add_pos_kernel(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ pos,
                                     __nv_bfloat16* __restrict__ Xp, int T, int tokens, int H) {
  const int NP = tokens - 1, chunks = H / 8;
  const long long total = (long long)T * NP * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks);
    const long long row = i / chunks;
    const int p = int(row % NP), t = int(row / NP);
    const uint4 a = *reinterpret_cast<const uint4*>(P + ((size_t)t * tokens + 1 + p) * H + ch * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(pos + (size_t)t * H + ch * 8);
    *reinterpret_cast<uint4*>(Xp + row * H + ch * 8) =
        make_uint4(pack_bf16x2(bf16_lo(a.x) + bf16_lo(b.x), bf16_hi(a.x) + bf16_hi(b.x)), pack_bf16x2(bf16_lo(a.y) + bf16_lo(b.y), bf16_hi(a.y) + bf16_hi(b.y)),
                   pack_bf16x2(bf16_lo(a.z) + bf16_lo(b.z), bf16_hi(a.z) + bf16_hi(b.z)), pack_bf16x2(bf16_lo(a.w) + bf16_lo(b.w), bf16_hi(a.w) + bf16_hi(b.w)));
  }
}

// attention of the LAST frame's query over the T frames, per patch position and head (torch MHA: q scaled by hd^-0.5).
// q [NP, H]; kv [T*NP, 2H] (k | v); out [NP, H].  grid (NP), block nhead warps (one head each).  T <= 64.
__global__ void delta_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                       __nv_bfloat16* __restrict__ out, int T, int NP, int H, int nhead) {
  const int p = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (head >= nhead) return;
  const int hd = H / nhead;
  const float scale = rsqrtf((float)hd);
  const __nv_bfloat16* qh = q + (size_t)p * H + head * hd;
  float sc0 = -INFINITY, sc1 = -INFINITY;            // lane holds the scores of keys t = lane and lane + 32
  for (int t = 0; t < T; ++t) {
    const __nv_bfloat16* kh = kv + ((size_t)t * NP + p) * 2 * H + head * hd;
    float d = 0.f;
    for (int e = lane * 2; e < hd; e += 64) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(qh + e), b = *reinterpret_cast<const uint32_t*>(kh + e);
      d += bf16_lo(a) * bf16_lo(b) + bf16_hi(a) * bf16_hi(b);
    }
    d = warp_sum(d) * scale;
    if (t == lane) sc0 = d;
    if (t == lane + 32) sc1 = d;
  }
  const float mx = warp_max(fmaxf(sc0, sc1));
  const float e0 = sc0 > -INFINITY ? __expf(sc0 - mx) : 0.f, e1 = sc1 > -INFINITY ? __expf(sc1 - mx) : 0.f;
  const float inv = 1.f / warp_sum(e0 + e1);
  for (int e = lane * 2; e < hd; e += 64) {
    float o0 = 0.f, o1 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float w = __shfl_sync(0xffffffffu, t < 32 ? e0 : e1, t & 31) * inv;
      const uint32_t b = *reinterpret_cast<const uint32_t*>(kv + ((size_t)t * NP + p) * 2 * H + H + head * hd + e);
      o0 = fmaf(w, bf16_lo(b), o0);
      o1 = fmaf(w, bf16_hi(b), o1);
    }
    *reinterpret_cast<uint32_t*>(out + (size_t)p * H + head * hd + e) = pack_bf16x2(o0, o1);
  }
}

// y = LayerNorm(x) * g + b over rows of length H (fp32 statistics, two passes over registers/global), in place allowed
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ g,
                                                             const float* __restrict__ b, __nv_bfloat16* __restrict__ y, int H,
                                                             float eps) {
  __shared__ float red[8], stat[2];
  const __nv_bfloat16* xr = x + (size_t)blockIdx.x * H;
  float s = 0.f;
  for (int i = threadIdx.x; i < H; i += 256) s += __bfloat162float(xr[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat[0] = t / H;
  }
  __syncthreads();
  const float mean = stat[0];
  float v = 0.f;
  for (int i = threadIdx.x; i < H; i += 256) {
    const float d = __bfloat162float(xr[i]) - mean;
    v += d * d;
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat[1] = rsqrtf(t / H + eps);
  }
  __syncthreads();
  const float rstd = stat[1];
  for (int i = threadIdx.x; i < H; i += 256)
    y[(size_t)blockIdx.x * H + i] = __float2bfloat16_rn((__bfloat162float(xr[i]) - mean) * rstd * g[i] + b[i]);
}

__global__ void relu_inplace_kernel(__nv_bfloat16* __restrict__ x, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 w = *reinterpret_cast<uint4*>(x + i * 8);
    const uint32_t in[4] = {w.x, w.y, w.z, w.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(fmaxf(bf16_lo(in[j]), 0.f), fmaxf(bf16_hi(in[j]), 0.f));
    *reinterpret_cast<uint4*>(x + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// vis [NP + T, H]: rows < NP = delta[p] + mean_t P[t, 1+p]  (:130-132), rows NP + t = P[t, 0] (frame CLS rows, :215)
__global__ void delta_finish_kernel(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ delta,
                                    __nv_bfloat16* __restrict__ vis, int T, int tokens, int H) {
  const int NP = tokens - 1, rows_out = NP + T, chunks = H / 8;
  const long long total = (long long)rows_out * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i % chunks), r = int(i / chunks);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < NP) {
      for (int t = 0; t < T; ++t) {
        const uint4 w = *reinterpret_cast<const uint4*>(P + ((size_t)t * tokens + 1 + r) * H + ch * 8);
        a[0] += bf16_lo(w.x); a[1] += bf16_hi(w.x); a[2] += bf16_lo(w.y); a[3] += bf16_hi(w.y);
        a[4] += bf16_lo(w.z); a[5] += bf16_hi(w.z); a[6] += bf16_lo(w.w); a[7] += bf16_hi(w.w);
      }
      const float inv = 1.f / T;
      const uint4 d = *reinterpret_cast<const uint4*>(delta + (size_t)r * H + ch * 8);
      a[0] = a[0] * inv + bf16_lo(d.x); a[1] = a[1] * inv + bf16_hi(d.x); a[2] = a[2] * inv + bf16_lo(d.y); a[3] = a[3] * inv + bf16_hi(d.y);
      a[4] = a[4] * inv + bf16_lo(d.z); a[5] = a[5] * inv + bf16_hi(d.z); a[6] = a[6] * inv + bf16_lo(d.w); a[7] = a[7] * inv + bf16_hi(d.w);
    } else {
      const uint4 w = *reinterpret_cast<const uint4*>(P + ((size_t)(r - NP) * tokens) * H + ch * 8);
      a[0] = bf16_lo(w.x); a[1] = bf16_hi(w.x); a[2] = bf16_lo(w.y); a[3] = bf16_hi(w.y);
      a[4] = bf16_lo(w.z); a[5] = bf16_hi(w.z); a[6] = bf16_lo(w.w); a[7] = bf16_hi(w.w);
    }
    *reinterpret_cast<uint4*>(vis + (size_t)r * H + ch * 8) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]),
                                                                         pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
  }
}

}  // namespace vly
