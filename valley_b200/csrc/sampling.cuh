// On-device token selection for the decode loop.
//
// Reference semantics (valley/serve/model_worker.py:388-397, and HF generate as called at valley_model.py:432):
//   temperature < 1e-4 : token = argmax(last_token_logits)                         (:390-391)
//   otherwise          : probs = softmax(last_token_logits / temperature); token = multinomial(probs, 1)   (:392-395)
//   token == eos       : the sequence stops (:396-397); HF pads finished rows of a batch with pad_token_id and stops
//                        when every row has finished.
//
// multinomial(softmax(z)) is drawn with the Gumbel-max identity: argmax_n(z_n + G_n), G_n i.i.d. standard Gumbel, has exactly
// that distribution -- so sampling is the SAME fused arg-max epilogue the greedy path already runs, with a counter-based noise
// term and no softmax pass, no prefix sum and no host round trip.  The noise is Philox4x32-10 keyed by the request seed with the
// counter (vocabulary index, batch row, position of the query token): every path (the persistent decode kernel's epilogue, the
// stand-alone kernel below) computes bit-identical scores for the same logits, which is what the tests check, next to a
// goodness-of-fit test of the drawn distribution against softmax(logits / T).  (The stream differs from torch's Philox offsets,
// so ids are not comparable draw-by-draw with torch.multinomial -- no sampler on different hardware is.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vly {

constexpr int kMaxSampleRows = 64;

struct SampleState {          // lives in device memory next to the KV cache; read by every decode step
  float inv_temp;             // 1 / temperature
  int enabled;                // 0 = greedy (scores are the raw logits: bit-identical to the plain arg-max)
  uint32_t seed_lo, seed_hi;
  long long eos, pad;         // eos < 0: no stop token
  long long stop2;            // second stop id (the worker's single-token stop string, model_worker.py:355-360); < 0: none
  int all_done;               // every row has produced eos: further steps exit at once
  int steps_valid;            // decode steps executed before all_done was raised (the one that raised it included)
  int done[kMaxSampleRows];
};

__device__ __forceinline__ uint32_t philox4x32_10_first(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

// score whose arg-max over n is the selected token
__device__ __forceinline__ float sample_score(float logit, float inv_temp, uint32_t seed_lo, uint32_t seed_hi, int n, int b, int pos) {
  const uint32_t x = philox4x32_10_first((uint32_t)n, (uint32_t)b, (uint32_t)pos, 0x56414c59u, seed_lo, seed_hi);
  const float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1), 24 bits
  return fmaf(logit, inv_temp, -logf(-logf(u)));
}

// eos / pad bookkeeping for one row's freshly selected token; returns the token to emit
__device__ __forceinline__ long long sample_finish_row(SampleState* s, int b, long long tok) {
  if (s->eos < 0 && s->stop2 < 0) return tok;
  if (s->done[b]) return s->pad;
  if (tok == s->eos || tok == s->stop2) s->done[b] = 1;     // (ids are >= 0, an unset -1 never matches)
  return tok;
}

// Stand-alone selection over a [B, V] fp32 logits block: the first token after a prefill (reset = 1), and the post-step
// selection of the per-op decode paths (B > 4).  One CTA; rows are handled one after the other.
//   pos = *seq_len - 1 : position of the query token that produced these logits
__global__ void __launch_bounds__(1024) sample_rows_kernel(const float* __restrict__ logits, int B, int V, SampleState* s,
                                                           const int* seq_len, const int* step, long long* next_tokens,
                                                           long long* out_tokens, int out_stride, int reset) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (reset) {
    if (tid < kMaxSampleRows) s->done[tid] = 0;
    if (tid == 0) { s->all_done = 0; s->steps_valid = 0; }
    __syncthreads();
  } else {
    if (!s->enabled && s->eos < 0 && s->stop2 < 0) return;   // plain greedy: the arg-max epilogue already wrote the token
    if (s->all_done) {                              // every row finished earlier: keep emitting pad
      if (tid < B) {
        next_tokens[tid] = s->pad;
        if (out_tokens) out_tokens[(size_t)tid * out_stride + (*step - 1)] = s->pad;
      }
      return;
    }
  }
  const int pos = *seq_len - 1;
  const bool on = s->enabled != 0;
  const float it = s->inv_temp;
  const uint32_t k0 = s->seed_lo, k1 = s->seed_hi;
  for (int b = 0; b < B; ++b) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = tid; n < V; n += 1024) {
      const float y = logits[(size_t)b * V + n];
      const float v = on ? sample_score(y, it, k0, k1, n, b, pos) : y;
      if (v > bv) { bv = v; bi = n; }               // ascending n per thread: first maximum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = sv[lane];
      bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        const long long tok = sample_finish_row(s, b, bi);
        next_tokens[b] = tok;
        if (out_tokens) out_tokens[(size_t)b * out_stride + (*step - 1)] = tok;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (!reset) s->steps_valid += 1;
    if (s->eos >= 0 || s->stop2 >= 0) {
      int all = 1;
      for (int b = 0; b < B; ++b) all &= s->done[b];
      s->all_done = all;
    }
  }
}

}  // namespace vly
