// Host-side, exact integer logic of Valley's visual-token splice ("prepare_inputs_labels_for_multimodal").
// Restates valley/model/valley_model.py:196-246 as a per-position source map; no GPU involved.
//
//   * a sample with no <im_patch> token is not multimodal: untouched, does NOT consume an image (:198-202, :246)
//   * count(<im_start>) != count(<im_end>)                         -> ValueError (:219-220)
//   * for every <im_start> at p (ascending): ids[p+257] must be <im_end> else ValueError("Seems that the image
//     is cut.") (:226-227); reading past the row is the IndexError torch raises; [p+1, p+257) <- pooled rows (:228)
//   * the video block (:231-244) sits in a bare try/except: unset vi_* ids, unbalanced <vi_start>/<vi_end>,
//     count(<vi_frame>) != T, a misplaced <vi_end> or an out-of-range index ALL fall back to the image-only result
#include <stdint.h>
#include <vector>

#include "../../include/valley_b200.h"

extern "C" int vly_build_splice_map(const int64_t* ids, int B, int S, int T, const vly_tokens* tok, int32_t* src_map,
                                     int32_t* img_idx) {
  extern void vly_set_error_(const char*);
  if (B == 0) return VLY_OK;   // empty batch: nothing to plan
  if (tok == nullptr || img_idx == nullptr || B < 0 || S < 0 || T < 0 || (S > 0 && (ids == nullptr || src_map == nullptr))) {
    vly_set_error_("vly_build_splice_map: null argument or negative size");
    return VLY_ERR_INVALID;
  }
  const int NP = 256;  // num_patches: hard-coded 256 in the reference (valley_model.py:192, :387)
  int cur = 0;
  std::vector<int32_t> vid(S);
  for (int b = 0; b < B; ++b) {
    const int64_t* row = ids + (size_t)b * S;
    int32_t* map = src_map + (size_t)b * S;
    for (int s = 0; s < S; ++s) map[s] = -1;
    long n_patch = 0, n_start = 0, n_end = 0;
    for (int s = 0; s < S; ++s) {
      n_patch += row[s] == tok->im_patch_token;
      n_start += row[s] == tok->im_start_token;
      n_end += row[s] == tok->im_end_token;
    }
    if (n_patch == 0) {
      img_idx[b] = -1;
      continue;
    }
    img_idx[b] = cur++;
    if (n_start != n_end) {
      vly_set_error_("The number of im_start_token and im_end_token should be the same");
      return VLY_ERR_IM_COUNT;
    }
    for (int p = 0; p < S; ++p) {
      if (row[p] != tok->im_start_token) continue;
      if (p + NP + 1 >= S) {
        vly_set_error_("index out of range reading the token after the image block");
        return VLY_ERR_INDEX;
      }
      if (row[p + NP + 1] != tok->im_end_token) {
        vly_set_error_("Seems that the image is cut.");
        return VLY_ERR_IM_CUT;
      }
      for (int j = 0; j < NP; ++j) map[p + 1 + j] = j;
    }
    // ---- video frames: any failure -> keep the image-only map ----
    bool ok = tok->vi_start_token >= 0 && tok->vi_end_token >= 0 && tok->vi_frame_token >= 0;
    if (ok) {
      long n_vs = 0, n_ve = 0, n_vf = 0;
      for (int s = 0; s < S; ++s) {
        n_vs += row[s] == tok->vi_start_token;
        n_ve += row[s] == tok->vi_end_token;
        n_vf += row[s] == tok->vi_frame_token;
      }
      ok = (n_vs == n_ve) && (n_vf == T);
    }
    if (ok) {
      for (int s = 0; s < S; ++s) vid[s] = map[s];
      for (int q = 0; q < S && ok; ++q) {
        if (row[q] != tok->vi_start_token) continue;
        if (q + T + 1 >= S || row[q + T + 1] != tok->vi_end_token) {
          ok = false;
          break;
        }
        for (int t = 0; t < T; ++t) vid[q + 1 + t] = NP + t;
      }
      if (ok)
        for (int s = 0; s < S; ++s) map[s] = vid[s];
    }
  }
  return VLY_OK;
}
