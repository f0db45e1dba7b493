/* valley_b200.h -- C ABI of libvalley_b200.so: the B200-native (sm_100a) implementation of Valley's
 * multimodal forward hot path (CLIP ViT-L/14 encode -> temporal pool + mm_projector -> LLaMA decoder
 * with KV cache -> greedy token).
 *
 * The reference (RupertLuo/Valley) has NO FFI / plugin interface: the boundary it offers is the Python
 * nn.Module surface of valley/model/valley_model.py, whose arithmetic is delegated to HuggingFace
 * transformers + ATen.  Each entry point below therefore cites the Python code path it stands in for;
 * valley_b200/model.py re-exposes them under the reference's own class/method names.
 *
 * Conventions
 *   - plain C: opaque handles, raw device pointers + explicit sizes, no C++/torch types.
 *   - every call returns VLY_OK (0) or a negative vly_status; vly_last_error() gives a thread-local message.
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. a torch tensor's data_ptr());
 *     the library owns only its packed weights, workspace and KV caches (tied to vly_ctx / vly_kv).
 *   - calls are stream-ordered on the cudaStream_t passed in (void* stream; NULL = legacy default stream).
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x vly_create fails.
 *   - hot calls do not allocate once the workspace has grown to the largest shapes seen, so the decode
 *     step is captured in a CUDA graph (vly_generate_greedy).
 *   - thread safety: a vly_ctx serialises its workspace-using calls with an internal mutex
 *     (model_worker.py:467-474 may call the model from up to 5 threads); distinct vly_kv are independent.
 */
#ifndef VALLEY_B200_H_
#define VALLEY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vly_ctx vly_ctx;
typedef struct vly_kv vly_kv;

typedef enum {
  VLY_OK = 0,
  VLY_ERR_INVALID = -1,          /* bad argument / shape                                          */
  VLY_ERR_CUDA = -2,             /* CUDA runtime / driver failure                                  */
  VLY_ERR_STATE = -3,            /* call order violated (e.g. weights not finalised)               */
  VLY_ERR_IM_COUNT = -10,        /* "The number of im_start_token and im_end_token should be the same" (valley_model.py:219-220) */
  VLY_ERR_IM_CUT = -11,          /* "Seems that the image is cut." (valley_model.py:226-227)       */
  VLY_ERR_INDEX = -12            /* IndexError the reference would raise reading past the ids row  */
} vly_status;

typedef enum { VLY_F32 = 0, VLY_BF16 = 1, VLY_F16 = 2 } vly_dtype;

/* ValleyConfig(LlamaConfig) keys + the CLIPVisionConfig of config.mm_vision_tower (valley_model.py:18-56). */
typedef struct {
  int32_t hidden_size, num_hidden_layers, num_attention_heads, intermediate_size, vocab_size;
  float rms_norm_eps, rope_theta;
  int32_t max_position_embeddings;   /* KV-cache capacity per sequence (model_max_length 2048, valley_stage2.yaml:54) */
  int32_t vit_hidden, vit_layers, vit_heads, vit_mlp, vit_patch, vit_image;
  float vit_eps;
  int32_t mm_vision_select_layer;    /* config.mm_vision_select_layer (default -1, yaml -2; valley_model.py:173) */
  int32_t device;                    /* CUDA device ordinal */
  int32_t patch_pooling_method;      /* vly_pooling: ValleyLlamaModel.patch_pooling_method (valley_model.py:27, :40-52, :205-213) */
} vly_config;

/* 'mean' (default), 'max', 'temporal_importance' (config.use_patch_importance_pooling: model.pooling_layer.{weight,bias}),
 * 'temporal_transformer' (config.use_delta_transformer: model.transformer_delta_encoder.layers.0.* + model.position_matrix) */
typedef enum { VLY_POOL_MEAN = 0, VLY_POOL_MAX = 1, VLY_POOL_TEMPORAL_IMPORTANCE = 2, VLY_POOL_TEMPORAL_TRANSFORMER = 3 } vly_pooling;

/* Sentinel token ids kept on vision_tower.config by every entry point (run_valley.py:13-18). -1 = unset. */
typedef struct {
  int64_t im_patch_token, im_start_token, im_end_token, vi_frame_token, vi_start_token, vi_end_token;
} vly_tokens;

const char* vly_last_error(void);
const char* vly_version(void);

/* ---- lifetime ---- replaces ValleyLlamaForCausalLM.from_pretrained / __init__ (valley_model.py:24-56, :260-267) */
int vly_create(const vly_config* cfg, vly_ctx** out);
void vly_destroy(vly_ctx* ctx);

/* ---- weights ---- accepts HF state_dict names: model.embed_tokens.weight, model.layers.{i}.*,
 * model.norm.weight, lm_head.weight, model.mm_projector.{weight,bias}, model.vision_tower.vision_model.*
 * Data is copied (converted to bf16) immediately; vly_finalize_weights packs the kernel layouts
 * (fused QKV, LayerNorm/RMSNorm gamma folded into the following Linear, RoPE pair interleave,
 * gate/up interleave) and frees the staging copies. */
int vly_load_weight(vly_ctx* ctx, const char* hf_name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim);
int vly_finalize_weights(vly_ctx* ctx);

/* ---- vision tower ---- vision_tower(images, output_hidden_states=True).hidden_states[select_layer]
 * (valley_model.py:172-184; HF modeling_clip.py:202-219, :363-385, :667-690).
 * pixels [F,3,224,224] of pixel_dtype -> out [F,257,1024] bf16.  Only the layers needed are run. */
int vly_vit_encode(vly_ctx* ctx, const void* pixels_dev, int pixel_dtype, int n_frames, int select_layer, void* out_dev,
                   void* stream);

/* ---- fused ViT encode + all-gather of frame features (multi-GPU; BASELINE north_star "all-gather of frame embeddings before
 * projection").  The reference has no inference collective; torch.distributed/NCCL is the plain path
 * (valley_b200/dist.py); this is the fused one: the LAST ViT layer's fc2+residual epilogue stores every finished tile
 * into the gather buffer of every rank over NVLink (peer-mapped pointers), followed by a flag exchange.
 *   vly_gather_create     : allocate this rank's gather buffer [rows_total, 1024] bf16 (+flags) and export its 64-byte CUDA IPC handle
 *   vly_gather_open_peers : map all ranks' buffers (handles gathered by the caller, e.g. torch.distributed.all_gather_object)
 *   vly_vit_encode_gather : encode n_frames local frames that start at global frame index frame_offset; on return (stream
 *                           order) the local gather buffer holds the features of ALL ranks' frames */
int vly_gather_create(vly_ctx* ctx, int64_t rows_total, void** local_buf_dev, void* ipc_handle_out_64B);
int vly_gather_open_peers(vly_ctx* ctx, const void* handles_64B_each, int world, int rank);
int vly_vit_encode_gather(vly_ctx* ctx, const void* pixels_dev, int pixel_dtype, int n_frames, int frame_offset, int select_layer,
                          void* stream);
/* the same with the local frames dealt round-robin: local frame i is global frame frame_offset + i * frame_stride (frame_stride =
 * world size, frame_offset = rank), so every video's frames are spread over all ranks and every rank needs remote frames */
int vly_vit_encode_gather_strided(vly_ctx* ctx, const void* pixels_dev, int pixel_dtype, int n_frames, int frame_offset,
                                  int frame_stride, int select_layer, void* stream);
/* enqueue after the kernels that read the gather buffer: lets the peers overwrite it in their next vly_vit_encode_gather */
int vly_gather_release(vly_ctx* ctx, void* stream);
/* A peer that never signals makes the device-side wait give up after ~10 s and raise a pinned flag in the context; from then on
 * vly_vit_encode_gather / vly_gather_release / vly_gather_status return VLY_ERR_STATE (the buffer holds stale rows).  Call
 * vly_gather_status after the synchronisation that follows a request (non-blocking read): *timed_out = 0 / 1. */
int vly_gather_status(vly_ctx* ctx, int* timed_out);

/* ---- frame preprocessing: what load_video does to the decoded uint8 frames before the vision tower
 * (valley/util/data_util.py:271-281): Resize(256) [PIL.Image.BILINEAR: video_transform.py:63-66 swaps the names] ->
 * CenterCrop(224) -> /255 -> CLIP mean/std.  Bit-exact with the reference (Pillow's 8-bit two-pass fixed-point convolution).
 *   vly_preprocess_plan   : host, exact: resized size (video_transform.py:56-60, :74-81) and crop origin (:542-543).  No GPU.
 *   vly_resample_coeffs   : host, exact: Pillow's precompute_coeffs + normalize_coeffs_8bpc for the triangle filter.
 *                           kk == NULL queries ksize only; else xmin[out], count[out], kk[out * ksize].  No GPU.
 *   vly_preprocess_frames : frames_dev [T,H,W,3] uint8 (decord's get_batch layout, data_util.py:262) -> out_dev [T,3,224,224]
 *                           of out_dtype (frames first, as every caller permutes it: model_worker.py:337, valley_model.py:430). */
int vly_preprocess_plan(int H, int W, int* new_h, int* new_w, int* crop_y, int* crop_x);
int vly_resample_coeffs(int in_size, int out_size, int* ksize_out, int32_t* xmin, int32_t* count, int32_t* kk);
int vly_preprocess_frames(vly_ctx* ctx, const uint8_t* frames_dev, int T, int H, int W, int out_dtype, void* out_dev, void* stream);

/* ---- mm_projector over every token, == encode_images' projection (valley_model.py:187-190):
 * feats [rows,1024] bf16 -> out [rows,hidden] bf16 */
int vly_project(vly_ctx* ctx, const void* feats_dev, int64_t rows, void* out_dev, void* stream);

/* ---- temporal pool + projector (valley_model.py:205-215 + :190), per cfg.patch_pooling_method:
 * feats [n_videos*T,257,1024] bf16 -> vis_rows [n_videos, 256+T, hidden] bf16
 * (rows 0..255 = pooled patch features in LLM space, rows 256.. = projected per-frame CLS features).
 * mean / temporal_importance pool first and project 256+T rows (the projector is linear and the weights sum to one);
 * max / temporal_transformer project all 257*T rows and pool in LLM space, as the reference does. */
int vly_pool_project(vly_ctx* ctx, const void* feats_dev, int n_videos, int T, void* vis_rows_dev, void* stream);

/* ---- splice plan (pure host integer logic, exact; valley_model.py:196-246).  ids_host [B,S] int64.
 * src_map_host [B,S] int32: -1 = keep the token embedding, j in [0,256) = pooled row j, 256+t = frame t's CLS row.
 * img_idx_host [B] int32: which entry of image_features the sample consumes (-1 = not multimodal).
 * Returns VLY_OK or VLY_ERR_IM_COUNT / VLY_ERR_IM_CUT / VLY_ERR_INDEX exactly where the reference raises.  Needs no GPU. */
int vly_build_splice_map(const int64_t* ids_host, int B, int S, int T, const vly_tokens* tok, int32_t* src_map_host,
                         int32_t* img_idx_host);

/* ---- embed_tokens gather + splice (valley_model.py:160, :223-247): -> inputs_embeds [B,S,hidden] bf16 */
int vly_embed_splice(vly_ctx* ctx, const int64_t* ids_dev, const int32_t* src_map_dev, const int32_t* img_idx_dev,
                     const void* vis_rows_dev, int rows_per_img, int B, int S, void* inputs_embeds_dev, void* stream);

/* ---- KV cache ---- replaces HF DynamicCache / the tuple cache (cache_utils.py:102-120): pre-allocated
 * [L][2][B][heads][max_seq][128] bf16, appended in place by the QKV epilogues. */
int vly_kv_create(vly_ctx* ctx, int batch, int max_seq, vly_kv** out);
void vly_kv_destroy(vly_kv* kv);
int vly_kv_seq_len(vly_kv* kv, int* out_len);   /* host-visible length (syncs the kv's stream state) */
/* which kernel a decode step of this context launches ("decode_step_umma_kernel<4>", "decode_step_kernel<1>", ...): written,
 * NUL-terminated, into name (capacity cap).  For reports (bench.py names the kernel its roofline describes); no GPU work. */
int vly_kv_decode_kernel(vly_kv* kv, char* name, int cap);
int vly_kv_reset(vly_kv* kv, void* stream);
/* HF's 2-D attention_mask (HF masking_utils: the padding mask is AND-ed into the causal mask; build_inputs /
 * tokenizer(padding=True) pad on the LEFT, valley_model.py:249-254 passes the mask through): mask_dev [B, len] uint8 on the
 * device, 0 = cache position k of sequence b must never be attended.  Covers cache positions [0, len) -- set it before the
 * prefill that appends them; later positions (decode) are attendable.  Position ids are NOT shifted by padding (the
 * reference never passes position_ids).  len = 0 clears the mask; vly_kv_reset clears it too. */
int vly_kv_set_key_mask(vly_kv* kv, const uint8_t* mask_dev, int len, void* stream);
/* copy layer l's K (which=0) or V (which=1) as HF-layout [B,heads,len,128] bf16 (de-interleaves K) -- for tests/drop-in */
int vly_kv_export(vly_ctx* ctx, vly_kv* kv, int layer, int which, void* out_dev, void* stream);

/* ---- LlamaModel.forward + lm_head on S new positions (valley_model.py:249-254, :304-305;
 * HF modeling_llama.py:375-426).  inputs_embeds [B,S,hidden] bf16.
 * logits_mode 0: none; 1: last position only -> logits_dev [B,V] fp32; 2: all positions -> [B,S,V] fp32.
 * hidden_out_dev (optional) receives the final-norm'ed hidden states is NOT provided: norm is folded into lm_head. */
int vly_llama_prefill(vly_ctx* ctx, vly_kv* kv, const void* inputs_embeds_dev, int B, int S, int logits_mode,
                      void* logits_dev, int64_t* next_tokens_dev, void* stream);

/* ---- one decode step for B sequences (model_worker.py:380-391): tokens_dev [B] int64 -> next_tokens_dev [B]
 * int64 = argmax (lowest index on ties); logits_dev optional [B,V] fp32. */
int vly_llama_decode(vly_ctx* ctx, vly_kv* kv, const int64_t* tokens_dev, int64_t* next_tokens_dev, void* logits_dev,
                     void* stream);

/* ---- n_steps greedy decode steps with no host round trip (CUDA graph replay): first_tokens_dev [B] is fed at step 0,
 * step i's argmax is fed to step i+1; out_tokens_dev [B, n_steps] int64 receives every argmax. */
int vly_generate_greedy(vly_ctx* ctx, vly_kv* kv, const int64_t* first_tokens_dev, int n_steps, int64_t* out_tokens_dev,
                        void* stream);

/* ---- loss of a forward with labels (valley_model.py:308-318): shift by one, CrossEntropyLoss() = mean of
 * logsumexp(logits[b,s,:]) - logits[b,s,labels[b,s+1]] over the labels != ignore_index (-100).  logits_dev [B,S,V] fp32
 * (vly_llama_prefill logits_mode 2), labels_dev [B,S] int64, loss_out_dev one fp32 (nan when no label counts, like torch). */
int vly_cross_entropy(vly_ctx* ctx, const float* logits_dev, const int64_t* labels_dev, int B, int S, int64_t ignore_index,
                      float* loss_out_dev, void* stream);

/* ---- token selection on the device (model_worker.py:388-397; HF generate as called at valley_model.py:432) ----
 * temperature < 1e-4: arg-max (model_worker.py:390-391); otherwise multinomial(softmax(logits / temperature)) (:392-395), drawn
 * with the Gumbel-max identity from counter-based Philox noise keyed by `seed` -- fused into the arg-max epilogue of the decode
 * step, so sampling costs no extra pass and no host round trip.  eos_token_id >= 0: a row that emits it is finished (:396-397);
 * finished rows emit pad_token_id (HF generate) and once every row has finished the remaining steps are skipped. */
typedef struct {
  float temperature;
  uint64_t seed;
  int64_t eos_token_id;      /* -1: none */
  int64_t pad_token_id;
  int64_t stop_token_id;     /* -1: none; a second id that ends a row: the worker's single-token stop string (model_worker.py:355-360, :396-397) */
} vly_sampling;

/* the first generated token: select from the prefill's last-position logits [B,V] fp32 (vly_llama_prefill logits_mode 1);
 * starts a generation (clears the finished flags). */
int vly_sample_logits(vly_ctx* ctx, vly_kv* kv, const float* logits_dev, const vly_sampling* sampling, int64_t* tokens_out_dev,
                      void* stream);
/* vly_generate_greedy with token selection per `sampling` (NULL = greedy, no stop token).  steps_done_dev (optional, device
 * int32) receives the number of steps actually executed: n_steps, or fewer when every row reached eos -- columns
 * [0, steps_done) of out_tokens_dev are valid. */
int vly_generate(vly_ctx* ctx, vly_kv* kv, const int64_t* first_tokens_dev, int n_steps, int64_t* out_tokens_dev,
                 const vly_sampling* sampling, int* steps_done_dev, void* stream);

/* ---- introspection for bench / tests ---- */
int vly_kernel_launch_count(vly_ctx* ctx, int64_t* out);   /* kernels launched by this ctx so far */
int vly_num_sms(vly_ctx* ctx, int* out);

/* in-kernel cycle counters of the last decode step / ViT attention launch, filled only when the process runs with VLY_MEGA_DBG=1 /
 * VLY_ATTN_DBG=1 (tools/bench_decode.py, tools/bench_vit.py): copies n int64 values to host_out; returns 0, -1 (never enabled) or -2. */
int vly_debug_mega_counters(long long* host_out, int n);
int vly_debug_attn_counters(long long* host_out, int n);
/* internal: lets the host-only translation units (host_splice.cpp, host_preprocess.cpp) set the thread-local error message */
void vly_set_error_(const char* message);

/* ---- low-level op hooks (per-kernel parity tests) ----
 * D[M,N] = epilogue(A[M,K] * W[N,K]^T): epi 0 bias->bf16, 3 bias+residual (in-place allowed)->bf16.  bias_dev fp32 or NULL. */
int vly_test_gemm(vly_ctx* ctx, const void* a_dev, const void* w_dev, int M, int N, int K, int epi, const float* bias_dev,
                  const void* residual_dev, void* out_dev, int block_n, void* stream);
/* ViT attention on a packed qkv [F*257, 3072] bf16 -> ctx [F*257,1024] bf16 */
int vly_test_vit_attention(vly_ctx* ctx, const void* qkv_dev, int n_frames, void* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VALLEY_B200_H_ */
