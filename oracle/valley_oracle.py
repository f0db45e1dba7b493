"""CPU ORACLE for the Valley multimodal forward hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, eager, functional) restatement of the reference's
algorithm for the path named by BASELINE.json's north_star.  It is the checker the CUDA
path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import it; nothing under ``valley_b200/``
does.  It is never the thing shipped or measured as the product.

Pinned how: the reference has no tests, fixtures or golden vectors for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against the reference ITSELF run in the
build container: ``oracle/make_golden.py`` imports ``/root/reference/valley/model/
valley_model.py`` (with decord/skimage stubbed) on top of the installed HuggingFace
transformers 5.5.0, checks this restatement against it on identical random-init weights
(fp32: bit-exact or <=1e-5 rel), and writes the reference's outputs to ``tests/golden/``.
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures on every run.

The arithmetic of the path lives in a third-party dependency that is not vendored in
/root/reference: ``transformers`` pinned at git cae78c46 (pyproject.toml:19).  File:line
citations of the form ``HF:`` refer to the installed transformers 5.5.0, whose fp32
results agree with the pin's to ~3e-7 (SURVEY.md 8c caveat 3).  Citations without a
prefix are relative to /root/reference.

Weights are passed as a flat dict with HuggingFace state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# CLIP ViT-L/14 vision tower
# ----------------------------------------------------------------------------------------------
def quick_gelu(x: Tensor) -> Tensor:
    """HF:activations.py:117-123 -- x * sigmoid(1.702 x)."""
    return x * torch.sigmoid(1.702 * x)


def vit_embeddings(w: Dict[str, Tensor], pixels: Tensor, pfx: str, patch: int) -> Tensor:
    """HF:models/clip/modeling_clip.py:202-219 (CLIPVisionEmbeddings.forward).

    conv2d(stride=kernel=patch, no bias) -> flatten -> transpose -> cat CLS -> + pos-emb.
    Pixels are cast to the weight dtype first (:208-209).
    """
    pw = w[pfx + "embeddings.patch_embedding.weight"]
    x = F.conv2d(pixels.to(pw.dtype), pw, bias=None, stride=patch)
    x = x.flatten(2).transpose(1, 2)                                     # [F, 256, D]
    cls = w[pfx + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)                                       # [F, 257, D]
    return x + w[pfx + "embeddings.position_embedding.weight"][None, : x.shape[1]]


def vit_attention(w: Dict[str, Tensor], x: Tensor, pfx: str, heads: int) -> Tensor:
    """HF:modeling_clip.py:300-336 (CLIPAttention.forward) with the eager kernel :261-279:
    scores = q k^T * scale; softmax in fp32, cast back; @ v; out_proj."""
    Fr, N, D = x.shape
    hd = D // heads
    q = F.linear(x, w[pfx + "q_proj.weight"], w[pfx + "q_proj.bias"]).view(Fr, N, heads, hd).transpose(1, 2)
    k = F.linear(x, w[pfx + "k_proj.weight"], w[pfx + "k_proj.bias"]).view(Fr, N, heads, hd).transpose(1, 2)
    v = F.linear(x, w[pfx + "v_proj.weight"], w[pfx + "v_proj.bias"]).view(Fr, N, heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
    p = F.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(p, v).transpose(1, 2).reshape(Fr, N, D)
    return F.linear(o, w[pfx + "out_proj.weight"], w[pfx + "out_proj.bias"])


def vit_layer(w: Dict[str, Tensor], x: Tensor, pfx: str, heads: int, eps: float) -> Tensor:
    """HF:modeling_clip.py:363-385 (CLIPEncoderLayer.forward): pre-LN MHA + pre-LN MLP."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), w[pfx + "layer_norm1.weight"], w[pfx + "layer_norm1.bias"], eps)
    x = x + vit_attention(w, h, pfx + "self_attn.", heads)
    h = F.layer_norm(x, (D,), w[pfx + "layer_norm2.weight"], w[pfx + "layer_norm2.bias"], eps)
    h = F.linear(h, w[pfx + "mlp.fc1.weight"], w[pfx + "mlp.fc1.bias"])
    h = quick_gelu(h)                                                     # HF:modeling_clip.py:347-351
    h = F.linear(h, w[pfx + "mlp.fc2.weight"], w[pfx + "mlp.fc2.bias"])
    return x + h


def vit_hidden_state(w: Dict[str, Tensor], pixels: Tensor, select_layer: int, *, num_layers: int,
                     heads: int = 16, patch: int = 14, eps: float = 1e-5,
                     pfx: str = "model.vision_tower.vision_model.") -> Tensor:
    """``vision_tower(images, output_hidden_states=True).hidden_states[select_layer]``
    (valley_model.py:172-184; HF:modeling_clip.py:667-690, :462-507).

    hidden_states[0] is the pre-LayerNorm'd embedding, hidden_states[k] the output of encoder
    layer k; no post_layernorm is applied to it.  Only the layers that
    are needed are evaluated (the reference evaluates all of them and discards the rest).
    """
    idx = select_layer if select_layer >= 0 else num_layers + 1 + select_layer
    assert 0 <= idx <= num_layers
    x = vit_embeddings(w, pixels, pfx, patch)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), w[pfx + "pre_layrnorm.weight"], w[pfx + "pre_layrnorm.bias"], eps)
    for i in range(idx):
        x = vit_layer(w, x, f"{pfx}encoder.layers.{i}.", heads, eps)
    return x


def encode_images(w: Dict[str, Tensor], images, select_layer: int, *, num_layers: int, heads: int = 16,
                  patch: int = 14, eps: float = 1e-5):
    """valley_model.py:163-190: per batch item run the tower, pick hidden_states[select_layer]
    (all 257 tokens), stack, then mm_projector.  ``images`` is a [B,T,3,H,W] tensor or a list
    of [T_i,3,H,W] tensors (kept as a list, :168-176, :187-188)."""
    def one(img):
        return vit_hidden_state(w, img, select_layer, num_layers=num_layers, heads=heads, patch=patch, eps=eps)

    pw, pb = w["model.mm_projector.weight"], w["model.mm_projector.bias"]
    if isinstance(images, (list, tuple)):
        return [F.linear(one(img), pw, pb) for img in images]
    feats = torch.stack([one(images[b]) for b in range(len(images))])     # [B,T,257,1024]
    return F.linear(feats, pw, pb)                                        # [B,T,257,H]


# ----------------------------------------------------------------------------------------------
# temporal pool + splice  ("prepare_inputs_labels_for_multimodal")
# ----------------------------------------------------------------------------------------------
class SentinelIds:
    """The token ids the reference keeps on ``vision_tower.config`` (run_valley.py:13-18)."""

    def __init__(self, im_patch, im_start, im_end, vi_frame=None, vi_start=None, vi_end=None):
        self.im_patch_token, self.im_start_token, self.im_end_token = im_patch, im_start, im_end
        self.vi_frame_token, self.vi_start_token, self.vi_end_token = vi_frame, vi_start, vi_end


def text_importance_pooling(w: Dict[str, Tensor], patch: Tensor) -> Tensor:
    """valley_model.py:113-121 ("temporal_importance", config.use_patch_importance_pooling, :40-43):
    Linear(256*H -> 1) on each frame's flattened patch block, softmax over the T frames, weighted sum.  patch [T,256,H]."""
    flat = torch.flatten(patch, start_dim=1)
    score = F.softmax(F.linear(flat, w["model.pooling_layer.weight"], w["model.pooling_layer.bias"]), dim=0)   # [T,1]
    return torch.sum(score.unsqueeze(2) * patch, dim=0)


def temporal_transformer_delta_adding(w: Dict[str, Tensor], patch: Tensor, nhead: int = 8, eps: float = 1e-5) -> Tensor:
    """valley_model.py:123-133 ("temporal_transformer", config.use_delta_transformer, :45-52): every patch position is a
    sequence over the T frames; + position_matrix[:T]; ONE post-LN nn.TransformerEncoderLayer(d_model=H, nhead=8,
    dim_feedforward=2048, relu, batch_first) in eval mode (dropout off); take the LAST frame's output, add the temporal mean.
    The layer is restated from torch/nn/modules/transformer.py (TransformerEncoderLayer.forward, norm_first=False:
    x = norm1(x + sa(x)); x = norm2(x + ff(x))) and torch/nn/functional.py multi_head_attention_forward
    (packed in_proj, q scaled by head_dim**-0.5 before q k^T, softmax, out_proj)."""
    x = patch.permute(1, 0, 2)                                            # [256,T,H]
    n, T, H = x.shape
    pos = w["model.position_matrix"][:T, :].unsqueeze(0).type_as(x)
    xp = x + pos
    pfx = "model.transformer_delta_encoder.layers.0."
    hd = H // nhead
    qkv = F.linear(xp, w[pfx + "self_attn.in_proj_weight"], w[pfx + "self_attn.in_proj_bias"])
    q, k, v = [t.view(n, T, nhead, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]   # [256,nhead,T,hd]
    att = F.softmax(torch.matmul(q * (hd ** -0.5), k.transpose(-1, -2)), dim=-1)
    a = torch.matmul(att, v).transpose(1, 2).reshape(n, T, H)
    a = F.linear(a, w[pfx + "self_attn.out_proj.weight"], w[pfx + "self_attn.out_proj.bias"])
    x1 = F.layer_norm(xp + a, (H,), w[pfx + "norm1.weight"], w[pfx + "norm1.bias"], eps)
    f = F.linear(F.relu(F.linear(x1, w[pfx + "linear1.weight"], w[pfx + "linear1.bias"])),
                 w[pfx + "linear2.weight"], w[pfx + "linear2.bias"])
    x2 = F.layer_norm(x1 + f, (H,), w[pfx + "norm2.weight"], w[pfx + "norm2.bias"], eps)
    return x2[:, -1, :] + torch.mean(x, dim=1)


def pool_patches(w: Dict[str, Tensor], feat: Tensor, method: str = "mean") -> Tensor:
    """valley_model.py:205-213: the four patch_pooling_method branches on cur_image_features[:,1:,:]  ([T,256,H])."""
    patch = feat[:, 1:, :]
    if method == "mean":
        return torch.mean(patch, dim=0)
    if method == "max":
        return torch.max(patch, dim=0)[0]
    if method == "temporal_importance":
        return text_importance_pooling(w, patch)
    if method == "temporal_transformer":
        return temporal_transformer_delta_adding(w, patch)
    raise ValueError(method)


def splice_one(ids: Tensor, embeds: Tensor, feat: Tensor, tok: SentinelIds, pooled: Optional[Tensor] = None) -> Tensor:
    """valley_model.py:203-245 for ONE multimodal sample.

    feat [T,257,H]: pooled = the patch_pooling_method over T of rows 1: ('mean' unless given); frames = row 0 of each
    frame.  Every <im_start> gets the same pooled block (:224-229); the video block is wrapped in a
    bare try/except (:231-244) so ANY failure silently yields the image-only result.
    """
    if pooled is None:
        pooled = torch.mean(feat[:, 1:, :], dim=0)                        # [256,H]   :207
    frames = feat[:, 0, :]                                                # [T,H]     :215
    npatch = pooled.shape[0]
    if (ids == tok.im_start_token).sum() != (ids == tok.im_end_token).sum():
        raise ValueError("The number of im_start_token and im_end_token should be the same")
    out = embeds.clone()
    for p in torch.where(ids == tok.im_start_token)[0].tolist():
        # ids[p + npatch + 1] raises IndexError if out of range, as in the reference (:226)
        if ids[p + npatch + 1] != tok.im_end_token:
            raise ValueError("Seems that the image is cut.")
        out = torch.cat((out[: p + 1], pooled, out[p + npatch + 1:]), dim=0)
    try:
        if (ids == tok.vi_start_token).sum() != (ids == tok.vi_end_token).sum():
            raise ValueError("The number of vi_start_token and vi_end_token should be the same")
        T = frames.shape[0]
        assert (ids == tok.vi_frame_token).sum() == T
        vid = out.clone()
        for q in torch.where(ids == tok.vi_start_token)[0].tolist():
            if ids[q + T + 1] != tok.vi_end_token:
                raise ValueError("Seems that the image is cut.")
            vid = torch.cat((vid[: q + 1], frames, vid[q + T + 1:]), dim=0)
    except Exception:                                                     # bare except in the reference
        vid = out.clone()
    return vid


def prepare_inputs_embeds(w: Dict[str, Tensor], input_ids: Tensor, image_features, tok: SentinelIds,
                          method: str = "mean") -> Tensor:
    """valley_model.py:155-160 + :192-247.  image_features as returned by encode_images."""
    embeds = F.embedding(input_ids, w["model.embed_tokens.weight"])
    out, cur = [], 0
    for ids, emb in zip(input_ids, embeds):
        if (ids == tok.im_patch_token).sum() == 0:                         # :198-202 (+0*dummy is exactly 0)
            out.append(emb)
            continue
        out.append(splice_one(ids, emb, image_features[cur], tok, pool_patches(w, image_features[cur], method)))
        cur += 1
    return torch.stack(out, dim=0)


# ----------------------------------------------------------------------------------------------
# LLaMA decoder
# ----------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """HF:models/llama/modeling_llama.py:62-67 -- variance in fp32, cast back, then * weight."""
    dt = x.dtype
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return weight * xf.to(dt)


def rope_cos_sin(positions: Tensor, head_dim: int, theta: float, dtype) -> Tuple[Tensor, Tensor]:
    """HF:modeling_llama.py:107-135: inv_freq fp32, freqs = pos*inv_freq (fp32), cat, cos/sin, cast to x dtype."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=positions.device).to(torch.float32) / head_dim))
    freqs = (inv[None, :, None] @ positions[:, None, :].to(torch.float32)).transpose(1, 2)   # [B,S,hd/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: Tensor) -> Tensor:
    """HF:modeling_llama.py:138-142."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class KVCache:
    """Growing K/V per layer == HF DynamicLayer.update (HF:cache_utils.py:102-120): torch.cat on dim -2."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[Tensor]] = [None] * n_layers
        self.v: List[Optional[Tensor]] = [None] * n_layers

    def get_seq_length(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[-2]

    def update(self, layer: int, k: Tensor, v: Tensor) -> Tuple[Tensor, Tensor]:
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat((self.k[layer], k), dim=-2)
            self.v[layer] = torch.cat((self.v[layer], v), dim=-2)
        return self.k[layer], self.v[layer]


def llama_layer(w: Dict[str, Tensor], x: Tensor, pfx: str, cos: Tensor, sin: Tensor, mask: Optional[Tensor],
                cache: Optional[KVCache], layer: int, heads: int, eps: float) -> Tensor:
    """HF:modeling_llama.py:303-333 (decoder layer), :251-289 (attention), :146-168 (RoPE),
    :199-222 (eager attention, softmax fp32), :182-184 (SwiGLU MLP)."""
    B, S, H = x.shape
    hd = H // heads
    h = rms_norm(x, w[pfx + "input_layernorm.weight"], eps)
    q = F.linear(h, w[pfx + "self_attn.q_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
    k = F.linear(h, w[pfx + "self_attn.k_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
    v = F.linear(h, w[pfx + "self_attn.v_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
    c, s_ = cos.unsqueeze(1), sin.unsqueeze(1)
    q = (q * c) + (rotate_half(q) * s_)
    k = (k * c) + (rotate_half(k) * s_)
    if cache is not None:
        k, v = cache.update(layer, k, v)
    att = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    if mask is not None:
        att = att + mask
    att = F.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, S, H)
    x = x + F.linear(o, w[pfx + "self_attn.o_proj.weight"])
    h = rms_norm(x, w[pfx + "post_attention_layernorm.weight"], eps)
    g = F.linear(h, w[pfx + "mlp.gate_proj.weight"])
    u = F.linear(h, w[pfx + "mlp.up_proj.weight"])
    return x + F.linear(F.silu(g) * u, w[pfx + "mlp.down_proj.weight"])


def llama_model(w: Dict[str, Tensor], inputs_embeds: Tensor, cache: Optional[KVCache], *, n_layers: int,
                heads: int, eps: float, theta: float = 10000.0, attention_mask: Optional[Tensor] = None) -> Tensor:
    """HF:modeling_llama.py:375-426 (LlamaModel.forward) with Valley's call (valley_model.py:249-254):
    position_ids = cache_len + arange(S) (never passed by Valley, Appendix A.8) -- NOT shifted by padding --, causal
    mask AND-ed with the 2-D ``attention_mask`` [B, past+S] over keys (HF masking_utils: padding mask), L decoder
    layers, final RMSNorm."""
    B, S, H = inputs_embeds.shape
    past = cache.get_seq_length() if cache is not None else 0
    dev = inputs_embeds.device          # (cpu in the tests; bench.py's eager-GPU baseline runs the same ops on cuda)
    pos = (past + torch.arange(S, device=dev))[None, :].expand(B, -1)
    cos, sin = rope_cos_sin(pos, H // heads, theta, inputs_embeds.dtype)
    mask = None
    if S > 1 or attention_mask is not None:
        neg = torch.finfo(inputs_embeds.dtype).min
        allowed = (torch.arange(past + S, device=dev)[None, :] <= (past + torch.arange(S, device=dev))[:, None])[None]      # [1,S,past+S]
        if attention_mask is not None:
            allowed = allowed & attention_mask[:, None, : past + S].bool().to(dev)                    # [B,S,past+S]
        mask = torch.zeros(allowed.shape, dtype=inputs_embeds.dtype, device=dev).masked_fill(~allowed, neg)[:, None]
    x = inputs_embeds
    for i in range(n_layers):
        x = llama_layer(w, x, f"model.layers.{i}.", cos, sin, mask, cache, i, heads, eps)
    return rms_norm(x, w["model.norm.weight"], eps)


# ----------------------------------------------------------------------------------------------
# ValleyLlamaForCausalLM.forward and the worker-style greedy loop
# ----------------------------------------------------------------------------------------------
class OracleConfig:
    def __init__(self, *, hidden_size, num_hidden_layers, num_attention_heads, intermediate_size, vocab_size,
                 rms_norm_eps=1e-5, rope_theta=10000.0, vit_layers=24, vit_heads=16, vit_patch=14,
                 vit_eps=1e-5, mm_vision_select_layer=-2, patch_pooling_method="mean"):
        self.hidden_size, self.num_hidden_layers = hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size = num_attention_heads, intermediate_size
        self.vocab_size, self.rms_norm_eps, self.rope_theta = vocab_size, rms_norm_eps, rope_theta
        self.vit_layers, self.vit_heads, self.vit_patch, self.vit_eps = vit_layers, vit_heads, vit_patch, vit_eps
        self.mm_vision_select_layer = mm_vision_select_layer
        self.patch_pooling_method = patch_pooling_method


def causal_lm_forward(w: Dict[str, Tensor], cfg: OracleConfig, tok: SentinelIds, input_ids: Tensor,
                      images=None, cache: Optional[KVCache] = None, attention_mask: Optional[Tensor] = None) -> Tensor:
    """ValleyLlamaForCausalLM.forward (valley_model.py:272-330) -> logits [B,S,V].
    Vision runs only when input_ids.shape[1] != 1 and images is not None (:163-164)."""
    if images is not None and input_ids.shape[1] != 1:
        feats = encode_images(w, images, cfg.mm_vision_select_layer, num_layers=cfg.vit_layers,
                              heads=cfg.vit_heads, patch=cfg.vit_patch, eps=cfg.vit_eps)
        embeds = prepare_inputs_embeds(w, input_ids, feats, tok, cfg.patch_pooling_method)
    else:
        embeds = F.embedding(input_ids, w["model.embed_tokens.weight"])
    hidden = llama_model(w, embeds, cache, n_layers=cfg.num_hidden_layers, heads=cfg.num_attention_heads,
                         eps=cfg.rms_norm_eps, theta=cfg.rope_theta, attention_mask=attention_mask)
    return F.linear(hidden, w["lm_head.weight"])                          # :304-305 (all positions)


@torch.no_grad()
def greedy_generate(w: Dict[str, Tensor], cfg: OracleConfig, tok: SentinelIds, input_ids: Tensor, images,
                    max_new_tokens: int, return_logits: bool = False, attention_mask: Optional[Tensor] = None):
    """The reference's own explicit decode loop, valley/serve/model_worker.py:371-397 with
    temperature < 1e-4 (argmax, :390-391), generalised from B=1 to B rows.  Step 0 = prefill
    with images; later steps feed the single new token with the cache.  ``attention_mask`` [B, S] (left padding)
    is extended by a column of ones per generated token (model_worker.py:382-383; HF generate does the same)."""
    cache = KVCache(cfg.num_hidden_layers)
    tokens, all_logits = [], []
    cur = input_ids
    for i in range(max_new_tokens):
        if attention_mask is not None and i > 0:
            attention_mask = torch.cat([attention_mask, torch.ones_like(attention_mask[:, :1])], dim=1)
        logits = causal_lm_forward(w, cfg, tok, cur, images if i == 0 else None, cache, attention_mask=attention_mask)
        last = logits[:, -1, :]
        nxt = torch.argmax(last, dim=-1)
        tokens.append(nxt)
        if return_logits:
            all_logits.append(last.to(torch.float32))
        cur = nxt[:, None]
    out = torch.stack(tokens, dim=1)
    return (out, torch.stack(all_logits, dim=1)) if return_logits else out


def causal_lm_loss(logits: Tensor, labels: Tensor, ignore_index: int = -100) -> Tensor:
    """valley_model.py:308-318: shift (tokens < n predict n), flatten, nn.CrossEntropyLoss() -- mean over the labels that are
    not ignore_index (-100, the IGNORE_INDEX the data pipeline writes over the prompt part)."""
    V = logits.shape[-1]
    sl = logits[..., :-1, :].contiguous().view(-1, V)
    tl = labels[..., 1:].contiguous().view(-1)
    return F.cross_entropy(sl, tl, ignore_index=ignore_index)
