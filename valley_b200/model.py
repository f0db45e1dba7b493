"""Reference-facing Python surface: the same class / method names, argument meaning and error
behaviour as valley/model/valley_model.py, with every tensor op executed by libvalley_b200.so.

    ValleyConfig                      valley_model.py:18
    ValleyLlamaModel                  valley_model.py:21   (.vision_tower.config sentinel ids, .forward)
    ValleyLlamaForCausalLM            valley_model.py:257  (.forward, .get_model, .generate,
                                      .prepare_inputs_for_generation, .build_inputs, .process_response)
  + encode_images / prepare_inputs_labels_for_multimodal: the two inline blocks valley_model.py:163-190
    and :192-247 factored out under the names BASELINE.json's north_star uses.

PyTorch is used here only for device memory, streams and (optionally) sampling; there is no
eager / CPU fallback for any op on the path.
"""
from __future__ import annotations

import os
import ctypes as C
import types
from typing import Iterable, List, Optional, Tuple, Union

import torch

from . import _lib
from ._lib import VlySampling, VlyConfig, VlyTokens, check

# valley/util/config.py:1-13
IGNORE_INDEX = -100
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
DEFAULT_VIDEO_FRAME_TOKEN = "<vi_frame>"
DEFAULT_VI_START_TOKEN = "<vi_start>"
DEFAULT_VI_END_TOKEN = "<vi_end>"

_DT = {torch.float32: _lib.VLY_F32, torch.bfloat16: _lib.VLY_BF16, torch.float16: _lib.VLY_F16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class ValleyConfig:
    """ValleyConfig(LlamaConfig), model_type 'valley' (valley_model.py:18-19) plus the extra keys the reference
    reads: mm_vision_tower, use_mm_proj, mm_hidden_size, mm_vision_select_layer, mm_use_im_start_end,
    use_patch_importance_pooling / use_delta_transformer (:40-52; the pooling variant is fixed at construction, as in the
    reference -- ``patch_pooling_method="max"`` stands for setting that attribute on the reference model)."""
    model_type = "valley"

    def __init__(self, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, intermediate_size=11008,
                 vocab_size=32008, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=2048,
                 mm_vision_tower="openai/clip-vit-large-patch14", mm_hidden_size=1024, mm_vision_select_layer=-2,
                 use_mm_proj=True, mm_use_im_start_end=True, vit_layers=24, vit_heads=16, vit_mlp=4096, vit_patch=14,
                 vit_image=224, vit_eps=1e-5, use_patch_importance_pooling=False, use_delta_transformer=False,
                 patch_pooling_method=None, bos_token_id=1, eos_token_id=2, pad_token_id=None, **kw):
        self.hidden_size, self.num_hidden_layers = hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size = num_attention_heads, intermediate_size
        self.vocab_size, self.rms_norm_eps, self.rope_theta = vocab_size, rms_norm_eps, rope_theta
        self.max_position_embeddings = max_position_embeddings
        self.mm_vision_tower, self.mm_hidden_size = mm_vision_tower, mm_hidden_size
        self.mm_vision_select_layer, self.use_mm_proj = mm_vision_select_layer, use_mm_proj
        self.mm_use_im_start_end = mm_use_im_start_end
        self.vit_layers, self.vit_heads, self.vit_mlp = vit_layers, vit_heads, vit_mlp
        self.vit_patch, self.vit_image, self.vit_eps = vit_patch, vit_image, vit_eps
        self.use_patch_importance_pooling, self.use_delta_transformer = use_patch_importance_pooling, use_delta_transformer
        if patch_pooling_method is None:          # valley_model.py:27, :40-52: the later flag wins
            patch_pooling_method = "temporal_transformer" if use_delta_transformer else (
                "temporal_importance" if use_patch_importance_pooling else "mean")
        if patch_pooling_method not in _lib.POOLING:
            raise ValueError(f"patch_pooling_method {patch_pooling_method!r} not in {sorted(_lib.POOLING)}")
        self.patch_pooling_method = patch_pooling_method
        # LlamaConfig defaults (HF:configuration_llama.py); HF generate() stops on config.eos_token_id unless told otherwise
        self.bos_token_id, self.eos_token_id, self.pad_token_id = bos_token_id, eos_token_id, pad_token_id
        self.use_return_dict, self.use_cache = True, True
        self.output_attentions = self.output_hidden_states = False
        for k, v in kw.items():
            setattr(self, k, v)

    @classmethod
    def from_spec(cls, spec, **kw):
        return cls(hidden_size=spec.hidden_size, num_hidden_layers=spec.num_hidden_layers,
                   num_attention_heads=spec.num_attention_heads, intermediate_size=spec.intermediate_size,
                   vocab_size=spec.vocab_size, rms_norm_eps=spec.rms_norm_eps, rope_theta=spec.rope_theta,
                   max_position_embeddings=spec.max_position_embeddings, mm_hidden_size=spec.vit_hidden,
                   mm_vision_select_layer=spec.mm_vision_select_layer, vit_layers=spec.vit_layers,
                   vit_heads=spec.vit_heads, vit_mlp=spec.vit_mlp, vit_patch=spec.vit_patch,
                   vit_image=spec.vit_image, vit_eps=spec.vit_eps,
                   patch_pooling_method=getattr(spec, "patch_pooling_method", "mean"),
                   **{"eos_token_id": None, **kw})          # synthetic specs have no eos: random-init ids must not stop a run


class CausalLMOutputWithPast(dict):
    """Attribute + index access like transformers.modeling_outputs.CausalLMOutputWithPast."""

    def __init__(self, loss=None, logits=None, past_key_values=None, hidden_states=None, attentions=None):
        super().__init__(loss=loss, logits=logits, past_key_values=past_key_values,
                         hidden_states=hidden_states, attentions=attentions)
        self.__dict__ = self

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in (self.loss, self.logits, self.past_key_values) if v is not None][k]
        return dict.__getitem__(self, k)


class _ShapeOnly:
    def __init__(self, shape):
        self.shape = torch.Size(shape)


class ValleyKVCache:
    """Opaque KV cache returned as ``past_key_values``.  Supports both access patterns callers use:
    ``past_key_values[0][0].shape[-2]`` (model_worker.py:253, :381, pinned-HF tuple cache) and
    ``.get_seq_length()`` (HF >= 4.36 DynamicCache).  Storage is the library's pre-allocated
    [L][2][B][heads][max_seq][128] bf16 buffer, appended in place by the QKV epilogues."""

    def __init__(self, model: "ValleyLlamaForCausalLM", batch: int, max_seq: int):
        self._model, self.batch, self.max_seq = model, batch, max_seq
        h = model._pop_handle(batch, max_seq)            # a handle (allocation + captured CUDA graph) freed by an earlier cache
        if h is None:
            h = C.c_void_p()
            check(model._lib.vly_kv_create(model._ctx, batch, max_seq, C.byref(h)))
        else:
            check(model._lib.vly_kv_reset(h, _stream()))
        self._h = h

    def get_seq_length(self, layer_idx: int = 0) -> int:
        n = C.c_int()
        check(self._model._lib.vly_kv_seq_len(self._h, C.byref(n)))
        return n.value

    def decode_kernel(self) -> str:
        """name of the kernel a decode step of this cache launches (for reports)"""
        buf = C.create_string_buffer(96)
        check(self._model._lib.vly_kv_decode_kernel(self._h, buf, 96))
        return buf.value.decode()

    def __len__(self):
        return self._model.config.num_hidden_layers

    def __bool__(self):
        return True

    def __getitem__(self, layer):
        c = self._model.config
        shp = (self.batch, c.num_attention_heads, self.get_seq_length(), c.hidden_size // c.num_attention_heads)
        return (_ShapeOnly(shp), _ShapeOnly(shp))

    def to_hf(self, layer: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Materialise layer's (key, value) in HuggingFace layout [B, heads, len, 128] (keys de-interleaved)."""
        c, n = self._model.config, self.get_seq_length()
        out = []
        for which in (0, 1):
            t = torch.empty(self.batch, c.num_attention_heads, n, 128, dtype=torch.bfloat16, device=self._model.device)
            check(self._model._lib.vly_kv_export(self._model._ctx, self._h, layer, which, t.data_ptr(), _stream()))
            out.append(t)
        return tuple(out)

    def reset(self):
        check(self._model._lib.vly_kv_reset(self._h, _stream()))

    def set_attention_mask(self, attention_mask: Optional[torch.Tensor], total_len: int):
        """HF's 2-D ``attention_mask`` [B, total_len] over cache positions 0..total_len-1 (past + new): a 0 means
        the key is never attended (left padding from ``build_inputs``).  Later positions stay attendable."""
        if attention_mask is None:
            return
        if attention_mask.dim() != 2 or tuple(attention_mask.shape) != (self.batch, total_len):
            raise ValueError(f"attention_mask shape {tuple(attention_mask.shape)} != (batch {self.batch}, past+new {total_len})")
        m = (attention_mask != 0).to(self._model.device, torch.uint8).contiguous()
        check(self._model._lib.vly_kv_set_key_mask(self._h, m.data_ptr(), total_len, _stream()))

    def __del__(self):
        # forward() without past_key_values hands a fresh cache to the caller on every request (model_worker.py:371-379);
        # when the caller drops it, the allocation (1 GB / sequence at 7B) goes back to the model instead of to cudaFree
        try:
            if getattr(self, "_h", None) is not None and self._model._ctx:
                if not self._model._push_handle(self.batch, self.max_seq, self._h):
                    self._model._lib.vly_kv_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _same_device(have: torch.device, want) -> bool:
    want = torch.device(want) if not isinstance(want, torch.device) else want
    return want.type == "cuda" and (want.index is None or want.index == have.index)


class _ModuleSurface:
    """The slice of ``nn.Module`` the reference's callers touch on the model and on ``vision_tower``
    (model_worker.py:78,86; run_valley.py:42-43; run_valley_conv.py:113,126): ``.to()``, ``.cuda()``, ``.half()``,
    ``.eval()``, ``.device``, ``.dtype``.  Weights live packed (bf16) inside the library on the device chosen at
    construction, so these calls VALIDATE and return self: a 16-bit float dtype is accepted (fp16 checkpoints were
    converted to bf16 at load), the owning CUDA device is accepted, anything else raises -- there is no CPU path."""
    device: torch.device
    dtype = torch.bfloat16
    training = False

    def to(self, *args, **kwargs):
        device, dtype = kwargs.get("device"), kwargs.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif isinstance(a, (str, torch.device, int)):
                device = a
            elif torch.is_tensor(a):
                device, dtype = a.device, a.dtype
        if device is not None:
            device = f"cuda:{device}" if isinstance(device, int) else device
            if not _same_device(self.device, device):
                raise _lib.VlyError(f"valley_b200 weights live on {self.device}; .to({device!r}) is not supported "
                                    "(there is no CPU path; build the model with device=... instead)")
        if dtype is not None and dtype not in (torch.float16, torch.bfloat16):
            raise _lib.VlyError(f".to({dtype}): the packed weights are bf16; only 16-bit float dtypes are accepted")
        return self

    def cuda(self, device=None):
        return self.to(device="cuda" if device is None else device)

    def half(self):
        return self.to(dtype=torch.float16)

    def bfloat16(self):
        return self.to(dtype=torch.bfloat16)

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise _lib.VlyError("valley_b200 is the inference hot path: forward only (SURVEY 8 f-4), no train() mode")
        return self.eval()

    def requires_grad_(self, requires_grad: bool = False):
        if requires_grad:
            raise _lib.VlyError("valley_b200 holds no autograd parameters")
        return self


class _VisionTower(_ModuleSurface):
    """Stands where CLIPVisionModel sits on the reference model: carries ``.config`` with the sentinel ids
    (run_valley.py:13-18, model_worker.py:80-84), takes the ``.to(device, dtype)`` the callers issue
    (model_worker.py:78, run_valley_conv.py:126) and is callable like the reference's use of it."""

    def __init__(self, model: "ValleyLlamaForCausalLM"):
        self._model = model
        self.device = model.device
        cfg = model.config
        self.config = types.SimpleNamespace(
            hidden_size=cfg.mm_hidden_size, image_size=cfg.vit_image, patch_size=cfg.vit_patch,
            num_hidden_layers=cfg.vit_layers, use_im_start_end=cfg.mm_use_im_start_end,
            im_patch_token=-1, im_start_token=-1, im_end_token=-1,
            _name_or_path=getattr(cfg, "mm_vision_tower", None))

    def __call__(self, pixel_values: torch.Tensor, output_hidden_states: bool = True, select_layer: Optional[int] = None):
        sel = self._model.config.mm_vision_select_layer if select_layer is None else select_layer
        hs = self._model._vit_encode(pixel_values, sel)
        return types.SimpleNamespace(selected_hidden_state=hs, select_layer=sel)


class _PackedLinear(_ModuleSurface):
    """Shape-carrying stand-in for an ``nn.Linear`` / ``nn.Embedding`` whose weight lives packed inside the library
    (fused / norm-folded / interleaved -- there is no per-module weight tensor to hand out)."""

    def __init__(self, owner, **dims):
        self.device = owner.device
        for k, v in dims.items():
            setattr(self, k, v)

    @property
    def weight(self):
        raise _lib.VlyError("weights are packed inside libvalley_b200.so (fused QKV, norm-folded, RoPE-interleaved); "
                            "load them with load_state_dict / from_pretrained -- they cannot be read back per module")


class ValleyLlamaModel(_ModuleSurface):
    """valley_model.py:21-254.  Holds the vision tower handle; ``forward`` returns final hidden states is NOT exposed
    separately (the final RMSNorm is folded into lm_head) -- callers in the reference only use the CausalLM wrapper.
    Plain attributes the callers set on it (``multi_image``, ``multi_image_mode``: model_worker.py:63-64 -- never read by
    the reference either) are accepted like on any Python object."""

    def __init__(self, owner: "ValleyLlamaForCausalLM"):
        self._owner = owner
        self.device = owner.device
        self.config = owner.config
        self.vision_tower = _VisionTower(owner)
        self.patch_pooling_method = owner.config.patch_pooling_method      # valley_model.py:27, :40-52
        self.mm_projector = _PackedLinear(owner, in_features=owner.config.mm_hidden_size, out_features=owner.config.hidden_size)
        self.embed_tokens = _PackedLinear(owner, num_embeddings=owner.config.vocab_size, embedding_dim=owner.config.hidden_size)


class KeywordsStoppingCriteria:
    """valley/util/data_util.py:40-56 (a ``transformers.StoppingCriteria``): stop when the text decoded from the ids generated
    so far contains a keyword.  Same quirk as the reference: the FIRST call only records the prompt length (so the first
    generated token is never tested on its own), later calls decode row 0 of ``output_ids[:, start_len:]``."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.tokenizer = tokenizer
        self.start_len = None
        self.input_ids = input_ids

    def __call__(self, output_ids, scores=None, **kwargs) -> bool:
        if self.start_len is None:
            self.start_len = self.input_ids.shape[1]
        else:
            outputs = self.tokenizer.batch_decode(output_ids[:, self.start_len:], skip_special_tokens=True)[0]
            for keyword in self.keywords:
                if keyword in outputs:
                    return True
        return False


_UNSET = object()


def _open_video_reader(path: str):
    """Default file reader of ``completion(tokenizer, path, ...)``: decord, exactly as load_video opens it
    (data_util.py:258-260).  Container decoding is CPU work outside the hot path; inject another reader factory through
    ``model.video_reader_factory`` (anything with ``len()``, ``get_batch(idx)`` -> [n,H,W,3] uint8, ``get_avg_fps()``)."""
    try:
        import decord
    except ImportError as e:
        raise ImportError("completion() was given a video path but decord is not installed; set model.video_reader_factory "
                          "to a callable path -> reader, or pass the decoded clip tensor [3,T,224,224]") from e
    return decord.VideoReader(path, num_threads=1, ctx=decord.cpu(0))


class ValleyLlamaForCausalLM(_ModuleSurface):
    """valley_model.py:257-439 behind libvalley_b200.so."""
    config_class = ValleyConfig
    video_reader_factory = staticmethod(_open_video_reader)

    def __init__(self, config: ValleyConfig, device: Union[int, str, torch.device] = 0):
        self._lib = _lib.load()
        self._ctx = None
        self.config = config
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _lib.VlyError("valley_b200 runs on a CUDA sm_100a device only; there is no CPU path")
        self.device = dev
        self.dtype = torch.bfloat16
        c = VlyConfig(config.hidden_size, config.num_hidden_layers, config.num_attention_heads, config.intermediate_size,
                      config.vocab_size, config.rms_norm_eps, config.rope_theta, config.max_position_embeddings,
                      config.mm_hidden_size, config.vit_layers, config.vit_heads, config.vit_mlp, config.vit_patch,
                      config.vit_image, config.vit_eps, config.mm_vision_select_layer, dev.index or 0,
                      _lib.POOLING[config.patch_pooling_method])
        h = C.c_void_p()
        check(self._lib.vly_create(C.byref(c), C.byref(h)))
        self._ctx = h
        self.model = ValleyLlamaModel(self)
        self.training = False
        self.logits_all_positions = True      # reference behaviour (valley_model.py:304-305); generate() uses last-only
        import threading
        self._cache_pool, self._pool_lock = {}, threading.Lock()
        self._free_handles = {}               # (batch, max_seq) -> [vly_kv handles] released by dropped ValleyKVCache objects

    # ---------------- lifetime / weights ----------------
    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                for lst in getattr(self, "_free_handles", {}).values():
                    for h in lst:
                        self._lib.vly_kv_destroy(h)
                self._free_handles = {}
                self._lib.vly_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def _pop_handle(self, batch: int, max_seq: int):
        with self._pool_lock:
            lst = self._free_handles.get((batch, max_seq))
            return lst.pop() if lst else None

    def _push_handle(self, batch: int, max_seq: int, h) -> bool:
        with self._pool_lock:
            lst = self._free_handles.setdefault((batch, max_seq), [])
            if len(lst) >= 2:
                return False
            lst.append(h)
            return True

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype=None, device=0, **kw) -> "ValleyLlamaForCausalLM":
        """``ValleyLlamaForCausalLM.from_pretrained(path, torch_dtype=torch.float16)`` (run_valley.py:39, model_worker.py:62): a
        local HF checkpoint directory (config.json + safetensors / .bin shards).  Weights are stored as bf16 whatever ``torch_dtype``
        says (fp16 checkpoints are converted on load)."""
        from . import checkpoint
        path = pretrained_model_name_or_path
        if checkpoint.is_lora_dir(path):       # run_valley.py:26-37: PeftModel.from_pretrained(base, path).merge_and_unload()
            base = checkpoint.resolve_lora_base(path)
            cfg = ValleyConfig(**{**checkpoint.read_config(base), **kw})
            m = cls(cfg, device)
            m.load_state_dict(checkpoint.iter_checkpoint_merged(base, path, device=m.device))
            return m
        cfg = ValleyConfig(**{**checkpoint.read_config(path), **kw})
        m = cls(cfg, device)
        m.load_state_dict(checkpoint.iter_checkpoint(path))
        return m

    @classmethod
    def from_state_dict(cls, config: ValleyConfig, state: Iterable, device=0) -> "ValleyLlamaForCausalLM":
        m = cls(config, device)
        m.load_state_dict(state)
        return m

    def load_state_dict(self, state, strict: bool = False):
        """Accepts a dict or an iterator of (hf_name, tensor).  Tensors may live on CPU or GPU, fp32/bf16/fp16.
        Unknown names (post_layernorm, rotary inv_freq, position_ids...) are ignored like HF non-strict loading."""
        items = state.items() if hasattr(state, "items") else state
        for name, t in items:
            if "post_layernorm" in name or name.endswith("position_ids") or name.endswith("inv_freq"):
                continue
            if "transforemr_adding_layer" in name:      # the template nn.TransformerEncoder deep-copies; never executed (valley_model.py:47-48)
                continue
            t = t.detach()
            if t.dtype not in _DT:
                t = t.float()
            t = t.to(self.device).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(self._lib.vly_load_weight(self._ctx, name.encode(), t.data_ptr(), _DT[t.dtype], shape, t.dim()))
        check(self._lib.vly_finalize_weights(self._ctx))
        return self

    def get_model(self) -> ValleyLlamaModel:      # valley_model.py:269
        return self.model

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return _PackedLinear(self, in_features=self.config.hidden_size, out_features=self.config.vocab_size)

    def resize_token_embeddings(self, new_num_tokens: Optional[int] = None):
        """Growing the embedding table is a training-time step (train.py:147 -> initialize_vision_tokenizer); released
        checkpoints already contain the added tokens.  Accepted as a no-op when nothing has to grow."""
        if new_num_tokens is not None and new_num_tokens > self.config.vocab_size:
            raise _lib.VlyError(f"resize_token_embeddings({new_num_tokens}) > loaded vocab {self.config.vocab_size}: growing the "
                                "embeddings is a training-time operation; load a checkpoint that already holds the added tokens")
        return self.model.embed_tokens

    def initialize_vision_tokenizer(self, tokenizer):
        """valley_model.py:354-379 for an inference checkpoint: register the sentinel tokens with the tokenizer and record
        their ids on ``vision_tower.config``.  The reference also grows the embeddings and initialises the new rows with the
        mean embedding -- that is checkpoint construction (training); here the tokenizer must fit the loaded vocabulary."""
        vc = self.get_model().vision_tower.config
        vc.use_im_start_end = True
        tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN, DEFAULT_VIDEO_FRAME_TOKEN], special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN, DEFAULT_VI_START_TOKEN, DEFAULT_VI_END_TOKEN],
                             special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        vc.im_start_token, vc.im_end_token = tokenizer.convert_tokens_to_ids([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN])
        vc.vi_start_token, vc.vi_end_token = tokenizer.convert_tokens_to_ids([DEFAULT_VI_START_TOKEN, DEFAULT_VI_END_TOKEN])
        vc.vi_frame_token = tokenizer.convert_tokens_to_ids(DEFAULT_VIDEO_FRAME_TOKEN)
        vc.im_patch_token = tokenizer.convert_tokens_to_ids([DEFAULT_IMAGE_PATCH_TOKEN])[0]

    def launches(self) -> int:
        n = C.c_int64()
        check(self._lib.vly_kernel_launch_count(self._ctx, C.byref(n)))
        return n.value

    # ---------------- vision ----------------
    def _tokens(self) -> VlyTokens:
        vc = self.model.vision_tower.config
        g = lambda k: int(getattr(vc, k, -1) if getattr(vc, k, None) is not None else -1)
        return VlyTokens(g("im_patch_token"), g("im_start_token"), g("im_end_token"),
                         g("vi_frame_token"), g("vi_start_token"), g("vi_end_token"))

    def _vit_encode(self, pixels: torch.Tensor, select_layer: int) -> torch.Tensor:
        """[F,3,224,224] -> hidden_states[select_layer] [F,257,1024] bf16."""
        if pixels.dim() != 4 or pixels.shape[1] != 3 or pixels.shape[2] != self.config.vit_image or pixels.shape[3] != self.config.vit_image:
            # HF:modeling_clip.py:203-207
            raise ValueError(f"Input image size ({pixels.shape[-2]}*{pixels.shape[-1]}) doesn't match model "
                             f"({self.config.vit_image}*{self.config.vit_image}).")
        if pixels.dtype not in _DT:
            pixels = pixels.float()
        pixels = pixels.to(self.device, non_blocking=True).contiguous()
        tokens = (self.config.vit_image // self.config.vit_patch) ** 2 + 1
        out = torch.empty(pixels.shape[0], tokens, self.config.mm_hidden_size, dtype=torch.bfloat16, device=self.device)
        check(self._lib.vly_vit_encode(self._ctx, pixels.data_ptr(), _DT[pixels.dtype], pixels.shape[0], select_layer,
                                       out.data_ptr(), _stream()))
        return out

    def encode_frames(self, pixels: torch.Tensor) -> torch.Tensor:
        """Vision tower only: [F,3,224,224] -> [F,257,1024] (config.mm_vision_select_layer)."""
        return self._vit_encode(pixels, getattr(self.config, "mm_vision_select_layer", -1))

    def _project(self, feats: torch.Tensor) -> torch.Tensor:
        rows = feats.numel() // feats.shape[-1]
        out = torch.empty(*feats.shape[:-1], self.config.hidden_size, dtype=torch.bfloat16, device=self.device)
        check(self._lib.vly_project(self._ctx, feats.data_ptr(), rows, out.data_ptr(), _stream()))
        return out

    @torch.no_grad()
    def encode_images(self, images):
        """valley_model.py:163-190: tensor [B,T,3,H,W] -> [B,T,257,hidden]; list of [T_i,3,H,W] -> list."""
        if isinstance(images, (list, tuple)):
            return [self._project(self.encode_frames(img)) for img in images]
        B, T = images.shape[:2]
        feats = self.encode_frames(images.reshape(B * T, *images.shape[2:]))
        return self._project(feats).view(B, T, feats.shape[1], self.config.hidden_size)

    def _pool_project(self, feats: torch.Tensor, n_videos: int, T: int) -> torch.Tensor:
        """feats [n_videos*T,257,1024] -> [n_videos, 256+T, hidden] (pool-first)."""
        rows = feats.shape[1] - 1 + T
        out = torch.empty(n_videos, rows, self.config.hidden_size, dtype=torch.bfloat16, device=self.device)
        check(self._lib.vly_pool_project(self._ctx, feats.data_ptr(), n_videos, T, out.data_ptr(), _stream()))
        return out

    def _splice_plan(self, input_ids: torch.Tensor, T: int):
        ids = input_ids.detach().to("cpu", torch.int64).contiguous()
        B, S = ids.shape
        smap = torch.empty(B, S, dtype=torch.int32)
        iidx = torch.empty(B, dtype=torch.int32)
        tok = self._tokens()
        check(self._lib.vly_build_splice_map(C.cast(ids.data_ptr(), C.POINTER(C.c_int64)), B, S, T, C.byref(tok),
                                             C.cast(smap.data_ptr(), C.POINTER(C.c_int32)),
                                             C.cast(iidx.data_ptr(), C.POINTER(C.c_int32))))
        return smap, iidx

    @torch.no_grad()
    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask=None, past_key_values=None, labels=None,
                                              images=None, frame_features: Optional[torch.Tensor] = None,
                                              n_frames: Optional[int] = None):
        """valley_model.py:155-247 -> (None, attention_mask, past_key_values, inputs_embeds, labels).

        Vision runs iff input_ids.shape[1] != 1 (or training) and images is not None (:163-164).
        ``frame_features`` ([n_videos*n_frames,257,1024], e.g. all-gathered from other ranks) skips the local ViT."""
        B, S = input_ids.shape
        ids_dev = input_ids.to(self.device, torch.int64).contiguous()
        embeds = torch.empty(B, S, self.config.hidden_size, dtype=torch.bfloat16, device=self.device)
        use_vision = (images is not None or frame_features is not None) and (S != 1 or self.training)
        if not use_vision:
            check(self._lib.vly_embed_splice(self._ctx, ids_dev.data_ptr(), None, None, None, 0, B, S, embeds.data_ptr(), _stream()))
            return None, attention_mask, past_key_values, embeds, labels
        if isinstance(images, (list, tuple)):
            # variable T per sample (valley_model.py:168-176): plan and pool per sample
            vis, smaps, cur = [], [], 0
            for b in range(B):
                T_b = images[min(cur, len(images) - 1)].shape[0]
                smap, iidx = self._splice_plan(input_ids[b:b + 1], T_b)
                if iidx[0] >= 0:
                    feats = self.encode_frames(images[cur])
                    vis.append(self._pool_project(feats, 1, T_b)[0])
                    cur += 1
                smaps.append((smap, int(iidx[0])))
            out = []
            for b, (smap, flag) in enumerate(smaps):
                e = torch.empty(1, S, self.config.hidden_size, dtype=torch.bfloat16, device=self.device)
                if flag < 0:
                    check(self._lib.vly_embed_splice(self._ctx, ids_dev[b:b + 1].data_ptr(), None, None, None, 0, 1, S, e.data_ptr(), _stream()))
                else:
                    v = vis.pop(0).contiguous()
                    sm = smap.to(self.device)
                    z = torch.zeros(1, dtype=torch.int32, device=self.device)
                    check(self._lib.vly_embed_splice(self._ctx, ids_dev[b:b + 1].data_ptr(), sm.data_ptr(), z.data_ptr(), v.data_ptr(),
                                                     v.shape[0], 1, S, e.data_ptr(), _stream()))
                out.append(e)
            return None, attention_mask, past_key_values, torch.cat(out, 0), labels
        if frame_features is None:
            Bi, T = images.shape[:2]
            frame_features = self.encode_frames(images.reshape(Bi * T, *images.shape[2:]))
        else:
            T = n_frames if n_frames is not None else images.shape[1]
            Bi = frame_features.shape[0] // T
        smap, iidx = self._splice_plan(input_ids, T)          # raises ValueError exactly where the reference does
        vis = self._pool_project(frame_features, Bi, T)        # [Bi, 256+T, H]
        n_mm = int((iidx >= 0).sum())
        if n_mm > Bi:
            raise IndexError("index out of range: more multimodal samples than images")   # image_features[cur_image_idx]
        sm, ii = smap.to(self.device, non_blocking=True), iidx.clamp(min=0).to(self.device, non_blocking=True)
        check(self._lib.vly_embed_splice(self._ctx, ids_dev.data_ptr(), sm.data_ptr(), ii.data_ptr(), vis.data_ptr(), vis.shape[1],
                                         B, S, embeds.data_ptr(), _stream()))
        return None, attention_mask, past_key_values, embeds, labels

    # ---------------- language model ----------------
    def new_cache(self, batch: int, max_seq: Optional[int] = None) -> ValleyKVCache:
        return ValleyKVCache(self, batch, max_seq or self.config.max_position_embeddings)

    def _borrow_cache(self, batch: int) -> ValleyKVCache:
        """generate() recycles its KV caches (and the CUDA graph captured on them) instead of paying a
        multi-GB cudaMalloc + graph capture per request.  The pool is per model instance and lock-protected
        (model_worker.py:467-474 calls the model from several threads)."""
        with self._pool_lock:
            lst = self._cache_pool.setdefault(batch, [])
            c = lst.pop() if lst else None
        if c is None:
            c = self.new_cache(batch)
        else:
            c.reset()
        return c

    def _return_cache(self, c: ValleyKVCache):
        with self._pool_lock:
            lst = self._cache_pool.setdefault(c.batch, [])
            if len(lst) < 2:
                lst.append(c)

    def _prefill(self, cache: ValleyKVCache, embeds: torch.Tensor, logits_mode: int):
        B, S, _ = embeds.shape
        V = self.config.vocab_size
        logits = None
        if logits_mode == 2:
            logits = torch.empty(B, S, V, dtype=torch.float32, device=self.device)
        elif logits_mode == 1:
            logits = torch.empty(B, 1, V, dtype=torch.float32, device=self.device)
        nxt = torch.empty(B, dtype=torch.int64, device=self.device)
        check(self._lib.vly_llama_prefill(self._ctx, cache._h, embeds.data_ptr(), B, S, logits_mode, _ptr(logits), nxt.data_ptr(), _stream()))
        return logits, nxt

    def _decode(self, cache: ValleyKVCache, tokens: torch.Tensor, want_logits: bool):
        B = tokens.shape[0]
        nxt = torch.empty(B, dtype=torch.int64, device=self.device)
        logits = torch.empty(B, 1, self.config.vocab_size, dtype=torch.float32, device=self.device) if want_logits else None
        tk = tokens.reshape(B).to(self.device, torch.int64).contiguous()
        check(self._lib.vly_llama_decode(self._ctx, cache._h, tk.data_ptr(), nxt.data_ptr(), _ptr(logits), _stream()))
        return logits, nxt

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, inputs_embeds=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, images=None, return_dict=None):
        """ValleyLlamaForCausalLM.forward (valley_model.py:272-330).  logits are fp32 [B,S,V]
        (the reference returns them in the model dtype; callers .float() them).  A 2-D attention_mask [B, past+S] masks
        keys exactly as HF does (padding mask AND causal mask; position ids are not shifted -- the reference never
        passes position_ids); rows that are themselves padding produce unspecified (finite) logits."""
        if output_hidden_states or output_attentions:
            # valley_model.py:285-300 forwards these to LlamaModel; the fused path never materialises per-layer hidden states
            # or attention probabilities (that is the point of it), so asking for them is an error rather than a silent None
            raise NotImplementedError("output_hidden_states / output_attentions are not produced by the fused kernels")
        if inputs_embeds is None:
            if input_ids is None:
                raise ValueError("You have to specify either input_ids or inputs_embeds")
            _, _, _, inputs_embeds, _ = self.prepare_inputs_labels_for_multimodal(
                input_ids, attention_mask, past_key_values, labels, images)
        else:
            if images is not None and input_ids is None:
                # the reference dereferences input_ids.shape here (valley_model.py:164)
                raise AttributeError("'NoneType' object has no attribute 'shape'")
            inputs_embeds = inputs_embeds.to(self.device, torch.bfloat16).contiguous()
        B, S, _ = inputs_embeds.shape
        cache = past_key_values if isinstance(past_key_values, ValleyKVCache) else None
        if cache is None:
            cache = self.new_cache(B)
        cache.set_attention_mask(attention_mask, cache.get_seq_length() + S)
        if S == 1 and cache.get_seq_length() > 0 and input_ids is not None:
            logits, nxt = self._decode(cache, input_ids, True)
        else:
            logits, nxt = self._prefill(cache, inputs_embeds, 2 if self.logits_all_positions else 1)
        loss = None
        if labels is not None:       # valley_model.py:308-318
            if logits.shape[1] != S:
                raise ValueError("labels need logits at every position (logits_all_positions=True, the reference behaviour)")
            lab = labels.to(self.device, torch.int64).contiguous()
            loss = torch.full((), float("nan"), dtype=torch.float32, device=self.device)     # S == 1: nothing to score (torch: nan)
            if S >= 2:
                check(self._lib.vly_cross_entropy(self._ctx, logits.data_ptr(), lab.data_ptr(), B, S, IGNORE_INDEX,
                                                  loss.data_ptr(), _stream()))
        out = CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=cache)
        out.next_tokens = nxt
        if return_dict is False:
            return tuple(v for v in (loss, logits, cache) if v is not None)
        return out

    __call__ = forward

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None, **kwargs):
        """valley_model.py:332-352 with the INTENDED semantics: slice to the last token only
        once the cache actually holds tokens."""
        if past_key_values is not None and (not hasattr(past_key_values, "get_seq_length") or past_key_values.get_seq_length() > 0):
            input_ids = input_ids[:, -1:]
        if inputs_embeds is not None and past_key_values is None:
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids}
        model_inputs.update({"past_key_values": past_key_values, "use_cache": kwargs.get("use_cache"),
                             "attention_mask": attention_mask, "images": kwargs.get("images", None)})
        return model_inputs

    @torch.no_grad()
    def generate(self, input_ids=None, images=None, max_new_tokens: int = 1024, do_sample: bool = False,
                 temperature: float = 1.0, stopping_criteria=None, eos_token_id=_UNSET, **kw):
        """Greedy (or temperature) generation == the loop of model_worker.py:371-397 / HF generate as called at
        valley_model.py:432.  Returns [B, S + n_new] like HF.  With no stopping criteria, decoding runs
        entirely on the device (CUDA-graph replay, no per-token host sync).  ``attention_mask`` [B, S] (left padding)
        is honoured like HF generate does; generated positions are always attendable.

        HF defaults that the reference's callers rely on are kept: ``eos_token_id`` / ``pad_token_id`` default to the config's
        (generation stops when every row has emitted eos; finished rows are padded), and when no ``attention_mask`` is given
        but the prompt contains ``pad_token_id`` (!= eos) the mask is inferred as ``input_ids != pad_token_id``
        (HF:generation/utils.py _prepare_attention_mask_for_generation).  Pass ``eos_token_id=None`` to run the full length."""
        B, S = input_ids.shape
        room = self.config.max_position_embeddings - S
        n_new = max(0, min(max_new_tokens, room))
        if n_new == 0:
            return input_ids.to(self.device)
        if eos_token_id is _UNSET:
            eos_token_id = getattr(self.config, "eos_token_id", None)
        pad_token_id = kw.get("pad_token_id", getattr(self.config, "pad_token_id", None))
        attention_mask = kw.get("attention_mask")
        if attention_mask is None and pad_token_id is not None and (eos_token_id is None or pad_token_id != eos_token_id):
            is_pad = input_ids == pad_token_id
            if bool(is_pad.any()):
                attention_mask = (~is_pad).to(torch.int64)
        _, _, _, embeds, _ = self.prepare_inputs_labels_for_multimodal(input_ids, None, None, None, images)
        cache = self._borrow_cache(B)
        try:
            cache.set_attention_mask(attention_mask, S)
            return self._generate_with_cache(cache, input_ids, embeds, n_new, do_sample, temperature, stopping_criteria, eos_token_id,
                                             pad_token_id)
        finally:
            self._return_cache(cache)

    def _generate_with_cache(self, cache, input_ids, embeds, n_new, do_sample, temperature, stopping_criteria, eos_token_id,
                             pad_token_id=None):
        B = input_ids.shape[0]
        greedy = (not do_sample) or temperature < 1e-4
        device_select = not stopping_criteria and not (greedy and eos_token_id is None) and B <= 64
        logits, nxt = self._prefill(cache, embeds, 0 if (greedy and not device_select) else 1)
        ids_dev = input_ids.to(self.device, torch.int64)
        if greedy and not stopping_criteria and eos_token_id is None:
            out = torch.empty(B, n_new, dtype=torch.int64, device=self.device)
            out[:, 0] = nxt
            if n_new > 1:
                rest = torch.empty(B, n_new - 1, dtype=torch.int64, device=self.device)
                check(self._lib.vly_generate_greedy(self._ctx, cache._h, nxt.data_ptr(), n_new - 1, rest.data_ptr(), _stream()))
                out[:, 1:] = rest
            return torch.cat([ids_dev, out], dim=1)
        if device_select:
            # temperature sampling and/or a stop token, still without a per-token host sync: the selection (Gumbel-max over
            # Philox noise) and the eos bookkeeping run inside the decode step; one 4-byte read at the end gives the length
            eos = -1 if eos_token_id is None else int(eos_token_id)
            pad = int(pad_token_id) if pad_token_id is not None else max(eos, 0)      # HF: pad defaults to eos
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())                         # torch.manual_seed() governs it
            sp = VlySampling(0.0 if greedy else float(temperature), seed, eos, pad)
            out = torch.empty(B, n_new, dtype=torch.int64, device=self.device)
            first = torch.empty(B, dtype=torch.int64, device=self.device)
            check(self._lib.vly_sample_logits(self._ctx, cache._h, logits.data_ptr(), C.byref(sp), first.data_ptr(), _stream()))
            out[:, 0] = first
            n_valid = 1
            if n_new > 1:
                rest = torch.empty(B, n_new - 1, dtype=torch.int64, device=self.device)
                done = torch.zeros(1, dtype=torch.int32, device=self.device)
                check(self._lib.vly_generate(self._ctx, cache._h, first.data_ptr(), n_new - 1, rest.data_ptr(), C.byref(sp),
                                             done.data_ptr(), _stream()))
                out[:, 1:] = rest
                n_valid += int(done.item())
            return torch.cat([ids_dev, out[:, :n_valid]], dim=1)
        # host-visible loop (stopping criteria present, or B > 64): one device->host sync per token, as in the reference.
        # HF semantics per row: a row that has emitted eos is finished and is fed / emits pad from then on
        seq = ids_dev
        finished = torch.zeros(B, dtype=torch.bool, device=self.device)
        pad = int(pad_token_id) if pad_token_id is not None else (int(eos_token_id) if eos_token_id is not None else 0)
        for i in range(n_new):
            if not greedy:
                probs = torch.softmax(logits[:, -1, :] / temperature, dim=-1)    # model_worker.py:393-394
                nxt = torch.multinomial(probs, num_samples=1).reshape(B)
            if eos_token_id is not None:
                nxt = torch.where(finished, torch.full_like(nxt, pad), nxt)
                finished = finished | (nxt == eos_token_id)
            seq = torch.cat([seq, nxt[:, None]], dim=1)
            if eos_token_id is not None and bool(finished.all()):
                break
            if stopping_criteria and any(sc(seq, None) for sc in stopping_criteria):
                break
            if i + 1 < n_new:
                logits, nxt = self._decode(cache, nxt, not greedy)
        return seq

    # ---------------- prompt helpers (pure string logic; valley_model.py:381-422) ----------------
    def build_inputs(self, tokenizer, messages):
        prompt = ''
        for m in messages:
            if m['role'] == 'system':
                prompt += m['content'] + '\n\n' + '###'
            elif m['role'] == 'user':
                replace_token = DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_PATCH_TOKEN * 256 + DEFAULT_IM_END_TOKEN + \
                    DEFAULT_VI_START_TOKEN + DEFAULT_VIDEO_FRAME_TOKEN * 8 + DEFAULT_VI_END_TOKEN
                if '<video>' in m['content'] or '<image>' in m['content']:
                    message = m['content'].replace('<video>', replace_token).replace('<image>', replace_token)
                    prompt += ' ' + 'Human' + ": " + message + ' \n' + '###'
            elif m['role'] == 'assistent':
                prompt += ' ' + 'Assistent' + ": " + m['content'] + ' \n' + '###'
            else:
                raise ValueError("Role is only suport \"assistent\", \"human\" and \"system\".")
        if DEFAULT_IM_START_TOKEN not in prompt:
            raise ValueError("You need to specify the <video> token in the query")
        tokenizer.padding_side = 'left'
        return tokenizer([prompt], padding=True)

    def process_response(self, outputs):
        output = []
        for out in outputs:
            while True:
                cur_len = len(out)
                out = out.strip()
                for pattern in ['###', 'Assistant:', 'Response:', 'Valley:']:
                    if out.startswith(pattern):
                        out = out[len(pattern):].strip()
                if len(out) == cur_len:
                    break
            if '###' not in out:
                out += '###'
            output.append(out[:out.index('###')].strip())
        return output

    @torch.no_grad()
    def completion(self, tokenizer, video, message: list, gen_kwargs: dict, device=None):
        """valley_model.py:424-439, same call: ``model.completion(tokenizer, args.video_file, message, gen_kwargs, device)``
        (run_valley.py:56).  ``video`` is a file path -- opened by ``self.video_reader_factory`` (decord by default, as
        load_video does; container decoding is CPU work outside the hot path) and then sampled (8 frames, 'fixed'),
        resized, cropped and normalised ON THE DEVICE, bit-identical to load_video (valley_b200/video.py) -- or the decoded
        clip tensor [3,T,224,224] load_video would have returned.  Generation stops like the reference's: the
        ``KeywordsStoppingCriteria(['###'])`` it passes (:431-432) plus HF generate's default stop on ``config.eos_token_id``
        (overridable through ``gen_kwargs``)."""
        inputs = self.build_inputs(tokenizer, message)
        input_ids = torch.as_tensor(inputs.input_ids).to(self.device)
        if torch.is_tensor(video):
            images = video.permute(1, 0, 2, 3).unsqueeze(0).half().to(self.device)
        else:
            from . import video as _video
            if os.path.isdir(str(video)):              # load_video's directory-of-images branch (data_util.py:282-302)
                images = _video.load_image_dir(str(video), getattr(self, "image_processor", None)).unsqueeze(0).half().to(self.device)
            else:
                reader = self.video_reader_factory(str(video))
                images = _video.load_video(self, reader, "fixed", 8, dtype=torch.float16).unsqueeze(0)      # [1,T,3,224,224]
        gen_kwargs = dict(gen_kwargs)
        if "attention_mask" not in gen_kwargs and getattr(inputs, "attention_mask", None) is not None:
            am = torch.as_tensor(inputs.attention_mask)
            if bool((am == 0).any()):               # build_inputs pads on the left (valley_model.py:400-402); B == 1 -> no padding
                gen_kwargs["attention_mask"] = am
        if "eos_token_id" not in gen_kwargs and getattr(tokenizer, "eos_token_id", None) is not None \
                and getattr(self.config, "eos_token_id", None) is None:
            gen_kwargs["eos_token_id"] = tokenizer.eos_token_id
        stopping_criteria = KeywordsStoppingCriteria(['###'], tokenizer, input_ids)
        output_ids = self.generate(input_ids=input_ids, images=images, stopping_criteria=[stopping_criteria], **gen_kwargs)
        input_token_len = input_ids.shape[1]
        n_diff_input_output = (input_ids != output_ids[:, :input_token_len]).sum().item()
        if n_diff_input_output > 0:
            print(f'[Warning] {n_diff_input_output} output_ids are not the same as the input_ids')
        outputs = tokenizer.batch_decode(output_ids[:, input_token_len:], skip_special_tokens=True)
        return self.process_response(outputs)
