"""Shared test plumbing: oracle config / sentinel ids from a ShapeSpec, bf16-rounded weights, error metrics."""
import torch

from oracle import valley_oracle as O
from valley_b200 import synthetic as syn


def oracle_cfg(spec):
    return O.OracleConfig(hidden_size=spec.hidden_size, num_hidden_layers=spec.num_hidden_layers,
                          num_attention_heads=spec.num_attention_heads, intermediate_size=spec.intermediate_size,
                          vocab_size=spec.vocab_size, rms_norm_eps=spec.rms_norm_eps, rope_theta=spec.rope_theta,
                          vit_layers=spec.vit_layers, vit_heads=spec.vit_heads, vit_patch=spec.vit_patch,
                          vit_eps=spec.vit_eps, mm_vision_select_layer=spec.mm_vision_select_layer,
                          patch_pooling_method=spec.patch_pooling_method)


def oracle_tok(spec):
    t = syn.sentinel_ids(spec)
    return O.SentinelIds(t["im_patch_token"], t["im_start_token"], t["im_end_token"], t["vi_frame_token"],
                         t["vi_start_token"], t["vi_end_token"])


def bf16_weights(spec, seed=0, **kw):
    """fp32 tensors holding bf16-representable values: what a bf16 checkpoint contains; handed to BOTH sides."""
    return {k: v.bfloat16().float() for k, v in syn.make_state_dict(spec, seed, **kw).items()}


def rel_fro(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def build_model(spec, sd, device=0):
    from valley_b200.model import ValleyConfig, ValleyLlamaForCausalLM
    m = ValleyLlamaForCausalLM.from_state_dict(ValleyConfig.from_spec(spec), sd, device=device)
    for k, v in syn.sentinel_ids(spec).items():
        setattr(m.get_model().vision_tower.config, k, v)
    return m
