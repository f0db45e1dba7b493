"""Reading a Valley checkpoint directory the way ``ValleyLlamaForCausalLM.from_pretrained(path)`` does in the reference
(HF ``save_pretrained`` layout: ``config.json`` + ``model*.safetensors`` / ``pytorch_model*.bin`` shards, optionally with an
``*.index.json``).  Pure host logic: tensors are streamed one at a time (a 13B checkpoint never sits in host memory twice)
and handed to ``vly_load_weight`` under their HuggingFace names.  The vision tower is a registered sub-module of the
reference model (valley_model.py:38), so its weights are in the same files under ``model.vision_tower.vision_model.*``.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict, Iterator, List, Tuple

import torch


def read_config(path: str) -> Dict:
    with open(os.path.join(path, "config.json")) as f:
        cfg = json.load(f)
    # CLIP geometry: a local mm_vision_tower directory carries its own config.json; a hub id means ViT-L/14 (the only tower the
    # reference supports: valley_stage2.yaml:42, the hard-coded zeros(256, 1024) at valley_model.py:192)
    vt = cfg.get("mm_vision_tower")
    vt_cfg = os.path.join(vt, "config.json") if isinstance(vt, str) and os.path.isdir(vt) else None
    if vt_cfg and os.path.exists(vt_cfg):
        with open(vt_cfg) as f:
            v = json.load(f)
        v = v.get("vision_config", v)
        cfg.setdefault("mm_hidden_size", v.get("hidden_size", 1024))
        cfg.update(vit_layers=v.get("num_hidden_layers", 24), vit_heads=v.get("num_attention_heads", 16),
                   vit_mlp=v.get("intermediate_size", 4096), vit_patch=v.get("patch_size", 14), vit_image=v.get("image_size", 224),
                   vit_eps=v.get("layer_norm_eps", 1e-5))
    return cfg


def weight_files(path: str) -> List[str]:
    """Shard list in a deterministic order; safetensors preferred when both formats are present (as HF does)."""
    for index, pattern in (("model.safetensors.index.json", "*.safetensors"), ("pytorch_model.bin.index.json", "pytorch_model*.bin")):
        ip = os.path.join(path, index)
        if os.path.exists(ip):
            with open(ip) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
            return [os.path.join(path, x) for x in files]
        files = sorted(glob.glob(os.path.join(path, pattern)))
        if files:
            return files
    raise FileNotFoundError(f"no model*.safetensors / pytorch_model*.bin under {path}")


def iter_checkpoint(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    for f in weight_files(path):
        if f.endswith(".safetensors"):
            from safetensors import safe_open
            with safe_open(f, framework="pt", device="cpu") as sf:
                for name in sf.keys():
                    yield name, sf.get_tensor(name)
        else:
            sd = torch.load(f, map_location="cpu", mmap=True, weights_only=True)
            for name in list(sd.keys()):
                yield name, sd.pop(name)
