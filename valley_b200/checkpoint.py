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
        files = sorted(f for f in glob.glob(os.path.join(path, pattern)) if not os.path.basename(f).startswith("adapter_model"))
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


# ----------------------------------------------------------------------------------------------
# LoRA adapters (run_valley.py:26-37: PeftModel.from_pretrained(base, path).merge_and_unload())
# ----------------------------------------------------------------------------------------------
def is_lora_dir(path: str) -> bool:
    return os.path.exists(os.path.join(path, "adapter_config.json"))


def read_lora(path: str):
    """-> (scaling, {hf_weight_name: (A [r,in], B [out,r])}).  peft's layout: ``adapter_config.json`` (r, lora_alpha) and
    ``adapter_model.{safetensors,bin}`` with keys ``base_model.model.<module>.lora_{A,B}[.default].weight``.
    peft itself is not in this image, so this follows its documented merge (LoraLayer.get_delta_weight:
    ``W += (B @ A) * lora_alpha / r``; train.py:153-157 uses r=16, alpha=32 on the seven decoder projections) and is NOT pinned
    against a peft run -- stated in DESIGN.md."""
    with open(os.path.join(path, "adapter_config.json")) as f:
        ac = json.load(f)
    if ac.get("fan_in_fan_out"):
        raise ValueError("fan_in_fan_out LoRA adapters are not supported (Linear layers only)")
    scaling = float(ac["lora_alpha"]) / float(ac["r"])
    st = os.path.join(path, "adapter_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in sd.items():
        for tag in ("lora_A", "lora_B"):
            mark = f".{tag}."
            if mark in k:
                mod = k.split(mark)[0]
                if mod.startswith("base_model.model."):
                    mod = mod[len("base_model.model."):]
                pairs.setdefault(mod + ".weight", {})[tag] = v
    out = {}
    for name, ab in pairs.items():
        if set(ab) != {"lora_A", "lora_B"}:
            raise ValueError(f"LoRA adapter for {name} is missing {'lora_B' if 'lora_A' in ab else 'lora_A'}")
        out[name] = (ab["lora_A"], ab["lora_B"])
    return scaling, out


def iter_checkpoint_merged(base_path: str, lora_path: str, device="cpu") -> Iterator[Tuple[str, torch.Tensor]]:
    """The base checkpoint with every adapted weight replaced by ``W + (B @ A) * scaling`` (fp32 accumulate, cast back)."""
    scaling, lora = read_lora(lora_path)
    seen = set()
    for name, w in iter_checkpoint(base_path):
        if name in lora:
            a, b = lora[name]
            if tuple(w.shape) != (b.shape[0], a.shape[1]) or a.shape[0] != b.shape[1]:
                raise ValueError(f"LoRA shapes {tuple(a.shape)}, {tuple(b.shape)} do not fit {name} {tuple(w.shape)}")
            delta = (b.to(device, torch.float32) @ a.to(device, torch.float32)) * scaling
            w = (w.to(device, torch.float32) + delta).to(w.dtype)
            seen.add(name)
        yield name, w
    missing = set(lora) - seen
    if missing:
        raise KeyError(f"LoRA adapter targets weights that are not in the base checkpoint: {sorted(missing)[:3]} ...")


def resolve_lora_base(path: str) -> str:
    """run_valley.py:27-31: the adapter directory itself when it also holds a config.json, else adapter_config's
    base_model_name_or_path."""
    if os.path.exists(os.path.join(path, "config.json")):
        return path
    with open(os.path.join(path, "adapter_config.json")) as f:
        return json.load(f)["base_model_name_or_path"]
