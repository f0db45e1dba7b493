"""Multi-GPU data parallelism for the hot path: one process per GPU, frames sharded over ranks,
ONE all-gather of frame embeddings before pool + projection (BASELINE.json north_star).

The reference has no inference-time collective at all; this is new capability.
Every frame's ViT forward is independent (valley_model.py:179-184 loops over batch items), and every
sequence's decode is independent, so the only exchange step is the gather of ``hidden_states[select]``
shards ([F_local,257,1024] bf16) that lets each rank pool/splice the videos it decodes.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing; the gather is not fused with
the last ViT GEMM yet ("what comes next").
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n_items over world ranks: the first (n_items % world) ranks get one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_bounds(n_items, world, r)[1] - shard_bounds(n_items, world, r)[0] for r in range(world)]


def gather_frame_features(local_feats: torch.Tensor, n_frames_total: int, group=None) -> torch.Tensor:
    """All-gather variable-size shards of frame features -> [n_frames_total, tokens, D] on every rank.

    Shards are padded to the largest shard so a single all_gather_into_tensor (one NCCL ncclAllGather)
    moves everything; padding rows are dropped afterwards."""
    world = dist.get_world_size(group)
    sizes = all_shard_sizes(n_frames_total, world)
    mx = max(sizes)
    assert local_feats.shape[0] == sizes[dist.get_rank(group)], (local_feats.shape, sizes)
    if local_feats.shape[0] < mx:
        pad = torch.zeros(mx - local_feats.shape[0], *local_feats.shape[1:], dtype=local_feats.dtype, device=local_feats.device)
        local_feats = torch.cat([local_feats, pad], 0)
    out = torch.empty(world * mx, *local_feats.shape[1:], dtype=local_feats.dtype, device=local_feats.device)
    dist.all_gather_into_tensor(out, local_feats.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], 0)


def encode_frames_sharded(encode_fn: Callable[[torch.Tensor], torch.Tensor], local_pixels: torch.Tensor,
                          n_frames_total: int, group=None) -> torch.Tensor:
    """rank r holds pixels of global frames shard_bounds(n_frames_total, world, r); returns ALL frame features."""
    local = encode_fn(local_pixels) if local_pixels.shape[0] > 0 else \
        torch.zeros(0, 257, 1024, dtype=torch.bfloat16, device=local_pixels.device)
    return gather_frame_features(local, n_frames_total, group)


def my_videos(n_videos: int, group=None) -> Tuple[int, int]:
    """Videos whose sequences this rank decodes (LLM replicated per GPU, batch sharded; no further collective)."""
    return shard_bounds(n_videos, dist.get_world_size(group), dist.get_rank(group))


def generate_sharded(model, input_ids: torch.Tensor, local_pixels: torch.Tensor, n_videos: int, n_frames: int,
                     max_new_tokens: int, group=None) -> torch.Tensor:
    """Config-4 style request: ``n_videos`` videos x ``n_frames`` frames, frames sharded over ranks for the ViT,
    one all-gather, then every rank pools/projects/splices and greedy-decodes its own videos.
    ``input_ids`` [n_videos_local, S] are this rank's prompts; returns this rank's generated ids."""
    feats = encode_frames_sharded(model.encode_frames, local_pixels, n_videos * n_frames, group)
    lo, hi = my_videos(n_videos, group)
    mine = feats.view(n_videos, n_frames, *feats.shape[1:])[lo:hi].reshape((hi - lo) * n_frames, *feats.shape[1:]).contiguous()
    B = hi - lo
    _, _, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(input_ids, None, None, None, None,
                                                                    frame_features=mine, n_frames=n_frames)
    cache = model._borrow_cache(B)
    try:
        S = input_ids.shape[1]
        return model._generate_with_cache(cache, input_ids, embeds, max_new_tokens, False, 1.0, None, None)[:, S:]
    finally:
        model._return_cache(cache)
