"""Clip sharding across ranks: clips are independent, so each rank runs whole
clips with replicated weights and NO collective inside the hot path; per-clip panoptic id maps are
exchanged with ONE all_gather after the shard finishes."""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Contiguous blocks (keeps a video's frames on one rank when clips are listed video-major);
    the first ``n_clips % world`` ranks get one extra clip."""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


WIRE_DTYPE = torch.int16      # panoptic ids are < stuff_num + n_slots (vps_temporal_slots.py:428): 2 bytes on the wire


def gather_id_maps(local: torch.Tensor, n_clips: int, dist=None, wire_dtype=None) -> torch.Tensor:
    """local [k_rank,H,W] id maps of this rank's shard -> [n_clips,H,W] on every rank (clip order).

    Shards may differ by one clip; they are padded to the largest so a single all_gather suffices.
    ``wire_dtype`` (e.g. WIRE_DTYPE) narrows the int64 maps for the exchange -- 4x fewer bytes over NVLink -- and
    the result is returned in that dtype (ids are small non-negative integers, so the cast is lossless)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local if wire_dtype is None else local.to(wire_dtype)
    world, rank = dist.get_world_size(), dist.get_rank()
    kmax = (n_clips + world - 1) // world
    if wire_dtype is not None:
        local = local.to(wire_dtype)
    pad = torch.zeros((kmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    flat = torch.empty((world * kmax,) + tuple(pad.shape[1:]), dtype=local.dtype, device=local.device)
    # concatenation along dim 0 (accepted by gloo and nccl); exchanged as raw bytes: neither backend has int16
    dist.all_gather_into_tensor(flat.view(torch.uint8), pad.view(torch.uint8))
    out = flat.view((world, kmax) + tuple(pad.shape[1:]))
    parts = [out[r, :len(shard_clips(n_clips, r, world))] for r in range(world)]
    return torch.cat(parts, 0)
