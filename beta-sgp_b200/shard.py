"""Sharding of independent restorations over the GPUs of one box.

Every stamp / subdivision / beta-init value is its own optimisation problem (the reference loops over
them one at a time: application_sgp_star_stamps.py:56-105, application_sgp_subdivisions.py:83-107), so
the batch index is simply dealt round-robin to the ranks and nothing is exchanged until the final
gather of the restored images and their per-image scalars.  One process per GPU; torch.distributed is
used for the gather only (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_items, rank, world_size):
    """Indices of the batch entries owned by `rank`: rank, rank + W, rank + 2W, ...  Round-robin spreads
    the (strongly varying, 2..160) iteration counts evenly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return np.arange(rank, n_items, world_size)


def shard_counts(n_items, world_size):
    return [len(range(r, n_items, world_size)) for r in range(world_size)]


def gather_to_all(local, n_items, rank, world_size, group=None):
    """All-gather per-image results.  `local` maps name -> tensor [n_local, ...] for the images of
    shard_indices(n_items, rank, world_size); returns name -> tensor [n_items, ...] in the original batch
    order on every rank.  Shards are padded to the largest shard so that one all_gather per field
    suffices."""
    import torch
    import torch.distributed as dist
    counts = shard_counts(n_items, world_size)
    cap = max(counts)
    out = {}
    for name, t in local.items():
        if t.shape[0] != counts[rank]:
            raise ValueError(f"{name}: expected {counts[rank]} local rows, got {t.shape[0]}")
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        if world_size == 1:
            parts = [pad]
        else:
            parts = [torch.empty_like(pad) for _ in range(world_size)]
            dist.all_gather(parts, pad, group=group)
        full = torch.empty((n_items,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for r in range(world_size):
            idx = torch.as_tensor(shard_indices(n_items, r, world_size), device=t.device, dtype=torch.long)
            full[idx] = parts[r][:counts[r]]
        out[name] = full
    return out
