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


def solve_batch_sharded(gn, psf, bkg, flux=None, betaParam=1.005, group=None, device=None, **kw):
    """Every rank passes the FULL batch (numpy arrays) and gets the full result back: rank r restores the images
    shard_indices(B, r, W) on its own GPU (engine.solve_batch, one persistent kernel launch) and the per-image
    results are all-gathered (NCCL when the process group is NCCL, gloo otherwise).  No data-path collective:
    the images are independent problems (application_sgp_star_stamps.py:56-105, application_sgp_subdivisions.py:83-107).

    psf: [ny,nx] shared or [B,ny,nx]; bkg: scalar, [B] or [B,ny,nx]; flux / betaParam: scalar or [B]."""
    import torch
    import torch.distributed as dist
    from . import engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if device is None:
        device = int(__import__("os").environ.get("LOCAL_RANK", rank))
    gn = np.asarray(gn)
    B = gn.shape[0]
    idx = shard_indices(B, rank, world)

    def take(a, per_image_ndim):
        a = np.asarray(a)
        return a[idx] if a.ndim == per_image_ndim + 1 and a.shape[0] == B else a

    psf = np.asarray(psf)
    flux_l = None if flux is None else np.broadcast_to(np.asarray(flux, dtype=np.float64).reshape(-1), (B,))[idx]
    beta_l = np.broadcast_to(np.asarray(betaParam, dtype=np.float64).reshape(-1), (B,))[idx]
    bkg = np.asarray(bkg)
    bkg_l = bkg[idx] if bkg.ndim >= 1 and bkg.shape[0] == B and bkg.size > 1 else bkg
    r = engine.solve_batch(gn[idx], psf[idx] if psf.ndim == 3 else psf, bkg_l, flux=flux_l, betaParam=beta_l, device=device, **kw)
    use_cuda = dist.is_initialized() and dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    local = {k: torch.as_tensor(np.ascontiguousarray(getattr(r, k))).to(dev)
             for k in ("x", "iters", "status", "discr", "times", "beta_final", "proj_evals", "ls_trials")}
    full = gather_to_all(local, B, rank, world, group=group)
    return {k: v.cpu().numpy() for k, v in full.items()}
