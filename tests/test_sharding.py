"""Multi-GPU host logic on CPU: round-robin sharding of independent images and the final gather,
with torch.distributed over gloo (world_size 2 and 3)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_module():
    spec = importlib.util.spec_from_file_location("_bsgp_shard", os.path.join(ROOT, "beta-sgp_b200", "shard.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shard_indices_partition():
    sh = _shard_module()
    for n in (1, 7, 320, 8192):
        for w in (1, 2, 3, 8):
            parts = [sh.shard_indices(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert [len(p) for p in parts] == sh.shard_counts(n, w)
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _worker(rank, world, n, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _shard_module()
    idx = sh.shard_indices(n, rank, world)
    # stand-in for the per-image solve: a deterministic function of the global image index
    x = torch.stack([torch.full((4, 4), float(i)) for i in idx]) if len(idx) else torch.zeros(0, 4, 4)
    iters = torch.tensor([3 * int(i) + 1 for i in idx], dtype=torch.int32)
    full = sh.gather_to_all({"x": x, "iters": iters}, n, rank, world)
    ok = bool((full["x"][:, 0, 0] == torch.arange(n, dtype=torch.float32)).all()) and \
        bool((full["iters"] == 3 * torch.arange(n, dtype=torch.int32) + 1).all())
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 11), (3, 7)])
def test_gather_over_gloo(world, n):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() + world * 17 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, n, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))
