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
            for ranked in (False, True):
                beta = np.random.default_rng(n + w).normal(1.0, 0.05, n)
                cr = sh.expected_cost_rank(n, beta) if ranked else None
                parts = [sh.shard_indices(n, r, w, cr) for r in range(w)]
                assert sorted(np.concatenate(parts).tolist()) == list(range(n))
                assert [len(p) for p in parts] == sh.shard_counts(n, w, snake=ranked)
                assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_cost_ranked_dealing_balances_the_beta_groups():
    """320 subdivisions x 5 beta inits over 8 ranks: every rank gets the same number of solves of each beta value (the
    iteration count grows as beta -> 1), which plain round-robin (index mod 8 on a period-5 pattern) also gives, but a
    batch sorted by tile does not guarantee; and the expected-longest solves are spread one per rank."""
    sh = _shard_module()
    b5 = np.array([1.0882, 1.0248, 0.9789, 1.0068, 0.9551])
    beta = np.tile(b5, 64)
    cr = sh.expected_cost_rank(320, beta)
    assert np.all(np.abs(beta[cr][:64] - 1.0) == np.abs(b5 - 1.0).min())
    for r in range(8):
        idx = sh.shard_indices(320, r, 8, cr)
        assert len(idx) == 40
        assert [int(np.sum(beta[idx] == b)) for b in b5] == [8] * 5


def _worker(rank, world, n, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _shard_module()
    idx = sh.shard_indices(n, rank, world)                 # plain round-robin (no cost ranking)
    # stand-in for the per-image solve: a deterministic function of the global image index
    x = torch.stack([torch.full((4, 4), float(i)) for i in idx]) if len(idx) else torch.zeros(0, 4, 4)
    iters = torch.tensor([3 * int(i) + 1 for i in idx], dtype=torch.int32)
    full = sh.gather_to_all({"x": x, "iters": iters}, n, rank, world)
    ok = bool((full["x"][:, 0, 0] == torch.arange(n, dtype=torch.float32)).all()) and \
        bool((full["iters"] == 3 * torch.arange(n, dtype=torch.int32) + 1).all())
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def _fake_result(bs, gn, psf, bkg, flux, betaParam, x0, obj, maxit):
    """Stand-in for the CUDA solve (there is no CPU path): a deterministic function of EVERY per-image argument, so a
    wrongly sliced argument shows up in the gathered result."""
    B = gn.shape[0]
    psf_s = psf.sum(axis=(-2, -1)) if psf.ndim == 3 else np.full(B, psf.sum())
    bk = bkg if np.ndim(bkg) == 3 else np.broadcast_to(np.asarray(bkg, dtype=np.float64).reshape(-1), (B,))[:, None, None]
    fl = np.broadcast_to(np.asarray(flux, dtype=np.float64).reshape(-1), (B,))
    b0 = np.broadcast_to(np.asarray(betaParam, dtype=np.float64).reshape(-1), (B,))
    x = gn * b0[:, None, None] + bk + psf_s[:, None, None] + (0 if x0 is None else 3.0 * x0) + (0 if obj is None else 7.0 * obj)
    discr = np.zeros((B, maxit + 1)); discr[:, 0] = fl; discr[:, 1] = b0
    return bs.BatchResult(x=x, iters=np.round(fl).astype(np.int32), status=np.zeros(B, np.int32), discr=discr, times=discr * 2, stop_value=None,
                          err=None, beta_final=b0 + 1.0, proj_evals=np.round(fl).astype(np.int32) * 2, ls_trials=np.ones(B, np.int32), scalars=None)


def _inputs(n):
    rng = np.random.default_rng(n)
    return dict(gn=rng.random((n, 4, 6)), psf=rng.random((n, 4, 6)), bkg=rng.random((n, 4, 6)), flux=np.arange(n) + 5.0,
                betaParam=rng.normal(1.0, 0.05, n), x0=rng.random((n, 4, 6)), obj=rng.random((n, 4, 6)))


def _front_end_worker(rank, world, n, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import beta_sgp_b200 as bs
    seen = {}

    def fake_solve(gn, psf, bkg, divergence="beta", flux=None, betaParam=1.005, x0=None, obj=None, **kw):
        seen["n"] = gn.shape[0]
        return _fake_result(bs, gn, psf, bkg, flux, betaParam, x0, obj, kw["MAXIT"])

    bs.engine.solve_batch = fake_solve
    a = _inputs(n)
    ok = True
    for variant in range(3):
        kw = dict(a)
        if variant == 1:                                   # shared PSF, scalar background, no x0 / obj
            kw.update(psf=a["psf"][0], bkg=np.float64(2.5), x0=None, obj=None)
        if variant == 2:                                   # per-image scalar background
            kw.update(bkg=a["bkg"][:, 0, 0].copy())
        full = bs.solve_batch_sharded(kw.pop("gn"), kw.pop("psf"), kw.pop("bkg"), MAXIT=5, width=None, **kw)
        args = dict(a) if variant == 0 else (dict(a, psf=a["psf"][0], bkg=np.float64(2.5), x0=None, obj=None) if variant == 1 else dict(a, bkg=a["bkg"][:, 0, 0]))
        want = _fake_result(bs, args["gn"], args["psf"], args["bkg"], args["flux"], args["betaParam"], args["x0"], args["obj"], 5)
        for k in ("x", "iters", "status", "discr", "times", "beta_final", "proj_evals", "ls_trials"):
            ok = ok and np.array_equal(full[k], getattr(want, k)) and full[k].dtype == getattr(want, k).dtype
        ok = ok and seen.get("n", 0) == bs.shard.shard_counts(n, world, snake=True)[rank]
        seen.clear()
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 11), (3, 7), (3, 2)])
def test_sharded_front_end_over_gloo(world, n):
    """solve_batch_sharded itself on `world` CPU processes (gloo) with the CUDA solve replaced by a stand-in that depends
    on every per-image argument: slicing of gn / psf / bkg / flux / betaParam / x0 / obj, cost-ranked dealing, gather
    and re-ordering, empty shards (n < world: the rank enters the collective with zero rows)."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 31500 + (os.getpid() + world * 17 + n) % 2000
    procs = [ctx.Process(target=_front_end_worker, args=(r, world, n, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))


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
