"""Which CTA configuration for a shard of the 320-tile field?  One GPU plays every rank of a world of W in turn: the shard of
each virtual rank (shard.shard_indices, the dealing bench.py --gpus W uses) is solved under each candidate configuration and
the slowest rank is reported, which is what a W-GPU step waits for (plus the gather)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import beta_sgp_b200 as bs
from beta_sgp_b200 import shard, engine
dev = torch.device("cuda", 0)
w = bs.synth.field_tiles(size=2048, tile=256, seed=2024, n_beta=5); kw = dict(bs.synth.TILE_KWARGS)
t = {k: torch.as_tensor(w[k], device=dev) for k in ("gn", "psf", "bkg", "flux")}
B = 320
cands = [(0, 0), (16, 128), (0, 256), (16, 256)]
for W in [int(v) for v in (sys.argv[1:] or ["2", "4", "8"])]:
    cr = shard.expected_cost_rank(B, w["beta0"])
    res = {c: [] for c in cands}
    for rank in range(W):
        idx = shard.shard_indices(B, rank, W, cr)
        sel = torch.as_tensor(idx, device=dev, dtype=torch.long)
        gn, bk, fl = t["gn"].index_select(0, sel), t["bkg"].index_select(0, sel), t["flux"].index_select(0, sel)
        for c in cands:
            plan = engine.get_plan(256, 256, "float64", 0, c[0], c[1])
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                r = engine.solve_batch(gn, t["psf"], bk, divergence="beta", flux=fl, betaParam=w["beta0"][idx], plan=plan, **kw)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[c].append(best)
    auto = engine.auto_config(256, 256, len(shard.shard_indices(B, 0, W, cr)))
    print(f"W={W} images/rank={B // W} auto={auto}: " + "  ".join(f"{c}: max {max(v):.2f} mean {np.mean(v):.2f} ms" for c, v in res.items()), flush=True)
