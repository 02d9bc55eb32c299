"""The same one-GPU emulation of a W-GPU shard for the 8192-stamp workload: slowest virtual rank per CTA configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import beta_sgp_b200 as bs
from beta_sgp_b200 import shard, engine
dev = torch.device("cuda", 0)
B = 8192
w = bs.synth.star_stamps(B, 32, seed=12345); kw = dict(bs.synth.STAMP_KWARGS)
t = {k: torch.as_tensor(w[k], device=dev) for k in ("gn", "psf", "bkg", "flux")}
cands = [(0, 0), (1, 256)]
for W in [int(v) for v in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    cr = shard.expected_cost_rank(B, w["beta0"])
    res = {c: [] for c in cands}
    for rank in range(W):
        idx = shard.shard_indices(B, rank, W, cr)
        sel = torch.as_tensor(idx, device=dev, dtype=torch.long)
        gn, ps, bk, fl = (t[k].index_select(0, sel) for k in ("gn", "psf", "bkg", "flux"))
        for c in cands:
            plan = engine.get_plan(32, 32, "float64", 0, c[0], c[1])
            plan.set_psf(ps)
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                r = engine.solve_batch(gn, ps, bk, divergence="beta", flux=fl, betaParam=w["beta0"][idx], plan=plan, **kw)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[c].append(best)
    it = r.iters.cpu().numpy()
    print(f"W={W} stamps/rank={B // W} (last shard: max iterations {it.max()}): " + "  ".join(f"{c}: max {max(v):.2f} mean {np.mean(v):.2f} ms" for c, v in res.items()), flush=True)
