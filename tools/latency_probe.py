"""Solve-kernel time as a function of batch size and CTA configuration (tiles256 workload): the data behind the
adaptive-width rule of the sharded path.  Usage: python tools/latency_probe.py [tiles|stamps]"""
import json
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import beta_sgp_b200 as bs

what = sys.argv[1] if len(sys.argv) > 1 else "tiles"
dev = torch.device("cuda", 0)
if what == "tiles":
    w = bs.synth.field_tiles(size=2048, tile=256, seed=2024, n_beta=5); kw = dict(bs.synth.TILE_KWARGS); shared = True
    sizes = [1, 8, 20, 40, 80, 160, 320]; configs = [(0, 128), (0, 256), (0, 512), (16, 128), (16, 256)]
else:
    w = bs.synth.star_stamps(8192, 32, seed=12345); kw = dict(bs.synth.STAMP_KWARGS); shared = False
    sizes = [1, 64, 296, 1024, 8192]; configs = [(0, 256), (0, 128), (0, 512)]
order = np.argsort(np.abs(w["beta0"] - 1.0), kind="stable")
rows = []
for (G, th) in configs:
    try:
        plan = bs.Plan(w["gn"].shape[1], w["gn"].shape[2], "float64", 0, cluster_size=G, threads=th)
    except Exception as e:
        print("config", G, th, "failed:", e); continue
    info = plan.info()
    for B in sizes:
        idx = np.sort(order[:: max(1, len(order) // B)][:B])          # spread over the beta groups like a round-robin shard
        t = {k: torch.as_tensor(np.ascontiguousarray(w[k][idx] if (k != "psf" or not shared) else w[k]), device=dev) for k in ("gn", "psf", "bkg", "flux", "beta0")}
        plan.set_psf(t["psf"])
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = bs.solve_batch(t["gn"], None, t["bkg"], divergence="beta", flux=t["flux"], betaParam=t["beta0"], plan=plan, psf_is_set=True, **kw)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        it = r.iters.cpu().numpy()
        rows.append(dict(G=info["cluster_size"], threads=info["threads"], slots=info["num_clusters"], B=B, ms=ms, sum_it=int(it.sum()), max_it=int(it.max()),
                         us_per_it_longest=1e3 * ms / it.max(), us_per_img_it=1e3 * ms / it.sum()))
        print(json.dumps(rows[-1]), flush=True)
    plan.close()
