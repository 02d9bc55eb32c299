"""Where does the time of one sharded step go?  Run under torchrun (N >= 1): stages of solve_batch_sharded timed with CUDA events."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import beta_sgp_b200 as bs
from beta_sgp_b200 import shard, engine
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local); dev = torch.device("cuda", local)
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
w = bs.synth.field_tiles(size=2048, tile=256, seed=2024, n_beta=5); kw = dict(bs.synth.TILE_KWARGS)
t = {k: torch.as_tensor(w[k], device=dev) for k in ("gn", "psf", "bkg", "flux")}
B = 320
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(4):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    h0 = time.perf_counter(); e = [ev()]
    cr = shard.expected_cost_rank(B, w["beta0"]); idx = shard.shard_indices(B, rank, world, cr)
    sel = torch.as_tensor(idx, device=dev, dtype=torch.long)
    gn, bk, fl = t["gn"].index_select(0, sel), t["bkg"].index_select(0, sel), t["flux"].index_select(0, sel)
    e.append(ev()); h1 = time.perf_counter()
    cs, th = engine.auto_config(256, 256, len(idx))
    plan = engine.get_plan(256, 256, "float64", local, cs, th)
    r = engine.solve_batch(gn, t["psf"], bk, divergence="beta", flux=fl, betaParam=w["beta0"][idx], plan=plan, **kw)
    e.append(ev()); h2 = time.perf_counter()
    loc = {k: getattr(r, k) for k in shard._FIELDS}
    full = shard.gather_to_all(loc, B, rank, world, cost_rank=cr) if world > 1 else loc
    e.append(ev()); h3 = time.perf_counter()
    torch.cuda.synchronize(); h4 = time.perf_counter()
    if rank == 0:
        print(f"rep {rep} device ms: slice {e[0].elapsed_time(e[1]):.2f} solve(+psf,alloc) {e[1].elapsed_time(e[2]):.2f} gather {e[2].elapsed_time(e[3]):.2f} total {e[0].elapsed_time(e[3]):.2f} | "
              f"host ms: slice {1e3*(h1-h0):.2f} solve-call {1e3*(h2-h1):.2f} gather-call {1e3*(h3-h2):.2f} sync {1e3*(h4-h3):.2f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
