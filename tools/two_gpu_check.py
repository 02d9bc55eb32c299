"""torchrun --nproc-per-node 2 tools/two_gpu_check.py : the sharded front end on two GPUs equals the single-GPU batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import beta_sgp_b200 as bs

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = bs.synth.star_stamps(37, 32, seed=5)
kw = dict(bs.synth.STAMP_KWARGS)
full = bs.solve_batch_sharded(st["gn"], st["psf"], st["bkg"], flux=st["flux"], betaParam=st["beta0"], divergence="beta", **kw)
if dist.get_rank() == 0:
    ref = bs.sgp_betaDiv_batch(st["gn"], st["psf"], st["bkg"], flux=st["flux"], betaParam=st["beta0"], device=local, **kw)
    assert np.array_equal(full["x"], ref.x) and np.array_equal(full["iters"], ref.iters)
    print("two-GPU sharded solve == single-GPU batch for", len(st["gn"]), "stamps; iters", full["iters"][:8])
dist.barrier()
dist.destroy_process_group()
