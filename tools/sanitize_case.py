"""Small solves for compute-sanitizer runs: a few stamps (one CTA each), one 64x64 image in a cluster of 2,
one 256x256 tile in a cluster of 8 (few iterations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import beta_sgp_b200 as bs
st = bs.synth.star_stamps(4, 32, seed=3)
kw = dict(bs.synth.STAMP_KWARGS, MAXIT=6)
r = bs.sgp_betaDiv_batch(st["gn"], st["psf"], st["bkg"], flux=st["flux"], betaParam=st["beta0"], **kw)
print("stamps", r.iters, r.status)
rng = np.random.default_rng(0)
psf = bs.synth.moffat_psf(64, 64, 3.0)
gn = rng.poisson(50.0 + 500.0 * rng.random((2, 64, 64))).astype(float)
plan = bs.Plan(64, 64, cluster_size=2)
r = bs.solve_batch(gn, psf, np.float64(50.0), divergence="kl", init_recon=3, stop_criterion=1, MAXIT=4, plan=plan)
print("kl 64 cluster 2", r.iters, r.status)
t = bs.synth.field_tiles(size=512, tile=256, seed=1, n_beta=2, max_tiles=1)
kw = dict(bs.synth.TILE_KWARGS, MAXIT=3)
r = bs.sgp_betaDiv_batch(t["gn"], t["psf"], t["bkg"], flux=t["flux"], betaParam=t["beta0"], **kw)
print("tiles", r.iters, r.status)
# pipelined host path (ready flags, zero-copy output), tiling kernels, PSF model
import torch
r2 = bs.sgp_betaDiv_batch(torch.as_tensor(t["gn"]).pin_memory(), t["psf"], torch.as_tensor(t["bkg"]).pin_memory(), flux=t["flux"], betaParam=t["beta0"], **kw)
print("tiles, pinned path", r2.iters, r2.status, bool(np.array_equal(r2.x.numpy(), r.x)))
frame = np.random.default_rng(2).random((300, 200))
tl, org = bs.tiles.create_subdivisions(frame, (128, 64), 9)
back = bs.tiles.reconstruct_full_image_from_patches(tl, org, frame.shape)
print("tiling", tuple(tl.shape), float((back.cpu() - torch.as_tensor(frame)).abs().max()))
pm = bs.psf_model.evaluate_batch(np.array([[0.0084, 0.99996, -0.1267, -0.1662, 0.548] + [0.1] + [0.0] * 11]), 2, 15, (32, 32))
print("psf model", float(pm.sum()))
