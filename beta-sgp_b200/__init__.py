"""beta-sgp_b200 — B200-native (sm_100a CUDA) SGP / beta-SGP image restoration.

Drop-in entry points (same names, arguments and returns as the reference's restoration/sgp.py and
restoration/flux_conserve_proj.py) plus batched variants.  The directory name carries a hyphen (repo
layout); import it as ``beta_sgp_b200`` (alias package at the repo root) or put this directory on
``sys.path`` and ``from sgp import sgp, sgp_betaDiv`` like the reference's scripts do.
"""
from . import _capi, engine, psf_model, shard, sweep, synth, tiles  # noqa: F401
from .sweep import sgp_betaDiv_sweep  # noqa: F401
from .engine import BatchResult, Plan, clear_plans, get_plan, project_batch, solve_batch  # noqa: F401
from .flux_conserve_proj import projectDF  # noqa: F401
from .shard import solve_batch_sharded  # noqa: F401
from .sgp import (DEFAULT_COLUMNS, DEFAULT_PARAMS, PsfOperator, betaDiv, betaDivDeriv, betaDivDerivwrtY,  # noqa: F401
                  lr_schedule, sgp, sgp_betaDiv)


def sgp_batch(gn, psf, bkg, **kw):
    """KL-SGP on a batch [B,ny,nx] (see engine.solve_batch)."""
    return solve_batch(gn, psf, bkg, divergence="kl", **kw)


def sgp_betaDiv_batch(gn, psf, bkg, **kw):
    """beta-SGP on a batch [B,ny,nx] (see engine.solve_batch)."""
    return solve_batch(gn, psf, bkg, divergence="beta", **kw)


__all__ = ["sgp", "sgp_betaDiv", "projectDF", "sgp_batch", "sgp_betaDiv_batch", "solve_batch", "solve_batch_sharded", "project_batch", "Plan",
           "get_plan", "clear_plans", "BatchResult", "PsfOperator", "betaDiv", "betaDivDeriv", "betaDivDerivwrtY",
           "lr_schedule", "DEFAULT_PARAMS", "DEFAULT_COLUMNS", "synth", "tiles", "psf_model", "sweep", "sgp_betaDiv_sweep"]
