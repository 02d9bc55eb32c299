"""Drop-in for the reference's ``restoration/flux_conserve_proj.py``: ``projectDF`` on the GPU.

Same signature and return value as flux_conserve_proj.py:7,144.  The bracketing + safeguarded-secant
state machine runs on the device (csrc/bsgp_project.cuh); one full pass over ``c``/``dia`` and one
block reduction per evaluation.
"""
from __future__ import annotations

import numpy as np

try:
    from . import engine
except ImportError:
    import sgp as _sgp_loader  # noqa: F401  (registers the package when this directory is on sys.path)
    from beta_sgp_b200 import engine

EPSILON = np.finfo(float).eps
DEVICE = 0


def projectDF(b, c, dia, scaling, ccd_sat_level=None, lambda_=0, dlambda_=1, tol_lam=1e-11, biter=0, siter=0,
              max_projs=1000):
    """
    Equation: min 0.5 * x' * diag(dia) * x - c' * x
                subj to sum(x) = b
                x >= 0 (and x <= ccd_sat_level/scaling - eps if a saturation level is given)
    """
    c = np.asarray(c).astype(np.float64, copy=False)
    dia = np.asarray(dia).astype(np.float64, copy=False)
    b = float(np.asarray(b).astype(np.float64, copy=False))
    cap = None if ccd_sat_level is None else ccd_sat_level / scaling - EPSILON
    # biter / siter are the initial values of the reference's counters: biter grows during the bracketing, the secant
    # budget is max_projs - biter afterwards (:103) and the loop runs while siter < budget (:106); the device code keeps both
    x, _, st = engine.project_batch(b, c.ravel(), dia.ravel(), sat_cap=cap, lambda_=lambda_, dlambda_=dlambda_,
                                    tol_lam=tol_lam, max_projs=max_projs, biter=biter, siter=siter, device=DEVICE)
    if int(st[0]) != 0:
        raise RuntimeError("projectDF: the multiplier could not be bracketed (the reference loops forever here)")
    return x[0].reshape(c.shape)
