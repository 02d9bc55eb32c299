"""Drop-in for the reference's ``restoration/sgp.py`` entry points, backed by the CUDA library.

    from sgp import sgp, sgp_betaDiv, DEFAULT_PARAMS, DEFAULT_COLUMNS, betaDiv, betaDivDeriv, betaDivDerivwrtY

works with this directory on ``sys.path`` exactly as it does with the reference's ``restoration/``
(callers: application_sgp_star_stamps.py:5,22, simulation_test_sgp.py:3, tests.py:1).  Signatures,
argument meaning, return values, exceptions and side effects (``./sgp.log``, the two ``print`` lines
of ``sgp_betaDiv``) follow sgp.py:41-47 / 506-513 / 438 / 895.  All arithmetic runs on the GPU
(libbsgp.so); this file only validates arguments, moves arrays and reproduces the host-side side
effects after the solve.  ``use_original_SGP_Afunction=False`` selects the zero-padded operator with
astropy ``convolve_fft`` semantics (sgp.py:121-161; any image / kernel size; parity unpinned).
Not supported (raises NotImplementedError): ``save=True`` (per-iteration FITS dumps through astropy).
"""
from __future__ import annotations

import logging
import math

import numpy as np

try:                                    # imported as a package module (beta_sgp_b200.sgp) ...
    from . import _capi, engine
except ImportError:                     # ... or as top-level `sgp` with this directory on sys.path
    import importlib.util as _ilu
    import os as _os
    import sys as _sys

    def _load_pkg():
        here = _os.path.dirname(_os.path.abspath(__file__))
        name = "beta_sgp_b200"
        if name not in _sys.modules:
            spec = _ilu.spec_from_file_location(name, _os.path.join(here, "__init__.py"), submodule_search_locations=[here])
            mod = _ilu.module_from_spec(spec)
            _sys.modules[name] = mod
            spec.loader.exec_module(mod)
        return _sys.modules[name]

    _pkg = _load_pkg()
    from beta_sgp_b200 import _capi, engine  # noqa: E402

# sgp.py:34-39
DEFAULT_PARAMS = (1000, 1e-4, 0.4, 1e-5, 1e5, 1e1, 3, 0.5, 1)
DEFAULT_COLUMNS = ['label', 'xcentroid', 'ycentroid', 'sky_centroid',
                   'bbox_xmin', 'bbox_xmax', 'bbox_ymin', 'bbox_ymax',
                   'area', 'semimajor_sigma', 'semiminor_sigma',
                   'orientation', 'eccentricity', 'min_value', 'max_value',
                   'local_background', 'segment_flux', 'segment_fluxerr', 'ellipticity', 'fwhm']

DEVICE = 0   # CUDA device used by the single-image entry points


def _check_psf(psf):
    """sgp.py:98-102 / 558-562."""
    check = np.abs(np.sum(np.asarray(psf).flatten()) - 1.)
    tol = 1e4 * np.finfo(float).eps
    if check > tol:
        errmsg = f"\n\tsum(psf) - 1. = {check}, tolerance = {tol}"
        raise ValueError(f'PSF is not normalized! Provide a normalized PSF! {errmsg}')


def _status_error(status):
    if status in (_capi.ST_BAD_FLUX, _capi.ST_EMPTY_BOUNDS):
        # the reference dies in np.min of an empty selection at sgp.py:269 / 713
        return ValueError("zero-size array to reduction operation minimum which has no identity "
                          f"[{_capi.STATUS_TEXT[status]}]")
    return RuntimeError(_capi.STATUS_TEXT.get(status, f"solver status {status}"))


def _write_log(res, kw, tol):
    """The per-iteration INFO lines of sgp.py:291-298, 351-352, 399-411, written after the solve."""
    if not kw["verbose"]:
        return
    crit, maxit, iters = kw["stop_criterion"], kw["MAXIT"], int(res.iters[0])
    discr, stopv = res.discr[0], res.stop_value[0]
    if crit == 2:
        logging.info('it 0 || x_k - x_(k-1) ||^2 / || x_k ||^2 0 \n')
    elif crit == 3:
        logging.info('it 0 | f_k - f_(k-1) | / | f_k | 0 \n')
    elif crit == 4:
        logging.info(f'it 0 D_k {discr[0]} \n')
    M = kw["M"]
    for k in range(1, iters + 1):
        window = discr[max(0, k - M):k]
        if discr[k] >= window.max():                       # fv >= fr, in units of Discr_coeff
            logging.warning("\tWarning, fv >= fr")
        if crit == 1:
            logging.info(f'it {k} of  {maxit}\n')
        elif crit == 2:
            logging.info(f'it {k} || x_k - x_(k-1) ||^2 / || x_k ||^2 {stopv[k]} tol {tol}\n')
        elif crit == 3:
            logging.info(f'it {k} | f_k - f_(k-1) | / | f_k | {stopv[k]} tol {tol}\n')
        elif crit == 4:
            logging.info(f'it {k} D_k {discr[k]} tol {tol}\n')


def _solve_one(divergence, gn, psf, bkg, kw, flux, betaParam, obj, save, use_original_SGP_Afunction):
    _check_psf(psf)
    logging.basicConfig(filename='sgp.log', level=logging.INFO, force=True)       # sgp.py:104 / 564
    if save:
        raise NotImplementedError("save=True (per-iteration FITS dumps, sgp.py:223-231,416-422) is out of scope")
    gn = np.asarray(gn)
    psf = np.asarray(psf)
    if gn.ndim != 2:
        raise ValueError("gn must be a 2-D image")
    if use_original_SGP_Afunction and psf.shape != gn.shape:
        # np.reshape(x, psf.shape) in the reference's closure fails the same way (sgp.py:112)
        raise ValueError(f"cannot reshape array of size {gn.size} into shape {psf.shape}")
    if kw["errflag"] and obj is None:
        raise ValueError("errflag was set to True but no ground-truth was passed.")             # sgp.py:237-238
    x0 = None
    if kw["init_recon"] == 1:                                                                  # sgp.py:168-170
        np.random.seed(42)
        x0 = np.random.randn(*gn.shape)[None]
    if not hasattr(bkg, "flatten"):
        # the reference calls bkg.flatten() (sgp.py:182 / 636): a Python float or int background fails there
        raise AttributeError(f"'{type(bkg).__name__}' object has no attribute 'flatten'")
    bkg_a = np.asarray(bkg, dtype=np.float64)
    if bkg_a.size == gn.size:
        bkg_a = bkg_a.reshape((1,) + gn.shape)
    elif bkg_a.size == 1:
        bkg_a = bkg_a.reshape(1)
    else:
        raise ValueError(f"operands could not be broadcast together with shapes ({gn.size},) ({bkg_a.size},)")
    res = engine.solve_batch(np.asarray(gn, dtype=np.float64)[None], psf, bkg_a, divergence=divergence,
                             flux=None if flux is None else [float(flux)], betaParam=float(betaParam), x0=x0,
                             obj=None if (obj is None or not kw["errflag"]) else np.asarray(obj, dtype=np.float64)[None],
                             device=DEVICE, padded=not use_original_SGP_Afunction, **kw)
    status = int(res.status[0])
    if status != _capi.ST_OK:
        raise _status_error(status)
    if isinstance(flux, np.ndarray) and flux.dtype.kind == "f":
        # `flux /= scaling` (sgp.py:211 / 666) is an in-place division when the caller hands over an ndarray (the 0-d result
        # of np.sum is one): the caller's array holds the scaled flux afterwards.  Reproduced; numpy scalars are rebound.
        flux /= res.scalars[0, 0]
    iters = int(res.iters[0])
    _write_log(res, kw, float(res.scalars[0, 4]))
    err = None
    if kw["errflag"]:
        if iters >= kw["MAXIT"]:
            # the reference writes err[iter_] after the increment (sgp.py:394-396) into an array of
            # MAXIT+1 entries, so a run that reaches MAXIT dies with this IndexError
            raise IndexError(f"index {kw['MAXIT'] + 1} is out of bounds for axis 0 with size {kw['MAXIT'] + 1}")
        err = res.err[0, :iters + 1].copy()
    return res.x[0], iters, res.discr[0, :iters + 1].copy(), res.times[0, :iters + 1].copy(), err, float(res.beta_final[0])


def sgp(
    gn, psf, bkg, init_recon=0, proj_type=0, stop_criterion=0, MAXIT=500,
    gamma=1e-4, beta=0.4, alpha=1.3, alpha_min=1e-5, alpha_max=1e5, M_alpha=3,
    tau=0.5, M=1, max_projs=1000, save=False, obj=None, verbose=True, flux=None,
    ccd_sat_level=None, scale_data=True, errflag=False, tol_convergence=1e-4,
    use_original_SGP_Afunction=True
):
    """KL-divergence scaled gradient projection (sgp.py:41-438), same arguments and returns:
    ``(x, iters, discr, times, err_or_None)``."""
    kw = dict(init_recon=init_recon, proj_type=proj_type, stop_criterion=stop_criterion, MAXIT=MAXIT, gamma=gamma,
              beta=beta, alpha=alpha, alpha_min=alpha_min, alpha_max=alpha_max, M_alpha=M_alpha, tau=tau, M=M,
              max_projs=max_projs, verbose=verbose, ccd_sat_level=ccd_sat_level, scale_data=scale_data,
              errflag=errflag, tol_convergence=tol_convergence)
    x, iters, discr, times, err, _ = _solve_one("kl", gn, psf, bkg, kw, flux, 1.0, obj, save, use_original_SGP_Afunction)
    return x, iters, discr, times, err if errflag else None


def sgp_betaDiv(
    gn, psf, bkg, init_recon=0, proj_type=0, stop_criterion=0, MAXIT=500,
    gamma=1e-4, beta=0.4, alpha=1.3, alpha_min=1e-5, alpha_max=1e5, M_alpha=3,
    tau=0.5, M=1, max_projs=1000, save=False, obj=None, verbose=True, flux=None,
    ccd_sat_level=None, scale_data=True, errflag=False, adapt_beta=True,
    betaParam=1.005, lr=1e-3, lr_exp_param=0.1, schedule_lr=False, tol_convergence=1e-4,
    use_original_SGP_Afunction=True
):
    """beta-divergence SGP (sgp.py:506-895), same arguments and returns: ``(x, iters, discr, times, None)``.
    ``errflag`` / ``obj`` are accepted and ignored like in the reference (sgp.py:514)."""
    kw = dict(init_recon=init_recon, proj_type=proj_type, stop_criterion=stop_criterion, MAXIT=MAXIT, gamma=gamma,
              beta=beta, alpha=alpha, alpha_min=alpha_min, alpha_max=alpha_max, M_alpha=M_alpha, tau=tau, M=M,
              max_projs=max_projs, verbose=verbose, ccd_sat_level=ccd_sat_level, scale_data=scale_data,
              errflag=False, tol_convergence=tol_convergence, adapt_beta=adapt_beta, lr=lr, lr_exp_param=lr_exp_param,
              schedule_lr=schedule_lr)
    x, iters, discr, times, _, beta_final = _solve_one("beta", gn, psf, bkg, kw, flux, betaParam, None, save,
                                                       use_original_SGP_Afunction)
    print(f'Beta parameter in beta-divergence (final value): {beta_final}')          # sgp.py:892-893
    print(f'No. of iterations: {iters}')
    return x, iters, discr, times, None


# ---------------------------------------------------------------------------------------------
# helpers exported by the reference module (sgp.py:441-503)
# ---------------------------------------------------------------------------------------------
def betaDiv(y, x, betaParam):
    """beta-divergence of data ``x`` from model ``y`` (sgp.py:441-458), computed on the GPU."""
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.zeros(1)
    _capi.check(_capi.lib().bsgp_beta_div_host(y.ctypes.data, x.ctypes.data, y.size, float(betaParam), out.ctypes.data,
                                               None, DEVICE))
    return out[0]


def betaDivDeriv(y, x, betaParam):
    """Per-element derivative of the beta-divergence w.r.t. beta (sgp.py:462-495); 0 for beta in {0, 1}."""
    if betaParam == 0 or betaParam == 1:
        return 0
    shape = np.shape(y)
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.zeros(1)
    d = np.empty_like(y)
    _capi.check(_capi.lib().bsgp_beta_div_host(y.ctypes.data, x.ctypes.data, y.size, float(betaParam), out.ctypes.data,
                                               d.ctypes.data, DEVICE))
    return d.reshape(shape)


def betaDivDerivwrtY(AT, den_arg, gn_arg, betaParam):
    """den^(beta-1) - AT(x = gn * den^(beta-2)) (sgp.py:498-499); the two powers are evaluated on the GPU,
    ``AT`` is the caller's operator (e.g. ``PsfOperator.AT``)."""
    den = np.ascontiguousarray(den_arg, dtype=np.float64).ravel()
    gnv = np.ascontiguousarray(gn_arg, dtype=np.float64).ravel()
    p1 = np.empty_like(den)
    u = np.empty_like(den)
    _capi.check(_capi.lib().bsgp_beta_grad_terms_host(den.ctypes.data, gnv.ctypes.data, den.size, float(betaParam),
                                                      p1.ctypes.data, u.ctypes.data, DEVICE))
    return p1 - AT(x=u)


def lr_schedule(init_lr, k, epoch):
    """sgp.py:502-503."""
    return init_lr * math.exp(-k * epoch)


class PsfOperator:
    """The reference's ``A`` / ``AT`` closures (sgp.py:108-120) on the GPU: flattened vector in, flattened
    vector out, circular convolution with TF = fftn(fftshift(psf)) or its conjugate."""

    def __init__(self, psf, device=None):
        psf = np.asarray(psf, dtype=np.float64)
        self.shape = psf.shape
        self.plan = engine.Plan(psf.shape[0], psf.shape[1], "float64", DEVICE if device is None else device)
        self.plan.set_psf(psf)

    def A(self, x):
        return self.plan.apply_psf(np.reshape(x, self.shape), adjoint=False).ravel()

    def AT(self, x):
        return self.plan.apply_psf(np.reshape(x, self.shape), adjoint=True).ravel()
