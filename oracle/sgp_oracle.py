"""CPU oracle for the SGP / beta-SGP restoration loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``beta-sgp_b200/`` may import this
module: it is the checker for the CUDA path (tests/, ``__graft_entry__.smoke``)
and the timed CPU arm of ``bench.py`` (``cpu_baseline`` / ``--impl reference``).

It is a numpy restatement of the reference algorithm, organised as one solver
class instead of the reference's two ~400-line functions, and it records a
per-iteration trace (step length, line-search trials, projection evaluations,
divergence parameter) that the reference does not expose.  Every arithmetic
expression keeps the reference's operand order so that, under the same numpy,
results are bit-identical to the unmodified reference; that is the pin:
``tests/golden/make_golden.py`` runs the unmodified reference (imported from
/root/reference through stub modules) next to this oracle and
``tests/test_oracle_pinned.py`` asserts equality against the committed vectors.

Reference citations (relative to /root/reference/restoration):
  sgp.py:41-438      sgp()            -> Solver(divergence="kl")
  sgp.py:506-895     sgp_betaDiv()    -> Solver(divergence="beta")
  sgp.py:441-503     betaDiv / betaDivDeriv / betaDivDerivwrtY / lr_schedule
  flux_conserve_proj.py:7-144  projectDF() -> flux_projection()
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from timeit import default_timer

import numpy as np

MACHINE_EPS = np.finfo(float).eps


# --------------------------------------------------------------------------
# beta-divergence pieces (sgp.py:441-503)
# --------------------------------------------------------------------------
def beta_divergence(model, data, b):
    """D_b(data | model); sgp.py:441-458 (argument order there is (y=model, x=data))."""
    if b == 0:
        return np.sum(data / model) - np.sum(np.log(data / model)) - data.size
    if b == 1:
        return np.sum(np.multiply(data, np.log(np.divide(data, model)))) - np.sum(data) + np.sum(model)
    k = 1 / (b * (b - 1))
    return np.sum(k * data ** b) + np.sum(k * (b - 1) * model ** b) - np.sum(k * b * data * model ** (b - 1))


def beta_divergence_dbeta(model, data, b):
    """Per-pixel d D_b / d b; sgp.py:462-495 (returns the scalar 0 for b in {0, 1})."""
    if b == 0 or b == 1:
        return 0
    y, x = model, data
    return (-x * y ** (b - 1) * np.log(y) / (b - 1) + x * y ** (b - 1) / (b - 1) ** 2
            + x ** b * np.log(x) / (b * (b - 1)) - x ** b / (b * (b - 1) ** 2)
            + y ** b * np.log(y) / b - x ** b / (b ** 2 * (b - 1)) - y ** b / b ** 2)


def beta_divergence_grad(adjoint, model, data, b):
    """Gradient w.r.t. the image through A^T; sgp.py:498-499."""
    return model ** (b - 1) - adjoint(data * model ** (b - 2))


def learning_rate(lr0, k, epoch):
    """sgp.py:502-503."""
    return lr0 * math.exp(-k * epoch)


# --------------------------------------------------------------------------
# PSF operator, numpy circular convolution (sgp.py:108-120 / 570-582)
# --------------------------------------------------------------------------
class CircularPsf:
    def __init__(self, psf):
        self.shape = psf.shape
        self.tf = np.fft.fftn(np.fft.fftshift(psf))
        self.ctf = np.conj(self.tf)

    def _apply(self, vec, spectrum):
        img = np.reshape(vec, self.shape)
        out = np.real(np.fft.ifftn(np.multiply(spectrum, np.fft.fftn(img))))
        return out.flatten()

    def forward(self, vec):
        return self._apply(vec, self.tf)

    def adjoint(self, vec):
        return self._apply(vec, self.ctf)


# --------------------------------------------------------------------------
# PSF operator, zero-padded FFT convolution (sgp.py:121-161 / 583-615)
# --------------------------------------------------------------------------
class PaddedPsf:
    """``use_original_SGP_Afunction=False``: A(x) = convolve_fft(x, psf, normalize_kernel=True,
    normalization_zero_tol=1e-4), A^T(x) = convolve_fft(x, psf.conj().T, ...) (sgp.py:138,157).

    PARITY UNPINNED.  ``astropy.convolution.convolve_fft`` is a third-party function that is neither vendored
    in the reference nor installed here (the reference pins no version), and no reference test exercises this
    branch.  This class restates the function's published algorithm for the arguments the reference passes and
    the remaining defaults (boundary='fill', fill_value=0.0, nan_treatment='interpolate', crop=True, and hence
    psf_pad = fft_pad = True): kernel divided by its sum; image and kernel placed in the centre of a square
    2^ceil(log2(max(image + kernel shape))) array (zeros elsewhere); kernel spectrum taken after ifftshift;
    product of spectra; the interpolation weights ifftn(fftn(ones) * kernfft) divide the result inside the
    image window; weights below 10 eps zero the result; the image window is cropped out."""

    def __init__(self, psf, shape):
        self.shape = tuple(shape)
        self.fwd = self._prepare(np.asarray(psf))
        self.adj = self._prepare(np.asarray(psf).conj().T)

    def _prepare(self, kernel):
        ksum = kernel.sum()
        if abs(ksum) < 1e-4:
            raise ValueError("The kernel can't be normalized, because its sum is close to zero.")
        kern = kernel / ksum
        fsize = int(2 ** np.ceil(np.log2(np.max(np.array(self.shape) + np.array(kern.shape)))))
        newshape = (fsize, fsize)
        aslc, kslc = [], []
        for n, na, nk in zip(newshape, self.shape, kern.shape):
            centre = n - (n + 1) // 2
            aslc.append(slice(centre - na // 2, centre + (na + 1) // 2))
            kslc.append(slice(centre - nk // 2, centre + (nk + 1) // 2))
        aslc, kslc = tuple(aslc), tuple(kslc)
        bigkernel = np.zeros(newshape, dtype=complex)
        bigkernel[kslc] = kern
        kernfft = np.fft.fftn(np.fft.ifftshift(bigkernel))
        bigimwt = np.ones(newshape, dtype=complex)
        wtsm = np.fft.ifftn(np.fft.fftn(bigimwt) * kernfft)
        bigimwt[aslc] = wtsm.real[aslc]
        return newshape, aslc, kernfft, bigimwt

    def _apply(self, vec, prep):
        newshape, aslc, kernfft, bigimwt = prep
        big = np.zeros(newshape, dtype=complex)
        big[aslc] = np.reshape(vec, self.shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            rifft = np.fft.ifftn(np.fft.fftn(big) * kernfft) / bigimwt
        rifft[bigimwt < 10 * np.finfo(bigimwt.dtype).eps] = 0.0
        return rifft[aslc].real.ravel()

    def forward(self, vec):
        return self._apply(vec, self.fwd)

    def adjoint(self, vec):
        return self._apply(vec, self.adj)


# --------------------------------------------------------------------------
# flux-conserving projection (flux_conserve_proj.py:7-144)
# --------------------------------------------------------------------------
class _Clamp:
    """x(lam) = min(cap, max(0, (c+lam)/dia)) and its residual; flux_conserve_proj.py:22-25."""

    def __init__(self, c, dia, target, cap):
        self.c, self.dia, self.target, self.cap = c, dia, target, cap
        self.evals = 0

    def point(self, lam):
        x = np.maximum(0, np.divide(self.c + lam, self.dia))
        if self.cap is not None:
            x = np.minimum(self.cap, x)
        return x

    def residual(self, lam):
        self.evals += 1
        x = self.point(lam)
        return x, np.sum(x) - self.target


def flux_projection(b, c, dia, scaling, ccd_sat_level=None, lambda_=0, dlambda_=1,
                    tol_lam=1e-11, biter=0, siter=0, max_projs=1000, counter=None):
    """min 1/2 x'diag(dia)x - c'x  s.t. sum(x)=b, 0<=x(<=sat): bracketing + safeguarded secant on
    the multiplier.  Restates flux_conserve_proj.py:7-144 including the `x = ...` slip at line 122
    (the secant ratio ``s`` is *not* refreshed in that branch).  ``counter`` (a list) receives the
    number of full-image evaluations."""
    c = np.asarray(c).astype(np.float64, copy=False)
    dia = np.asarray(dia).astype(np.float64, copy=False)
    b = np.asarray(b).astype(np.float64, copy=False)[()]
    cap = None if ccd_sat_level is None else ccd_sat_level / scaling - MACHINE_EPS
    f = _Clamp(c, dia, b, cap)
    tol_r = 1e-11 * b
    lam, dlam = lambda_, dlambda_

    def done(x):
        if counter is not None:
            counter.append(f.evals)
        return x

    x, r = f.residual(lam)                                   # :22-25
    if abs(r) < tol_r:                                       # :27-28
        return done(x)

    if r < 0:                                                # :30-54 bracket upwards
        lam_lo, r_lo = lam, r
        lam = lam + dlam
        x, r = f.residual(lam)
        while r < 0:
            biter += 1
            lam_lo = lam
            s = np.max([r_lo / r - 1, 0.1])
            dlam = dlam + dlam / s
            lam = lam + dlam
            r_lo = r
            x, r = f.residual(lam)
        lam_hi, r_hi = lam, r
    else:                                                    # :55-81 bracket downwards
        lam_hi, r_hi = lam, r
        lam = lam - dlam
        x, r = f.residual(lam)
        while r > 0:
            biter += 1
            lam_hi = lam
            s = np.max([r_hi / r - 1, 0.1])
            try:                                             # :68-72 FP-exception escape
                with np.errstate(all='raise'):
                    dlam = dlam + dlam / s
            except Exception:
                break
            lam = lam - dlam
            r_hi = r
            x, r = f.residual(lam)
        lam_lo, r_lo = lam, r

    if abs(r_hi) < tol_r:                                    # :84-93 end-point checks
        return done(f.point(lam_hi))
    if abs(r_lo) < tol_r:
        return done(f.point(lam_lo))

    s = 1 - r_lo / r_hi                                      # :96-103 secant start
    dlam = dlam / s
    lam = lam_hi - dlam
    x, r = f.residual(lam)
    budget = max_projs - biter

    while abs(r) > tol_r and dlam > tol_lam * (1 + abs(lam)) and siter < budget:   # :106-142
        siter += 1
        if r > 0:
            if s <= 2:
                lam_hi, r_hi = lam, r
                s = 1 - r_lo / r_hi
                dlam = (lam_hi - lam_lo) / s
                lam = lam_hi - dlam
            else:
                s = np.max([r_hi / r - 1, 0.1])
                dlam = (lam_hi - lam) / s
                lam_new = np.max([lam - dlam, 0.75 * lam_lo + 0.25 * lam])
                lam_hi, r_hi = lam, r
                lam = lam_new
                # reference line 122 assigns this ratio to `x`, not `s`: s keeps its value
        else:
            if s >= 2:
                lam_lo, r_lo = lam, r
                s = 1 - r_lo / r_hi
                dlam = (lam_hi - lam_lo) / s
                lam = lam_hi - dlam
            else:
                s = np.max([r_lo / r - 1, 0.1])
                dlam = (lam - lam_lo) / s
                lam_new = np.min([lam + dlam, 0.75 * lam_hi + 0.25 * lam])
                lam_lo, r_lo = lam, r
                lam = lam_new
                s = (lam_hi - lam_lo) / (lam_hi - lam)
        x, r = f.residual(lam)
    return done(x)


# --------------------------------------------------------------------------
# solver
# --------------------------------------------------------------------------
@dataclass
class Trace:
    """Per-iteration controller record (index k = iteration k, 1-based like the reference log)."""
    alpha: list = field(default_factory=list)        # step length chosen for the NEXT iteration
    lam: list = field(default_factory=list)          # accepted line-search step
    fv: list = field(default_factory=list)           # objective after the accepted step
    beta_param: list = field(default_factory=list)   # divergence parameter after the iteration
    trials: list = field(default_factory=list)       # line-search evaluations in the iteration
    proj_evals: list = field(default_factory=list)   # full-image projection evaluations
    stop_value: list = field(default_factory=list)   # quantity compared with tol (criteria 2-4)
    init_proj_evals: int = 0
    x_low: float = float("nan")
    x_upp: float = float("nan")
    flux: float = float("nan")
    scaling: float = float("nan")


@dataclass
class OracleResult:
    x: np.ndarray
    iters: int
    discr: np.ndarray
    times: np.ndarray
    err: np.ndarray | None
    beta_param: float
    trace: Trace


def solve(gn, psf, bkg, divergence="kl", init_recon=0, proj_type=0, stop_criterion=0, MAXIT=500,
          gamma=1e-4, beta=0.4, alpha=1.3, alpha_min=1e-5, alpha_max=1e5, M_alpha=3, tau=0.5, M=1,
          max_projs=1000, obj=None, verbose=True, flux=None, ccd_sat_level=None, scale_data=True,
          errflag=False, adapt_beta=True, betaParam=1.005, lr=1e-3, lr_exp_param=0.1,
          schedule_lr=False, tol_convergence=1e-4, use_original_SGP_Afunction=True):
    """One restoration.  ``divergence="kl"`` follows sgp.py:41-438, ``"beta"`` sgp.py:506-895
    (numpy A/A^T closure only).  Returns an OracleResult; ``(x, iters, discr, times, err)`` are
    what the reference returns."""
    is_beta = divergence == "beta"
    if abs(np.sum(psf.flatten()) - 1.) > 1e4 * MACHINE_EPS:              # :98-102
        raise ValueError("PSF is not normalized! Provide a normalized PSF!")
    shape = gn.shape
    op = CircularPsf(psf) if use_original_SGP_Afunction else PaddedPsf(psf, shape)
    tr = Trace()
    lr0 = lr
    t0 = default_timer()

    if init_recon == 0:                                                   # :166-177
        x = np.zeros_like(gn)
    elif init_recon == 1:
        np.random.seed(42)
        x = np.random.randn(*gn.shape)
    elif init_recon == 2:
        x = gn.copy()
    elif init_recon == 3:
        x = (np.sum(gn - bkg) if flux is None else flux) / gn.size * np.ones_like(gn)

    gn = gn.flatten()
    x = x.flatten()
    bkg = np.asarray(bkg).flatten()

    tol = None                                                            # :185-190
    if stop_criterion in (2, 3):
        tol = tol_convergence
    elif stop_criterion == 4:
        tol = 1 + 1 / np.mean(gn)

    if scale_data:                                                        # :193-199
        scaling = np.max(gn)
        gn = gn / scaling
        bkg = bkg / scaling
        x = x / scaling
    else:
        scaling = 1.

    floor = np.min(gn[gn > 0])                                            # :202-204
    gn[gn <= 0] = floor * MACHINE_EPS * MACHINE_EPS

    npix = gn.size
    flux = np.sum(gn - bkg) if flux is None else flux / scaling           # :208-211
    tr.flux, tr.scaling = float(flux), float(scaling)

    iter_ = 1
    alpha_hist = alpha_max * np.ones(M_alpha)
    f_hist = -1e30 * np.ones(M)
    discr_coeff = 2 / npix * scaling
    ones = np.ones(npix)
    discr = np.zeros(MAXIT + 1)
    times = np.zeros(MAXIT + 1)

    if errflag and obj is None:
        raise ValueError("errflag was set to True but no ground-truth was passed.")
    if errflag:
        err = np.zeros(MAXIT + 1)
        truth = obj.flatten() / scaling
        truth_sq = np.sum(truth * truth)

    cnt = []
    if proj_type == 0:                                                    # :248-253
        x[x < 0] = 0
    elif proj_type == 1:
        x = flux_projection(flux, x, np.ones_like(x), scaling, ccd_sat_level=ccd_sat_level,
                            max_projs=max_projs, counter=cnt)
        tr.init_proj_evals = cnt[-1]
    if errflag:
        e = x - truth
        err[0] = np.sqrt(np.sum(e * e) / truth_sq)

    def objective(model, x_blur, b):
        if is_beta:
            return beta_divergence(model, gn, b)
        ratio = np.divide(gn, model)
        return np.sum(np.multiply(gn, np.log(ratio))) + np.sum(x_blur) - flux

    def gradient(model, b):
        if is_beta:
            return beta_divergence_grad(op.adjoint, model, gn, b)
        return ones - op.adjoint(np.divide(gn, model))

    x_tf = op.forward(x)                                                  # :260-265
    den = x_tf + bkg
    g = gradient(den, betaParam)
    fv = objective(den, x_tf, betaParam)

    y = np.multiply((flux / (flux + bkg)), op.adjoint(gn))                # :268-273
    x_low = np.min(y[y > 0])
    x_upp = np.max(y)
    if x_upp / x_low < 50:
        x_low = x_low / 10
        x_upp = x_upp * 10
    tr.x_low, tr.x_upp = float(x_low), float(x_upp)

    discr[0] = discr_coeff * fv

    def bounded(v):
        s = v.copy()
        s[s < x_low] = x_low
        s[s > x_upp] = x_upp
        return s

    X = np.ones_like(x) if init_recon == 0 else bounded(x)               # :279-288
    if proj_type == 1:
        D = np.divide(1, X)
    if verbose and stop_criterion == 2:                                   # :291-298
        tol = tol * tol

    keep_going = True
    epoch = 0
    while keep_going:
        epoch += 1
        x_before = x.copy()
        alpha_hist[0:M_alpha - 1] = alpha_hist[1:M_alpha]
        f_hist[0:M - 1] = f_hist[1:M]
        f_hist[M - 1] = fv

        y = x - alpha * np.multiply(X, g)                                 # :311-318
        n_eval = 0
        if proj_type == 0:
            y[y < 0] = 0
        elif proj_type == 1:
            y = flux_projection(flux, np.multiply(y, D), D, scaling, ccd_sat_level=ccd_sat_level,
                                max_projs=max_projs, counter=cnt)
            n_eval = cnt[-1]
        d = y - x

        gd = np.dot(d, g)                                                 # :321-326
        lam = 1
        d_tf = op.forward(d)
        f_ref = max(f_hist)
        n_trials = 0
        while True:                                                       # :328-349
            n_trials += 1
            x_try = x + lam * d
            x_tf_try = x_tf + lam * d_tf
            den = x_tf_try + bkg
            fv = objective(den, x_tf_try, betaParam)
            if fv <= f_ref + gamma * lam * gd or lam < 1e-12:
                x = x_try.copy()
                sk = lam * d
                x_tf = x_tf_try
                g_new = gradient(den, betaParam)
                yk = g_new - g
                g = g_new.copy()
                break
            lam = lam * beta
            if is_beta and adapt_beta:                                    # :798-800
                betaParam = betaParam - lr * beta_divergence_dbeta(den, gn, betaParam).mean()

        X = bounded(x)                                                    # :355-386
        D = np.divide(1, X)
        sk_s = np.multiply(sk, D)
        yk_s = np.multiply(yk, X)
        bk = np.dot(sk_s, yk)
        ck = np.dot(yk_s, sk)
        if bk <= 0:
            a1 = min(10 * alpha, alpha_max)
        else:
            a1 = min(alpha_max, max(alpha_min, np.sum(np.dot(sk_s, sk_s)) / bk))
        if ck <= 0:
            a2 = min(10 * alpha, alpha_max)
        else:
            a2 = min(alpha_max, max(alpha_min, ck / np.sum(np.dot(yk_s, yk_s))))
        alpha_hist[M_alpha - 1] = a2
        if iter_ <= 20:
            alpha = min(alpha_hist)
        elif a2 / a1 < tau:
            alpha = min(alpha_hist)
            tau = tau * 0.9
        else:
            alpha = a1
            tau = tau * 1.1

        if is_beta and schedule_lr:                                       # :842-844
            lr = learning_rate(lr0, lr_exp_param, epoch)

        iter_ += 1                                                        # :390-392
        times[iter_ - 1] = default_timer() - t0
        discr[iter_ - 1] = discr_coeff * fv
        if errflag:                                                       # :394-396 (index quirk kept)
            e = x - truth
            err[iter_] = np.sqrt(np.sum(e * e) / truth_sq)

        stop_value = float("nan")
        if stop_criterion == 2:                                           # :399-411
            stop_value = np.dot(sk, sk) / np.dot(x, x)
            keep_going = stop_value > tol
        elif stop_criterion == 3:
            stop_value = (f_hist[M - 1] - fv) / fv
            keep_going = stop_value > tol and stop_value >= 0
        elif stop_criterion == 4:
            stop_value = discr[iter_ - 1]
            keep_going = stop_value > tol
        if iter_ > MAXIT:
            keep_going = False

        tr.alpha.append(float(alpha)); tr.lam.append(float(lam)); tr.fv.append(float(fv))
        tr.beta_param.append(float(betaParam)); tr.trials.append(n_trials)
        tr.proj_evals.append(n_eval); tr.stop_value.append(float(stop_value))

        if not keep_going:                                                # :424-425
            x = x_before
        if is_beta and epoch == MAXIT:                                    # :881-882
            break

    x = x.reshape(shape) * scaling                                        # :428-438
    return OracleResult(x=x, iters=iter_ - 1, discr=discr[0:iter_], times=times[0:iter_],
                        err=(err[0:iter_] if errflag else None), beta_param=float(betaParam), trace=tr)


def sgp(gn, psf, bkg, **kw):
    """Reference-shaped wrapper: returns (x, iters, discr, times, err); sgp.py:41-438."""
    r = solve(gn, psf, bkg, divergence="kl", **kw)
    return r.x, r.iters, r.discr, r.times, r.err


def sgp_betaDiv(gn, psf, bkg, **kw):
    """Reference-shaped wrapper: returns (x, iters, discr, times, None); sgp.py:506-895."""
    r = solve(gn, psf, bkg, divergence="beta", **kw)
    return r.x, r.iters, r.discr, r.times, None
