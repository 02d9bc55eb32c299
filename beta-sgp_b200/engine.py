"""Batched front end of libbsgp: plans, buffers and the `solve_batch` call.

Host arrays (numpy) go through the library's ``*_host`` entry points; CUDA tensors (torch, used
only as device-memory handles) go through the asynchronous device entry points on torch's current
stream.  Nothing here computes: all arithmetic of the restoration loop runs in the CUDA kernels
of csrc/bsgp_kernels.cu, and there is no CPU fallback.

Mirrors the call shape of the reference's batched callers, which loop over images one at a time:
application_sgp_star_stamps.py:56-105 and application_sgp_subdivisions.py:83-107.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import BsgpError, DIV_BETA, DIV_KL, check, lib

_NP = {"float64": np.float64, "float32": np.float32}
_DT = {"float64": _capi.BSGP_F64, "float32": _capi.BSGP_F32}

_SOLVER_KEYS = ("region", "div_a", "div_at", "adjoint_second_psf", "init_recon", "proj_type", "stop_criterion", "MAXIT", "gamma", "beta", "alpha", "alpha_min", "alpha_max",
                "M_alpha", "tau", "M", "max_projs", "verbose", "ccd_sat_level", "scale_data", "errflag",
                "tol_convergence", "adapt_beta", "lr", "lr_exp_param", "schedule_lr")


class Plan:
    """One bsgp_plan: (ny, nx, dtype, device) -> twiddles, per-cluster scratch, PSF spectra."""

    def __init__(self, ny, nx, dtype="float64", device=0, cluster_size=0, threads=0):
        self.ny, self.nx, self.dtype, self.device = int(ny), int(nx), str(dtype), int(device)
        if self.dtype not in _DT:
            raise ValueError("dtype must be 'float64' or 'float32'")
        self._h = C.c_void_p()
        check(lib().bsgp_plan_create(self.ny, self.nx, _DT[self.dtype], self.device, C.byref(self._h)))
        if cluster_size or threads:
            check(lib().bsgp_plan_configure(self._h, int(cluster_size), int(threads)))
        self._psf_key = None

    @property
    def handle(self):
        return self._h

    def info(self):
        i = _capi.PlanInfo()
        check(lib().bsgp_plan_get_info(self._h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in i._fields_}

    def _image_tensor(self, t, name, lead=None):
        """Validate a CUDA tensor handed to the library as raw memory: [ny,nx] or [n,ny,nx] with the plan's image shape,
        on the plan's device; converted to the plan's dtype and made contiguous.  `lead`: allowed leading sizes."""
        import torch
        if not t.is_cuda:
            raise ValueError(f"{name} must be a CUDA tensor (or a numpy array)")
        if (t.device.index or 0) != self.device:
            raise ValueError(f"{name} lives on {t.device}, the plan on cuda:{self.device}")
        if t.dim() not in (2, 3) or tuple(t.shape[-2:]) != (self.ny, self.nx):
            raise ValueError(f"{name} has shape {tuple(t.shape)}; expected [{self.ny},{self.nx}] or [n,{self.ny},{self.nx}]")
        n = 1 if t.dim() == 2 else int(t.shape[0])
        if lead is not None and n not in lead:
            raise ValueError(f"{name} has {n} images; expected one of {sorted(set(lead))}")
        want = torch.float64 if self.dtype == "float64" else torch.float32
        return t.to(want).contiguous(), n

    def set_psf(self, psf):
        """psf: numpy [ny,nx] / [n,ny,nx] or a CUDA tensor of the same shapes."""
        if _is_tensor(psf):
            t, n = self._image_tensor(psf, "psf")
            check(lib().bsgp_set_psf(self._h, t.data_ptr(), n, _stream_ptr()))
            self._psf_keepalive = t
            return n
        a = np.ascontiguousarray(psf, dtype=_NP[self.dtype])
        n = 1 if a.ndim == 2 else a.shape[0]
        if a.shape[-2:] != (self.ny, self.nx):
            raise ValueError(f"PSF shape {a.shape[-2:]} must equal the image shape {(self.ny, self.nx)} "
                             "(the reference's numpy A/AT closure has the same requirement, sgp.py:108-120)")
        check(lib().bsgp_set_psf_host(self._h, a.ctypes.data, n))
        return n

    def set_psf_adjoint(self, psf):
        """Second kernel, used for A^T by the zero-padded operator (sgp.py:157); numpy or CUDA tensor, [ny,nx] / [n,ny,nx]."""
        if _is_tensor(psf):
            t, n = self._image_tensor(psf, "adjoint psf")
            check(lib().bsgp_set_psf_adjoint(self._h, t.data_ptr(), n, _stream_ptr()))
            self._psf_adj_keepalive = t
            return n
        a = np.ascontiguousarray(psf, dtype=_NP[self.dtype])
        n = 1 if a.ndim == 2 else a.shape[0]
        if a.shape[-2:] != (self.ny, self.nx):
            raise ValueError("adjoint kernel must be embedded in the plan's grid")
        check(lib().bsgp_set_psf_adjoint_host(self._h, a.ctypes.data, n))
        return n

    def apply_psf(self, x, adjoint=False):
        """A(x) / AT(x) for numpy [ny,nx] or [n,ny,nx] (sgp.py:111-120)."""
        a = np.ascontiguousarray(x, dtype=_NP[self.dtype])
        if a.ndim not in (2, 3) or a.shape[-2:] != (self.ny, self.nx):
            raise ValueError(f"x has shape {a.shape}; expected [{self.ny},{self.nx}] or [n,{self.ny},{self.nx}]")
        n = 1 if a.ndim == 2 else a.shape[0]
        y = np.empty_like(a)
        check(lib().bsgp_apply_psf_host(self._h, a.ctypes.data, y.ctypes.data, n, int(bool(adjoint))))
        return y

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().bsgp_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plans = {}


def get_plan(ny, nx, dtype="float64", device=0, cluster_size=0, threads=0):
    """Process-wide plan cache, one plan per (shape, dtype, device, tuning).  A plan holds mutable device state (PSF
    spectra, scratch, work queue); the library serialises the launches of one plan on the device (an event chained
    through every entry point, include/bsgp.h "Streams"), so sharing a cached plan between CUDA streams is safe but
    not concurrent: callers that want overlap create their own `Plan` per stream."""
    key = (int(ny), int(nx), str(dtype), int(device), int(cluster_size), int(threads))
    p = _plans.get(key)
    if p is None:
        p = Plan(*key)
        _plans[key] = p
    return p


# Width of the CTA configuration as a function of the number of images a GPU holds (measured on B200, 256 x 256 beta-SGP,
# tools/latency_probe.py): the default (clusters of 8 CTAs x 128 threads, 4 CTAs of different images per SM, 71 images in
# flight) has the best throughput but 200 us per iteration of ONE image; 16 x 128 reaches 118 us with 31 images in flight,
# 8 x 256 112 us with 15, 16 x 256 74 us with 7.  A GPU that holds fewer images than slots is bounded by its longest solve,
# so it trades slots for latency.  Thresholds from tools/width_probe.py (one GPU plays every rank of a world of W on the
# 320-tile field; slowest rank in ms for default / 16 x 128 / 8 x 256 / 16 x 256): 160 images per rank 38.8 / 42.7 / 54.7 /
# 73.6; 80: 28.8 / 24.5 / 28.8 / 38.0; 40: 23.3 / 16.2 / 17.7 / 20.8; 20: 20.4 / 13.1 / 13.1 / 11.6; 10: 19.8 / 11.8 / 13.0 /
# 8.4.  8 x 256 never wins.  (threshold on images per default slot, (cluster_size, threads)); first match wins.
_WIDTH_RULES = ((1.7, (0, 0)), (0.42, (16, 128)), (0.0, (16, 256)))


def auto_config(ny, nx, batch, dtype="float64"):
    """(cluster_size, threads) for `batch` images of ny x nx on one GPU; (0, 0) = the library's default."""
    npix = int(ny) * int(nx)
    pow2 = all(v >= 16 and v & (v - 1) == 0 for v in (int(ny), int(nx)))
    if not pow2 or npix * (8 if dtype == "float64" else 4) <= 64 * 1024 or npix >= (1 << 20):
        return (0, 0)                                   # stamps (one CTA each), embedded (dense / wrapped) plans and frame mode: one configuration
    slots = 71.0                                        # images in flight of the default configuration on 148 SMs
    for thr, cfg in _WIDTH_RULES:
        if batch >= thr * slots:
            if cfg[0] == 16 and not (ny % 64 == 0 and (nx // 2) % 32 == 0):
                continue
            return cfg
    return (0, 0)


def clear_plans():
    for p in _plans.values():
        p.close()
    _plans.clear()


def _is_tensor(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class BatchResult:
    x: object                 # [B,ny,nx] restored images
    iters: object             # [B]
    status: object            # [B] 0 = ok
    discr: object             # [B,MAXIT+1]
    times: object             # [B,MAXIT+1]
    stop_value: object
    err: object
    beta_final: object
    proj_evals: object        # [B] total projection evaluations
    ls_trials: object         # [B] total line-search evaluations
    scalars: object           # [B,8]
    trace: dict | None = None


def _queue_order(divergence, beta0):
    """Order in which the persistent clusters pull images from the work queue.  Iteration counts of beta-SGP
    grow as beta approaches 1 (28..162 iterations over the reference's five beta initialisations,
    application_sgp_subdivisions.py:69-107), so the images with the smallest |beta - 1| go first: the long
    solves do not end up alone at the tail of the queue.  Results do not depend on the order."""
    if divergence != "beta" or beta0.size < 2 or np.all(beta0 == beta0[0]):
        return None
    return np.ascontiguousarray(np.argsort(np.abs(beta0 - 1.0), kind="stable").astype(np.int32))


class PaddedGeometry:
    """Placement of an ny x nx image and a ky x kx kernel on the square power-of-two grid that
    astropy.convolution.convolve_fft uses with boundary='fill' (fft_pad = psf_pad = True): side
    2^ceil(log2(max(image + kernel shape))), both centred (sgp.py:138,157 call it with the defaults)."""

    def __init__(self, ny, nx, ky, kx):
        self.ny, self.nx, self.ky, self.kx = ny, nx, ky, kx
        self.P = int(2 ** np.ceil(np.log2(max(ny + ky, nx + kx))))
        c = self.P - (self.P + 1) // 2
        self.rows = slice(c - ny // 2, c + (ny + 1) // 2)
        self.cols = slice(c - nx // 2, c + (nx + 1) // 2)
        self.c = c
        self.region = (self.rows.start, self.rows.stop, self.cols.start, self.cols.stop)

    def embed(self, a, dtype):
        """[B,ny,nx] -> [B,P,P], zeros in the padding."""
        out = np.zeros(a.shape[:-2] + (self.P, self.P), dtype=dtype)
        out[..., self.rows, self.cols] = a
        return out

    def kernel(self, k, dtype):
        """Normalised kernel(s) [.., kh, kw] centred on the grid; returns (grid array, weight constant)."""
        k = np.asarray(k, dtype=np.float64)
        kh, kw = k.shape[-2:]
        k = k / k.sum(axis=(-2, -1), keepdims=True)
        out = np.zeros(k.shape[:-2] + (self.P, self.P), dtype=dtype)
        out[..., self.c - kh // 2:self.c + (kh + 1) // 2, self.c - kw // 2:self.c + (kw + 1) // 2] = k
        return out, float(out.reshape(-1, self.P * self.P)[0].sum())


def _params(divergence, kw, has_flux):
    unknown = set(kw) - set(_SOLVER_KEYS)
    if unknown:
        raise TypeError(f"unexpected keyword argument(s): {sorted(unknown)}")
    return _capi.make_params(DIV_KL if divergence == "kl" else DIV_BETA, has_flux=has_flux, **kw)


def solve_batch(gn, psf, bkg, divergence="beta", flux=None, betaParam=1.005, x0=None, obj=None, dtype="float64",
                device=0, trace=False, plan=None, psf_is_set=False, padded=False, x_out=None, **kw):
    """Restore a batch of independent images in one persistent kernel launch.

    gn [B,ny,nx]; psf [ny,nx] (shared) or [B,ny,nx]; bkg scalar, [B] or [B,ny,nx]; flux None or [B];
    betaParam scalar or [B].  Keyword arguments are those of sgp()/sgp_betaDiv() (sgp.py:41-47,
    506-513).  numpy in -> numpy out; CUDA tensors in -> CUDA tensors out (asynchronous on the
    current stream); PINNED CPU tensors in -> pinned CPU tensors out through the pipelined host path
    (upload, solve and download overlap; see bsgp_solve_batch_pinned in include/bsgp.h; ``x_out`` may name a
    pinned tensor that receives the restored images, so that a caller in a loop reuses its buffer); pageable CPU tensors
    are treated like numpy arrays.  ``padded=True`` selects the zero-padded operator of
    use_original_SGP_Afunction=False (sgp.py:121-161): images of any size, kernel of any (smaller or equal) size."""
    if divergence not in ("kl", "beta"):
        raise ValueError("divergence must be 'kl' or 'beta'")
    if padded:
        if _is_tensor(gn) and not gn.is_cuda:             # CPU tensors (pinned or not) take the numpy route of the padded operator
            def host(a):
                return a.numpy() if _is_tensor(a) and not a.is_cuda else a
            gn, psf, bkg, flux, betaParam, x0, obj = (host(a) for a in (gn, psf, bkg, flux, betaParam, x0, obj))
            x_out = None
        return _solve_batch_padded(gn, psf, bkg, divergence, flux, betaParam, x0, obj, dtype, device, trace, kw)
    if x_out is not None and not (_is_tensor(gn) and not gn.is_cuda and gn.is_pinned()):
        raise ValueError("x_out is only supported with pinned CPU tensor inputs")
    if _is_tensor(gn) and not gn.is_cuda and gn.is_pinned():
        return _solve_batch_pinned(gn, psf, bkg, divergence, flux, betaParam, x0, obj, device, trace, plan, psf_is_set, x_out, kw)
    if _is_tensor(gn) and gn.is_cuda:
        return _solve_batch_device(gn, psf, bkg, divergence, flux, betaParam, x0, obj, trace, plan, psf_is_set, kw)
    if _is_tensor(gn):                                    # pageable CPU tensor: the plain host path
        gn = gn.numpy()
        psf, bkg, flux, betaParam, x0, obj = (a.numpy() if _is_tensor(a) and not a.is_cuda else a for a in (psf, bkg, flux, betaParam, x0, obj))
    npdt = _NP[dtype]
    gn = np.ascontiguousarray(gn, dtype=npdt)
    if gn.ndim != 3:
        raise ValueError("gn must be [batch, ny, nx]")
    B, ny, nx = gn.shape
    p = _params(divergence, kw, flux is not None)
    plan = plan or get_plan(ny, nx, dtype, device)
    if not psf_is_set:
        n_psf = plan.set_psf(psf)
        if n_psf not in (1, B):
            raise ValueError("psf must be one image or one per batch entry")
    bkg = np.asarray(bkg, dtype=npdt)
    if bkg.ndim == 3:
        bkg_img, bkg_a = 1, np.ascontiguousarray(bkg)
        if bkg_a.shape != gn.shape:
            raise ValueError("bkg image stack must have the shape of gn")
    else:
        bkg_img, bkg_a = 0, np.ascontiguousarray(np.broadcast_to(bkg.reshape(-1), (B,)).astype(npdt))
    fl = None if flux is None else np.ascontiguousarray(np.broadcast_to(np.asarray(flux, dtype=np.float64).reshape(-1), (B,)))
    b0 = np.ascontiguousarray(np.broadcast_to(np.asarray(betaParam, dtype=np.float64).reshape(-1), (B,)))
    x0a = None if x0 is None else np.ascontiguousarray(x0, dtype=npdt)
    obja = None if obj is None else np.ascontiguousarray(obj, dtype=npdt)
    for name, a in (("x0", x0a), ("obj", obja)):
        if a is not None and a.shape != gn.shape:
            raise ValueError(f"{name} has shape {a.shape}; expected the shape of gn {gn.shape}")
    if (plan.ny, plan.nx) != (ny, nx) or plan.dtype != dtype:
        raise ValueError(f"plan is for {plan.ny}x{plan.nx} {plan.dtype} images, gn is {ny}x{nx} {dtype}")
    if p.init_recon == 1 and x0a is None:
        raise ValueError("init_recon=1 needs x0 (the sgp()/sgp_betaDiv() wrappers draw it with seed 42)")
    T = p.maxit + 1
    out = dict(x=np.empty((B, ny, nx), npdt), iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
               discr=np.zeros((B, T)), times=np.zeros((B, T)), stop_value=np.zeros((B, T)),
               err=np.zeros((B, T + 1)) if p.errflag else None, beta_final=np.zeros(B), proj_evals=np.zeros(B, np.int32),
               ls_trials=np.zeros(B, np.int32), scalars=np.zeros((B, _capi.NSCALARS)))
    tr = None
    if trace:
        tr = dict(alpha=np.zeros((B, T)), lam=np.zeros((B, T)), beta=np.zeros((B, T)), trials=np.zeros((B, T), np.int32),
                  evals=np.zeros((B, T), np.int32))

    def ptr(a):
        return None if a is None else a.ctypes.data

    order = _queue_order(divergence, b0)
    ci = _capi.Inputs(ptr(gn), ptr(bkg_a), bkg_img, ptr(fl), ptr(b0), ptr(x0a), ptr(obja), ptr(order))
    co = _capi.Outputs(ptr(out["x"]), ptr(out["iters"]), ptr(out["status"]), ptr(out["discr"]), ptr(out["times"]),
                       ptr(out["stop_value"]), ptr(out["err"]), ptr(out["beta_final"]), ptr(out["proj_evals"]),
                       ptr(out["ls_trials"]), ptr(out["scalars"]),
                       ptr(tr["alpha"]) if tr else None, ptr(tr["lam"]) if tr else None, ptr(tr["beta"]) if tr else None,
                       ptr(tr["trials"]) if tr else None, ptr(tr["evals"]) if tr else None)
    check(lib().bsgp_solve_batch_host(plan.handle, C.byref(p), B, C.byref(ci), C.byref(co)))
    return BatchResult(trace=tr, **out)


def _solve_batch_padded(gn, psf, bkg, divergence, flux, betaParam, x0, obj, dtype, device, trace, kw):
    """Embed into the padded grid, solve there with the window mask, crop (see PaddedGeometry).  numpy inputs give
    numpy outputs; CUDA-tensor images stay on the device (the small kernels may be numpy or tensors)."""
    on_dev = _is_tensor(gn)
    if on_dev:
        import torch
        dtype = {torch.float64: "float64", torch.float32: "float32"}[gn.dtype]
        device = gn.device.index or 0
    npdt = _NP[dtype]
    if not on_dev:
        gn = np.asarray(gn, dtype=npdt)
    if gn.ndim != 3:
        raise ValueError("gn must be [batch, ny, nx]")
    B, ny, nx = gn.shape
    psf_h = psf.detach().cpu().numpy() if _is_tensor(psf) else np.asarray(psf)
    psf_h = np.asarray(psf_h, dtype=np.float64)
    geo = PaddedGeometry(ny, nx, psf_h.shape[-2], psf_h.shape[-1])
    if geo.P > 8192:
        raise ValueError(f"padded grid {geo.P} exceeds the largest supported side (8192)")
    plan = get_plan(geo.P, geo.P, dtype, device)
    big, div_a = geo.kernel(psf_h, npdt)
    bigt, div_at = geo.kernel(np.conj(np.swapaxes(psf_h, -1, -2)), npdt)          # psf.conj().T, sgp.py:157
    extra = dict(region=geo.region, div_a=div_a, div_at=div_at, adjoint_second_psf=True)
    if on_dev:
        import torch
        with torch.cuda.device(gn.device):
            if plan.set_psf(torch.as_tensor(big, device=gn.device)) not in (1, B) or \
                    plan.set_psf_adjoint(torch.as_tensor(bigt, device=gn.device)) not in (1, B):
                raise ValueError("psf must be one kernel or one per batch entry")

            def emb(t):
                out = torch.zeros(t.shape[:-2] + (geo.P, geo.P), dtype=gn.dtype, device=gn.device)
                out[..., geo.rows, geo.cols] = t
                return out

            bkg_p = emb(bkg) if (_is_tensor(bkg) and bkg.dim() == 3) else bkg
            res = _solve_batch_device(emb(gn), None, bkg_p, divergence, flux, betaParam, None if x0 is None else emb(x0),
                                      None if obj is None else emb(obj), trace, plan, True, dict(kw, **extra))
            res.x = res.x[:, geo.rows, geo.cols].contiguous()
        return res
    if plan.set_psf(big) not in (1, B) or plan.set_psf_adjoint(bigt) not in (1, B):
        raise ValueError("psf must be one kernel or one per batch entry")
    bkg = np.asarray(bkg, dtype=npdt)
    bkg_p = geo.embed(bkg, npdt) if bkg.ndim == 3 else bkg
    res = solve_batch(geo.embed(gn, npdt), None, bkg_p, divergence=divergence, flux=flux, betaParam=betaParam,
                      x0=None if x0 is None else geo.embed(np.asarray(x0, dtype=npdt), npdt),
                      obj=None if obj is None else geo.embed(np.asarray(obj, dtype=npdt), npdt), dtype=dtype, device=device,
                      trace=trace, plan=plan, psf_is_set=True, **extra, **kw)
    res.x = np.ascontiguousarray(res.x[:, geo.rows, geo.cols])
    return res


def _solve_batch_device(gn, psf, bkg, divergence, flux, betaParam, x0, obj, trace, plan, psf_is_set, kw):
    import torch
    if not gn.is_cuda:
        raise BsgpError("tensor inputs must live on a CUDA device (there is no CPU path)")
    dev = gn.device
    if gn.dtype not in (torch.float64, torch.float32):
        raise ValueError("gn must be float64 or float32")
    dtype = {torch.float64: "float64", torch.float32: "float32"}[gn.dtype]
    if gn.dim() != 3:
        raise ValueError("gn must be [batch, ny, nx]")
    gn = gn.contiguous()
    B, ny, nx = gn.shape
    p = _params(divergence, kw, flux is not None)
    plan = plan or get_plan(ny, nx, dtype, dev.index or 0)
    if (plan.ny, plan.nx) != (ny, nx) or plan.dtype != dtype or plan.device != (dev.index or 0):
        raise ValueError(f"plan is for {plan.ny}x{plan.nx} {plan.dtype} images on cuda:{plan.device}, gn is {ny}x{nx} {dtype} on {dev}")

    def image_stack(t, name):
        """[B,ny,nx] tensor of gn's dtype on gn's device (raw pointers go to the library: a short or foreign array would be read out of bounds)"""
        if t is None:
            return None
        if not _is_tensor(t):
            t = torch.as_tensor(np.ascontiguousarray(t), device=dev)
        if t.device != dev or tuple(t.shape) != (B, ny, nx):
            raise ValueError(f"{name} must have the shape {(B, ny, nx)} of gn and live on {dev}; got {tuple(t.shape)} on {t.device}")
        return t.to(gn.dtype).contiguous()

    with torch.cuda.device(dev):
        if not psf_is_set:
            n_psf = plan.set_psf(psf if _is_tensor(psf) else torch.as_tensor(np.ascontiguousarray(psf), dtype=gn.dtype, device=dev))
            if n_psf not in (1, B):
                raise ValueError("psf must be one image or one per batch entry")
        if _is_tensor(bkg) and bkg.dim() == 3:
            bkg_img, bkg_t = 1, image_stack(bkg, "bkg")
        else:
            bkg_img = 0
            bkg_t = (bkg if _is_tensor(bkg) else torch.as_tensor(np.asarray(bkg, dtype=np.float64), device=dev)).to(gn.dtype).reshape(-1).expand(B).contiguous()
        f64 = dict(dtype=torch.float64, device=dev)
        fl = None if flux is None else (flux if _is_tensor(flux) else torch.as_tensor(np.asarray(flux, dtype=np.float64), device=dev)).to(torch.float64).reshape(-1).expand(B).contiguous()
        b0 = (betaParam if _is_tensor(betaParam) else torch.as_tensor(np.asarray(betaParam, dtype=np.float64), device=dev)).to(torch.float64).reshape(-1).expand(B).contiguous()
        x0t, objt = image_stack(x0, "x0"), image_stack(obj, "obj")
        if p.init_recon == 1 and x0t is None:
            raise ValueError("init_recon=1 needs x0 (the sgp()/sgp_betaDiv() wrappers draw it with seed 42)")
        for name, t in (("bkg", bkg_t), ("flux", fl), ("betaParam", b0)):
            if t is not None and t.device != dev:
                raise ValueError(f"{name} lives on {t.device}, gn on {dev}")
        T = p.maxit + 1
        i32 = dict(dtype=torch.int32, device=dev)
        out = dict(x=torch.empty_like(gn), iters=torch.zeros(B, **i32), status=torch.zeros(B, **i32),
                   discr=torch.zeros(B, T, **f64), times=torch.zeros(B, T, **f64), stop_value=torch.zeros(B, T, **f64),
                   err=torch.zeros(B, T + 1, **f64) if p.errflag else None, beta_final=torch.zeros(B, **f64),
                   proj_evals=torch.zeros(B, **i32), ls_trials=torch.zeros(B, **i32),
                   scalars=torch.zeros(B, _capi.NSCALARS, **f64))
        tr = None
        if trace:
            tr = dict(alpha=torch.zeros(B, T, **f64), lam=torch.zeros(B, T, **f64), beta=torch.zeros(B, T, **f64),
                      trials=torch.zeros(B, T, **i32), evals=torch.zeros(B, T, **i32))

        def ptr(t):
            return None if t is None else t.data_ptr()

        order = None
        if divergence == "beta" and B > 1:
            # longest expected solve first (see _queue_order); computed on the device, no host sync
            order = torch.argsort((b0 - 1.0).abs(), stable=True).to(torch.int32)
        ci = _capi.Inputs(ptr(gn), ptr(bkg_t), bkg_img, ptr(fl), ptr(b0), ptr(x0t), ptr(objt), ptr(order))
        co = _capi.Outputs(ptr(out["x"]), ptr(out["iters"]), ptr(out["status"]), ptr(out["discr"]), ptr(out["times"]),
                           ptr(out["stop_value"]), ptr(out["err"]), ptr(out["beta_final"]), ptr(out["proj_evals"]),
                           ptr(out["ls_trials"]), ptr(out["scalars"]),
                           ptr(tr["alpha"]) if tr else None, ptr(tr["lam"]) if tr else None, ptr(tr["beta"]) if tr else None,
                           ptr(tr["trials"]) if tr else None, ptr(tr["evals"]) if tr else None)
        check(lib().bsgp_solve_batch(plan.handle, C.byref(p), B, C.byref(ci), C.byref(co), _stream_ptr()))
        res = BatchResult(trace=tr, **out)
        res._keepalive = (gn, bkg_t, fl, b0, x0t, objt, order)
    return res


def _solve_batch_pinned(gn, psf, bkg, divergence, flux, betaParam, x0, obj, device, trace, plan, psf_is_set, x_out, kw):
    """Page-locked CPU tensors -> page-locked CPU tensors; one call of bsgp_solve_batch_pinned (returns when the results
    are in host memory).  `device` selects the GPU.  Image-shaped arguments (bkg stack, x0, obj) must be pinned too."""
    import torch
    dtype = {torch.float64: "float64", torch.float32: "float32"}[gn.dtype]
    npdt = _NP[dtype]
    gn = gn.contiguous()
    if gn.dim() != 3:
        raise ValueError("gn must be [batch, ny, nx]")
    B, ny, nx = gn.shape
    p = _params(divergence, kw, flux is not None)
    plan = plan or get_plan(ny, nx, dtype, device)
    if (plan.ny, plan.nx) != (ny, nx) or plan.dtype != dtype:
        raise ValueError(f"plan is for {plan.ny}x{plan.nx} {plan.dtype} images, gn is {ny}x{nx} {dtype}")
    with torch.cuda.device(plan.device):
        if not psf_is_set:
            psf_t = psf if _is_tensor(psf) else torch.as_tensor(np.ascontiguousarray(psf))
            n_psf = plan.set_psf(psf_t.to(device=f"cuda:{plan.device}", dtype=gn.dtype, non_blocking=True))
            if n_psf not in (1, B):
                raise ValueError("psf must be one image or one per batch entry")

        def pinned_image(t, name):
            if t is None:
                return None
            if not (_is_tensor(t) and not t.is_cuda and t.is_pinned() and t.dtype == gn.dtype and tuple(t.shape) == tuple(gn.shape)):
                raise ValueError(f"{name} must be a pinned CPU tensor of the shape and dtype of gn")
            return t.contiguous()

        if _is_tensor(bkg) and bkg.dim() == 3:
            bkg_img, bkg_t = 1, pinned_image(bkg, "bkg")
            bkg_ptr = bkg_t.data_ptr()
        else:
            bkg_img, bkg_t = 0, np.ascontiguousarray(np.broadcast_to(np.asarray(bkg, dtype=npdt).reshape(-1), (B,)).astype(npdt))
            bkg_ptr = bkg_t.ctypes.data

        def small(a):
            return None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1), (B,)))

        fl, b0 = small(flux), small(betaParam)
        x0t, objt = pinned_image(x0, "x0"), pinned_image(obj, "obj")
        if p.init_recon == 1 and x0t is None:
            raise ValueError("init_recon=1 needs x0 (the sgp()/sgp_betaDiv() wrappers draw it with seed 42)")
        T = p.maxit + 1
        if x_out is None:
            x = torch.empty((B, ny, nx), dtype=gn.dtype, pin_memory=True)
        else:
            x = x_out
            if not (_is_tensor(x) and not x.is_cuda and x.is_pinned() and x.dtype == gn.dtype and tuple(x.shape) == (B, ny, nx) and x.is_contiguous()):
                raise ValueError("x_out must be a contiguous pinned CPU tensor of the shape and dtype of gn")

        # the small outputs are page-locked as well (torch's caching host allocator makes repeated calls cheap), handed
        # back as numpy views: the copies behind the kernel run at full PCIe speed and never block on a staging buffer
        def hbuf(shape, dt=np.float64):
            return torch.empty(shape, dtype=torch.float64 if dt is np.float64 else torch.int32, pin_memory=True).numpy()

        out = dict(iters=hbuf(B, np.int32), status=hbuf(B, np.int32), discr=hbuf((B, T)), times=hbuf((B, T)),
                   stop_value=hbuf((B, T)), err=hbuf((B, T + 1)) if p.errflag else None, beta_final=hbuf(B),
                   proj_evals=hbuf(B, np.int32), ls_trials=hbuf(B, np.int32), scalars=hbuf((B, _capi.NSCALARS)))
        tr = None
        if trace:
            tr = dict(alpha=hbuf((B, T)), lam=hbuf((B, T)), beta=hbuf((B, T)), trials=hbuf((B, T), np.int32), evals=hbuf((B, T), np.int32))

        def ptr(a):
            return None if a is None else (a.data_ptr() if _is_tensor(a) else a.ctypes.data)

        order = _queue_order(divergence, b0)
        ci = _capi.Inputs(gn.data_ptr(), bkg_ptr, bkg_img, ptr(fl), ptr(b0), ptr(x0t), ptr(objt), ptr(order))
        co = _capi.Outputs(x.data_ptr(), ptr(out["iters"]), ptr(out["status"]), ptr(out["discr"]), ptr(out["times"]),
                           ptr(out["stop_value"]), ptr(out["err"]), ptr(out["beta_final"]), ptr(out["proj_evals"]),
                           ptr(out["ls_trials"]), ptr(out["scalars"]),
                           ptr(tr["alpha"]) if tr else None, ptr(tr["lam"]) if tr else None, ptr(tr["beta"]) if tr else None,
                           ptr(tr["trials"]) if tr else None, ptr(tr["evals"]) if tr else None)
        check(lib().bsgp_solve_batch_pinned(plan.handle, C.byref(p), B, C.byref(ci), C.byref(co), _stream_ptr()))
    for i in np.nonzero(out["status"] == _capi.ST_INPUT_TIMEOUT)[0]:      # skipped images: defined (zero) pixels, status says why
        x[int(i)].zero_()
    return BatchResult(x=x, trace=tr, **out)


def project_batch(b, c, dia, sat_cap=None, lambda_=0.0, dlambda_=1.0, tol_lam=1e-11, max_projs=1000, biter=0, siter=0, device=0):
    """projectDF for a batch of problems: c, dia [B,n]; b [B] (flux_conserve_proj.py:7-144)."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    dia = np.ascontiguousarray(dia, dtype=np.float64)
    if c.ndim == 1:
        c, dia = c[None], dia[None]
    if c.ndim != 2 or dia.shape != c.shape:
        raise ValueError(f"c and dia must both be [B, n]; got {c.shape} and {dia.shape}")
    B, n = c.shape
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64).reshape(-1), (B,)))
    x = np.empty_like(c)
    ev = np.zeros(B, np.int32)
    st = np.zeros(B, np.int32)
    cap = float("nan") if sat_cap is None else float(sat_cap)
    check(lib().bsgp_project_df_host(b.ctypes.data, c.ctypes.data, dia.ctypes.data, n, B, cap, float(lambda_), float(dlambda_),
                                     float(tol_lam), int(max_projs), int(biter), int(siter), x.ctypes.data, ev.ctypes.data, st.ctypes.data,
                                     int(device)))
    return x, ev, st
