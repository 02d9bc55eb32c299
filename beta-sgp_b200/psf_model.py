"""DIAPL PSF model on the device: the upstream side of the per-stamp PSFs of application_sgp_star_stamps.py.

Mirrors psf/psf_calculate.py of the reference: ``PSF(txt_file)`` reads a DIAPL ``getpsf`` coefficient file
(psf_calculate.py:9-46), ``get_psf_mat()`` evaluates the spatially constant part of the model on a 31 x 31 grid
(:52-111: two elliptical Gaussians x a degree-2 local polynomial) and ``normalize_psf_mat()`` divides by the sum
(:130-139).  The arithmetic runs in ``bsgp_psf_model_eval`` (CUDA); there is no CPU path.  ``evaluate_batch`` writes
many PSFs straight into image-shaped CUDA tensors, centred where ``Plan.set_psf`` expects them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib

LDEG = 2            # psf_calculate.py:23 (the local polynomial degree the reference evaluates)


def evaluate_batch(params, ngauss, hw, shape, normalize=True, dtype="float64", device=0):
    """params [n, 5 + 6*ngauss] = (cos, sin, ax, ay, sigma_inc, coefficients) -> CUDA tensor [n, ny, nx]."""
    import torch
    p = np.ascontiguousarray(params, dtype=np.float64)
    if p.ndim != 2 or p.shape[1] != 5 + 6 * int(ngauss):
        raise ValueError("params must be [n, 5 + 6*ngauss]")
    ny, nx = int(shape[0]), int(shape[1])
    with torch.cuda.device(device):
        pd = torch.as_tensor(p, device=f"cuda:{device}")
        out = torch.empty((p.shape[0], ny, nx), dtype=torch.float64 if dtype == "float64" else torch.float32, device=f"cuda:{device}")
        check(lib().bsgp_psf_model_eval(pd.data_ptr(), p.shape[0], int(ngauss), int(hw), ny, nx, int(bool(normalize)),
                                        _capi.BSGP_F64 if dtype == "float64" else _capi.BSGP_F32, out.data_ptr(), int(device),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out._keepalive = pd
    return out


class PSF:
    """psf_calculate.py:8-46: the attributes of a DIAPL PSF file, in the file's order."""

    def __init__(self, txt_file):
        self.ldeg = LDEG
        self.sdeg = 1
        with open(txt_file) as f:
            data = [float(line.rstrip("\n")) for line in f if line.strip()]
        self.hw, self.ndeg_spat, self.ndeg_local, self.ngauss = (int(v) for v in data[:4])
        (self.recenter, self.cos, self.sin, self.ax, self.ay, self.sigma_inc, self.sigma_mscale, self.fitrad, self.x_orig,
         self.y_orig) = data[4:14]
        self.vec_coeffs = data[14:]
        self.ntot = self.ngauss * (self.ndeg_local + 1) * (self.ndeg_local + 2) / 2
        self.ntot *= (self.ndeg_spat + 1) * (self.ndeg_spat + 2) / 2

    @property
    def coeffs(self):
        return self.vec_coeffs

    def params(self):
        """The row bsgp_psf_model_eval consumes: the first ngauss * 6 coefficients are the ones calc_psf_pix reads."""
        ncomp = self.ngauss * (self.ldeg + 1) * (self.ldeg + 2) // 2
        return np.array([self.cos, self.sin, self.ax, self.ay, self.sigma_inc] + list(self.vec_coeffs[:ncomp]), dtype=np.float64)

    def _mat(self, normalize, device=0):
        side = 2 * self.hw + 1
        out = evaluate_batch(self.params()[None], self.ngauss, self.hw, (side, side), normalize=normalize, device=device)
        return out[0].cpu().numpy()

    def get_psf_mat(self, device=0):
        """31 x 31 (2 hw + 1 squared) model values, psf_calculate.py:92-111."""
        self.psf_mat = self._mat(False, device)
        return self.psf_mat

    def normalize_psf_mat(self, device=0):
        """get_psf_mat() / sum, psf_calculate.py:130-139."""
        return self._mat(True, device)

    def embedded(self, shape, dtype="float64", device=0):
        """Normalised PSF centred at (ny/2, nx/2) of an image-shaped CUDA tensor (ready for Plan.set_psf)."""
        return evaluate_batch(self.params()[None], self.ngauss, self.hw, shape, normalize=True, dtype=dtype, device=device)[0]
