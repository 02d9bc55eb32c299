"""Synthetic star-field workloads of the shapes named in BASELINE.json.

The M13 frames used by the reference's application scripts are not available offline, so the
benchmarks and the large parity tests use seeded Moffat-PSF star fields with Poisson noise and a
sky background, shaped like the inputs of
  * application_sgp_star_stamps.py:56-105   (config 3: 32x32 stamps, one PSF per stamp)
  * application_sgp_subdivisions.py:83-107  (config 4: 256x256 tiles of a crowded field, 2-D sky)
  * a single large frame                      (config 5)
numpy only; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

# application_sgp_star_stamps.py:69-75 draws its five beta initialisations with the legacy
# numpy generator: seed s -> normal(1, 0.05)
BETA_INIT_SEEDS = (0, 42, 951, 93, 810)


def beta_inits():
    out = []
    for s in BETA_INIT_SEEDS:
        out.append(float(np.random.RandomState(s).normal(loc=1, scale=0.05)))
    return out


# keyword set used by application_sgp_star_stamps.py:82-89 (DEFAULT_PARAMS at sgp.py:34)
STAMP_KWARGS = dict(gamma=1e-4, beta=0.4, alpha_min=1e-5, alpha_max=1e5, alpha=1e1, M_alpha=3, tau=0.5, M=1,
                    proj_type=1, max_projs=1000, init_recon=2, stop_criterion=3, verbose=True,
                    ccd_sat_level=65000, scale_data=True, lr=1e-3, lr_exp_param=0.1, schedule_lr=True,
                    adapt_beta=True, MAXIT=500, tol_convergence=1e-4)

# keyword set used by application_sgp_subdivisions.py:84-91 (numpy A/A^T closure)
TILE_KWARGS = dict(gamma=1e-4, beta=0.4, alpha_min=1e-5, alpha_max=1e5, alpha=1e1, M_alpha=3, tau=0.5, M=1,
                   proj_type=1, max_projs=1000, init_recon=2, stop_criterion=3, verbose=True,
                   ccd_sat_level=65000, scale_data=True, lr=1e-3, lr_exp_param=0.1, schedule_lr=True,
                   adapt_beta=False, MAXIT=500, tol_convergence=1e-5)


def moffat_psf(ny, nx, fwhm, index=2.5, axis_ratio=1.0, theta=0.0, centre=None):
    """Elliptical Moffat profile sampled on an ny x nx grid, peak at ``centre`` (default (ny//2, nx//2),
    which np.fft.fftshift moves to the origin for even sizes), normalised to unit sum."""
    cy, cx = (ny // 2, nx // 2) if centre is None else centre
    yy, xx = np.mgrid[0:ny, 0:nx].astype(np.float64)
    dy, dx = yy - cy, xx - cx
    ct, st = np.cos(theta), np.sin(theta)
    u = ct * dx + st * dy
    v = (-st * dx + ct * dy) * axis_ratio
    a = fwhm / (2.0 * np.sqrt(2.0 ** (1.0 / index) - 1.0))
    p = (1.0 + (u * u + v * v) / (a * a)) ** (-index)
    p /= p.sum()
    p /= p.sum()
    return p


def _circ_conv(img, psf):
    tf = np.fft.rfft2(np.fft.fftshift(psf, axes=(-2, -1)))
    return np.fft.irfft2(tf * np.fft.rfft2(img), s=img.shape[-2:])


def star_stamps(count, size=32, seed=12345):
    """Config 3.  Returns dict(gn[B,s,s], psf[B,s,s], bkg[B], flux[B], beta0[B], obj[B,s,s])."""
    rng = np.random.default_rng(seed)
    fwhm = rng.uniform(2.5, 4.5, count)
    ratio = rng.uniform(1.0, 1.3, count)
    theta = rng.uniform(0.0, np.pi, count)
    off = rng.uniform(-2.0, 2.0, (count, 2))
    amp = 10.0 ** rng.uniform(4.0, 5.5, count)
    sky = rng.uniform(50.0, 800.0, count)
    psf = np.empty((count, size, size))
    obj = np.zeros((count, size, size))
    c = size // 2
    for i in range(count):
        psf[i] = moffat_psf(size, size, fwhm[i], 2.5, ratio[i], theta[i])
        # bilinear splat of one star at a sub-pixel position near the centre
        y, x = c + off[i, 0], c + off[i, 1]
        y0, x0 = int(np.floor(y)), int(np.floor(x))
        fy, fx = y - y0, x - x0
        obj[i, y0, x0] += amp[i] * (1 - fy) * (1 - fx)
        obj[i, y0, x0 + 1] += amp[i] * (1 - fy) * fx
        obj[i, y0 + 1, x0] += amp[i] * fy * (1 - fx)
        obj[i, y0 + 1, x0 + 1] += amp[i] * fy * fx
    mean = np.maximum(_circ_conv(obj, psf), 0.0) + sky[:, None, None]
    gn = rng.poisson(mean).astype(np.float64)
    flux = (gn - sky[:, None, None]).sum(axis=(1, 2))
    assert np.all(flux > 0), "generator produced a stamp with non-positive flux"
    b5 = beta_inits()
    beta0 = np.array([b5[i % 5] for i in range(count)])
    return dict(gn=gn, psf=psf, bkg=sky.copy(), flux=flux, beta0=beta0, obj=obj)


def star_cutouts(count, psf, seed=31):
    """Star cut-outs of the PSF's own shape with ONE PSF shared by all of them: the literal call shape of
    application_sgp_star_stamps.py:24,58,82-89 (31 x 31 cut-outs restored with the subdivision's 31 x 31 PSF image,
    e.g. the one the reference ships as psf/psfccfbrd210048_1_1_img.fits).  One star within +-2 px of the centre,
    flux 10^U[4,5.5], sky U[50,800], Poisson noise; the image is formed with the reference's own operator
    (circular, fftshift placement: sgp.py:109-117), whatever the parity of the side.
    Returns dict(gn[B,n,n], psf[n,n], bkg[B], flux[B], beta0[B], obj[B,n,n])."""
    psf = np.asarray(psf, dtype=np.float64)
    ny, nx = psf.shape
    rng = np.random.default_rng(seed)
    off = rng.uniform(-2.0, 2.0, (count, 2))
    amp = 10.0 ** rng.uniform(4.0, 5.5, count)
    sky = rng.uniform(50.0, 800.0, count)
    obj = np.zeros((count, ny, nx))
    for i in range(count):
        y, x = ny // 2 + off[i, 0], nx // 2 + off[i, 1]
        y0, x0 = int(np.floor(y)), int(np.floor(x))
        fy, fx = y - y0, x - x0
        obj[i, y0, x0] += amp[i] * (1 - fy) * (1 - fx)
        obj[i, y0, x0 + 1] += amp[i] * (1 - fy) * fx
        obj[i, y0 + 1, x0] += amp[i] * fy * (1 - fx)
        obj[i, y0 + 1, x0 + 1] += amp[i] * fy * fx
    tf = np.fft.fftn(np.fft.fftshift(psf))
    mean = np.maximum(np.real(np.fft.ifftn(tf * np.fft.fftn(obj, axes=(-2, -1)), axes=(-2, -1))), 0.0) + sky[:, None, None]
    gn = rng.poisson(mean).astype(np.float64)
    flux = (gn - sky[:, None, None]).sum(axis=(1, 2))
    assert np.all(flux > 0), "generator produced a cut-out with non-positive flux"
    b5 = beta_inits()
    beta0 = np.array([b5[i % 5] for i in range(count)])
    return dict(gn=gn, psf=psf, bkg=sky.copy(), flux=flux, beta0=beta0, obj=obj)


def crowded_field(size=2048, seed=2024, density=1.0 / 400.0, fwhm=3.5, tile=256):
    """A crowded frame: stars of flux 10^U[3,5], sky 300 + smooth gradient, Moffat PSF.
    Returns dict(frame, sky, truth, psf_tile) where psf_tile is the PSF embedded at
    (tile//2, tile//2) of a tile x tile array."""
    rng = np.random.default_rng(seed)
    nstar = int(size * size * density)
    ys = rng.uniform(0, size - 1, nstar)
    xs = rng.uniform(0, size - 1, nstar)
    amp = 10.0 ** rng.uniform(3.0, 5.0, nstar)
    truth = np.zeros((size, size))
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    fy = ys - y0; fx = xs - x0
    y1 = np.minimum(y0 + 1, size - 1); x1 = np.minimum(x0 + 1, size - 1)
    np.add.at(truth, (y0, x0), amp * (1 - fy) * (1 - fx))
    np.add.at(truth, (y0, x1), amp * (1 - fy) * fx)
    np.add.at(truth, (y1, x0), amp * fy * (1 - fx))
    np.add.at(truth, (y1, x1), amp * fy * fx)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64) / size
    sky = 300.0 + 40.0 * xx + 25.0 * yy + 15.0 * np.sin(2 * np.pi * xx) * np.cos(2 * np.pi * yy)
    psf_full = moffat_psf(size, size, fwhm)
    mean = np.maximum(_circ_conv(truth, psf_full), 0.0) + sky
    frame = rng.poisson(mean).astype(np.float64)
    return dict(frame=frame, sky=sky, truth=truth, psf_tile=moffat_psf(tile, tile, fwhm))


def tile_boxes(height, width, tile, overlap=0):
    """Tile origins; same enumeration as utils.py:332-375 (`calculate_slice_bboxes`) with an absolute
    overlap in pixels: rows outer, columns inner, last tile of a row/column clamped to the border."""
    boxes = []
    y_max = y_min = 0
    while y_max < height:
        x_min = x_max = 0
        y_max = y_min + tile
        while x_max < width:
            x_max = x_min + tile
            xe, ye = min(width, x_max), min(height, y_max)
            boxes.append((max(0, ye - tile), max(0, xe - tile)))
            x_min = x_max - overlap
        y_min = y_max - overlap
    return boxes


def field_tiles(size=2048, tile=256, seed=2024, n_beta=5, max_tiles=None):
    """Config 4.  tiles x n_beta independent solves: dict(gn[B,t,t], bkg[B,t,t], psf[t,t] (shared),
    flux[B], beta0[B], origin[B,2])."""
    f = crowded_field(size, seed, tile=tile)
    boxes = tile_boxes(size, size, tile, 0)
    if max_tiles is not None:
        boxes = boxes[:max_tiles]
    b5 = beta_inits()[:n_beta]
    gn, bkg, flux, beta0, origin = [], [], [], [], []
    for (y, x) in boxes:
        g = f["frame"][y:y + tile, x:x + tile]
        s = f["sky"][y:y + tile, x:x + tile]
        fl = float((g - s).sum())
        assert fl > 0
        for b in b5:
            gn.append(g); bkg.append(s); flux.append(fl); beta0.append(b); origin.append((y, x))
    return dict(gn=np.ascontiguousarray(np.stack(gn)), bkg=np.ascontiguousarray(np.stack(bkg)),
                psf=f["psf_tile"], flux=np.array(flux), beta0=np.array(beta0), origin=np.array(origin))


def single_frame(size=8192, seed=77, fwhm=3.5):
    """Config 5: one large frame with the PSF embedded in a frame-sized array."""
    f = crowded_field(size, seed, fwhm=fwhm, tile=size)
    flux = float((f["frame"] - f["sky"]).sum())
    return dict(gn=f["frame"], bkg=f["sky"], psf=f["psf_tile"], flux=flux)
