"""Subdivision of a frame into overlapping tiles and re-assembly: the callers either side of the batched solve.

Mirrors restoration/utils.py of the reference: ``calculate_slice_bboxes`` (utils.py:332-375, same signature and return
value), ``create_subdivisions`` (utils.py:378-389; returns arrays instead of astropy Cutout2D objects) and
``reconstruct_full_image_from_patches`` (utils.py:392-397; the reference re-projects FITS files with
``reproject_and_coadd``, which needs WCS headers — here the tiles are cross-faded over their overlap, so the result is
NOT comparable bit for bit).  ``restore_frame`` chains extract -> one batched beta-SGP launch -> assemble, i.e. the
shape of BASELINE config 4 with a frame going in and a frame coming out.  The index arithmetic runs in libbsgp (host),
the data movement in two CUDA kernels; there is no CPU path for the latter.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib

_DT = {"float64": _capi.BSGP_F64, "float32": _capi.BSGP_F32}


def calculate_slice_bboxes(image_height, image_width, slice_height=512, slice_width=512, overlap_height_ratio=0.2,
                           overlap_width_ratio=0.2):
    """Bounding boxes [xmin, ymin, xmax, ymax] of the overlapping slices, in the reference's order (utils.py:332-375)."""
    n = C.c_int(0)
    args = (int(image_height), int(image_width), int(slice_height), int(slice_width), float(overlap_height_ratio), float(overlap_width_ratio))
    check(lib().bsgp_tile_boxes(*args, None, 0, C.byref(n)))
    boxes = np.zeros((n.value, 4), np.int32)
    check(lib().bsgp_tile_boxes(*args, boxes.ctypes.data, n.value, C.byref(n)))
    return boxes.tolist()


def tile_origins(shape, subdiv_shape=(100, 100), overlap=10):
    """(y0, x0) of every tile for create_subdivisions' arguments (utils.py:381-384: ratios overlap / subdiv_shape)."""
    boxes = calculate_slice_bboxes(shape[0], shape[1], subdiv_shape[0], subdiv_shape[1], overlap / subdiv_shape[0], overlap / subdiv_shape[1])
    return np.ascontiguousarray([[b[1], b[0]] for b in boxes], dtype=np.int32)


def _torch():
    import torch
    return torch


def _as_device(a, device, dtype=None):
    torch = _torch()
    if type(a).__module__.startswith("torch"):
        t = a
        if not t.is_cuda:
            t = t.to(f"cuda:{device}", non_blocking=True)
    else:
        t = torch.as_tensor(np.ascontiguousarray(a), device=f"cuda:{device}")
    if dtype is not None:
        t = t.to(dtype)
    return t.contiguous()


def create_subdivisions(image, subdiv_shape=(100, 100), overlap=10, device=0):
    """Tiles [n, h, w] (CUDA tensor) and their origins [n, 2] (numpy, (y0, x0)) of `image` (numpy or tensor)."""
    torch = _torch()
    img = _as_device(image, device)
    if img.dtype not in (torch.float64, torch.float32):
        img = img.to(torch.float64)
    dt = "float64" if img.dtype == torch.float64 else "float32"
    H, W = img.shape
    org = tile_origins((H, W), subdiv_shape, overlap)
    n, (th, tw) = len(org), subdiv_shape
    with torch.cuda.device(img.device):
        org_d = torch.as_tensor(org, device=img.device)
        tiles = torch.empty((n, th, tw), dtype=img.dtype, device=img.device)
        check(lib().bsgp_extract_tiles(img.data_ptr(), H, W, _DT[dt], org_d.data_ptr(), n, th, tw, tiles.data_ptr(), img.device.index or 0,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        tiles._keepalive = (img, org_d)
    return tiles, org


def reconstruct_full_image_from_patches(tiles, origins, shape, feather=None, device=0):
    """Frame [H, W] (CUDA tensor) from tiles [n, h, w] and origins [n, 2]: cross-fade of `feather` pixels at tile borders
    (default: half the tile's smaller side, i.e. a full linear blend across any overlap; 1 = plain average)."""
    torch = _torch()
    t = _as_device(tiles, device)
    dt = "float64" if t.dtype == torch.float64 else "float32"
    n, th, tw = t.shape
    H, W = int(shape[0]), int(shape[1])
    if feather is None:
        feather = max(1, min(th, tw) // 2)
    with torch.cuda.device(t.device):
        org_d = torch.as_tensor(np.ascontiguousarray(origins, dtype=np.int32), device=t.device)
        out = torch.empty((H, W), dtype=t.dtype, device=t.device)
        check(lib().bsgp_assemble_tiles(t.data_ptr(), org_d.data_ptr(), n, th, tw, _DT[dt], int(feather), out.data_ptr(), H, W, t.device.index or 0,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out._keepalive = (t, org_d)
    return out


def restore_frame(frame, psf, bkg, subdiv_shape=(256, 256), overlap=0, betaParam=1.005, divergence="beta", feather=None, device=0,
                  on_failure="raise", **kw):
    """Frame in, frame out: tiles of `subdiv_shape` (any shape; powers of two run on their own FFT grid) are cut on the device, restored
    in one persistent-kernel launch (each tile conserves its own flux sum(gn - bkg), the solver's default, sgp.py:661-666)
    and cross-faded back.  `bkg` is a scalar or a background map of the frame's shape; `psf` has the tile's shape.
    Returns (restored frame as a CUDA tensor, BatchResult of the tiles, origins)."""
    from .engine import solve_batch
    torch = _torch()
    gn, org = create_subdivisions(frame, subdiv_shape, overlap, device)
    if np.ndim(bkg) == 2 or (type(bkg).__module__.startswith("torch") and bkg.dim() == 2):
        bk, _ = create_subdivisions(bkg, subdiv_shape, overlap, device)
        bk = bk.to(gn.dtype)
    else:
        bk = torch.full((gn.shape[0],), float(bkg), dtype=gn.dtype, device=gn.device)
    kw.setdefault("proj_type", 1)
    # flux = None: every tile conserves its own sum(gn - bkg), the default of sgp.py:207-211 / 661-666
    res = solve_batch(gn, psf, bk, divergence=divergence, betaParam=betaParam, **kw)
    # A tile that was not restored must not be blended into the frame silently: with proj_type=1 a background-only
    # tile whose sum(gn - bkg) is not positive ends with BSGP_ST_BAD_FLUX (the reference raises ValueError at
    # sgp.py:269/713 for it).  on_failure="raise" (default) reports the tiles; "keep_input" substitutes the observed tile.
    status = res.status.cpu().numpy()
    bad = np.nonzero(status != 0)[0]
    if bad.size:
        from ._capi import STATUS_TEXT
        what = ", ".join(f"tile {int(i)} at {tuple(int(v) for v in org[i])}: {STATUS_TEXT.get(int(status[i]), status[i])}" for i in bad[:8])
        if on_failure == "raise":
            raise ValueError(f"{bad.size} of {len(status)} tiles were not restored ({what}); pass on_failure='keep_input' to keep the observed tiles there")
        if on_failure != "keep_input":
            raise ValueError("on_failure must be 'raise' or 'keep_input'")
        import warnings
        warnings.warn(f"{bad.size} of {len(status)} tiles were not restored and keep their observed pixels ({what})")
        idx = torch.as_tensor(bad, device=gn.device)
        res.x[idx] = gn[idx]
    out = reconstruct_full_image_from_patches(res.x, org, frame.shape, feather, device)
    return out, res, org
