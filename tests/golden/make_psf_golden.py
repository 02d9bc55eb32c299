"""Golden vectors for the DIAPL PSF model: runs the UNMODIFIED ``PSF`` class of the reference (psf/psf_calculate.py:8-139;
its AST node is executed on its own because the module imports astropy / matplotlib, absent here) on the coefficient
file the reference ships, and reads the 31 x 31 image the reference itself wrote from that file
(psf/psfccfbrd210048_1_1_img.fits, BITPIX -64; parsed by hand, astropy is not available).  Only numbers are committed
(psf_golden.npz).  Run in the build container: python tests/golden/make_psf_golden.py
"""
import ast
import os

import numpy as np

PSF_DIR = "/root/reference/psf"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_class():
    src = os.path.join(PSF_DIR, "psf_calculate.py")
    tree = ast.parse(open(src).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PSF")
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), src, "exec"), ns)
    return ns["PSF"]


def read_fits_f64(path):
    raw = open(path, "rb").read()
    cards = [raw[i:i + 80].decode("ascii") for i in range(0, 2880, 80)]
    hdr = {c[:8].strip(): c[10:30].strip() for c in cards if "=" in c[:10]}
    assert hdr["BITPIX"] == "-64" and hdr["NAXIS"] == "2"
    nx, ny = int(hdr["NAXIS1"]), int(hdr["NAXIS2"])
    return np.frombuffer(raw[2880:2880 + 8 * nx * ny], dtype=">f8").reshape(ny, nx).astype(np.float64)


if __name__ == "__main__":
    txt = os.path.join(PSF_DIR, "psfccfbrd210048_1_1.bin.txt")
    PSF = reference_class()
    p = PSF(txt)
    mat = p.get_psf_mat().copy()
    norm = p.normalize_psf_mat().copy()
    shipped = read_fits_f64(os.path.join(PSF_DIR, "psfccfbrd210048_1_1_img.fits"))
    print("reference class vs the image it shipped: max |diff| =", np.abs(norm - shipped).max(), " sum =", shipped.sum())
    lines = [float(l) for l in open(txt).read().split()]
    np.savez(os.path.join(HERE, "psf_golden.npz"), file_values=np.array(lines), mat=mat, norm=norm, shipped=shipped)
