"""Import the UNMODIFIED reference (``/root/reference/restoration``) in the build container.

``sgp.py`` imports astropy / photutils / matplotlib / ``utils`` at module level (sgp.py:14-31); none of
them is touched by ``sgp()`` / ``sgp_betaDiv()`` with ``use_original_SGP_Afunction=True, save=False``.
They are absent here, so placeholder modules are registered before the import.  The reference writes
``./sgp.log``; callers should chdir to a scratch directory.  This file is only used by
``make_golden.py`` and by the optional oracle-vs-reference test; it never runs on the GPU box
(``/root/reference`` does not exist there).
"""
import os
import sys
import types

REFERENCE_DIR = "/root/reference/restoration"
_STUBS = ["astropy", "astropy.units", "astropy.io", "astropy.io.fits", "astropy.wcs", "astropy.wcs.utils",
          "astropy.nddata", "astropy.stats", "astropy.coordinates", "astropy.convolution", "photutils",
          "photutils.background", "photutils.segmentation", "matplotlib", "matplotlib.pyplot", "utils"]


class _Placeholder(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _missing(*a, **k):
            raise RuntimeError(f"{self.__name__}.{name} is not available offline")
        return _missing


def available():
    return os.path.isdir(REFERENCE_DIR)


def load():
    """Returns (sgp_module, flux_conserve_proj_module) of the unmodified reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    saved = {}
    for n in _STUBS:
        if n not in sys.modules:
            saved[n] = None
            sys.modules[n] = _Placeholder(n)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    # the product package also ships modules called `sgp` / `flux_conserve_proj`; make sure we get
    # the reference's
    for n in ("sgp", "flux_conserve_proj"):
        m = sys.modules.get(n)
        if m is not None and not getattr(m, "__file__", "").startswith(REFERENCE_DIR):
            del sys.modules[n]
    import flux_conserve_proj as ref_proj
    import sgp as ref_sgp
    assert ref_sgp.__file__.startswith(REFERENCE_DIR), ref_sgp.__file__
    return ref_sgp, ref_proj
