"""The oracle (oracle/sgp_oracle.py) against the committed vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  make_golden asserted bit equality when it ran; here a
tight tolerance is used so that a different CPU (other SIMD code paths in numpy's pow/log) cannot
fail the suite spuriously."""
import numpy as np
import pytest

from cases import CASES
from oracle import sgp_oracle as orc

FAST = [n for n in CASES if not n.endswith("_332")]


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference(name, get_case, golden):
    gn, psf, bkg, div, kw = get_case(name)
    r = orc.solve(gn.copy(), psf.copy(), bkg, divergence=div, **kw)
    assert r.iters == int(golden[name + "/iters"])
    np.testing.assert_allclose(r.discr, golden[name + "/discr"], rtol=1e-9)
    np.testing.assert_allclose(r.x[::8, ::8], golden[name + "/x_sub"], rtol=0, atol=1e-7 * np.abs(golden[name + "/x_sub"]).max())
    if name + "/x" in golden.files:
        np.testing.assert_allclose(r.x, golden[name + "/x"], rtol=0, atol=1e-7 * np.abs(golden[name + "/x"]).max())
    assert np.array_equal(np.array(r.trace.proj_evals), golden[name + "/proj_evals"])
    if div == "beta":
        assert r.beta_param == pytest.approx(float(golden[name + "/beta_final"]), rel=1e-12)


def test_oracle_long_run_rel_err(get_case, golden, fixtures):
    """simulation_test_sgp.py:37-54: KL-SGP on the satellite simulation at the SGP-dec-optimal 332 iterations;
    the reference only returns the relative reconstruction error."""
    gn, psf, bkg, div, kw = get_case("sat_kl_332")
    r = orc.solve(gn.copy(), psf.copy(), bkg, divergence=div, **kw)
    obj = fixtures["sat/obj"]
    rel = np.sqrt(np.sum((r.x - obj) ** 2) / np.sum(obj * obj))
    assert r.iters == 332
    assert rel == pytest.approx(float(golden["sat_kl_332/rel_err"]), rel=1e-6)


def test_oracle_projection_known_answers(golden):
    for i in range(12):
        k = f"proj{i:02d}"
        sat = float(golden[k + "/sat"])
        cnt = []
        x = orc.flux_projection(np.float64(golden[k + "/b"]), golden[k + "/c"].copy(), golden[k + "/dia"].copy(),
                                float(golden[k + "/scaling"]), ccd_sat_level=None if np.isnan(sat) else sat, counter=cnt)
        np.testing.assert_allclose(x, golden[k + "/x"], rtol=1e-12, atol=1e-15)
        assert cnt[-1] == int(golden[k + "/evals"])
        assert abs(x.sum() - float(golden[k + "/b"])) <= 1e-9 * float(golden[k + "/b"])


def test_oracle_beta_divergence_helpers(golden):
    """tests.py:9-19 and :54-68 restated without torchnmf: closed form of the beta-divergence and a central
    difference of it in beta."""
    x, y = golden["betadiv/x"], golden["betadiv/y"]
    b = 1.5
    closed = np.sum((x ** b + (b - 1) * y ** b - b * x * y ** (b - 1)) / (b * (b - 1)))
    assert np.isclose(orc.beta_divergence(y, x, b), closed)
    assert np.isclose(orc.beta_divergence(y, x, b), float(golden["betadiv/value_1p5"]))
    h = 1e-6
    fd = (orc.beta_divergence(y, x, 1.7 + h) - orc.beta_divergence(y, x, 1.7 - h)) / (2 * h)
    assert np.isclose(orc.beta_divergence_dbeta(y, x, 1.7).sum(), fd, rtol=1e-6)
    np.testing.assert_allclose(orc.beta_divergence_dbeta(y, x, 1.7), golden["betadiv/deriv_1p7"], rtol=1e-13)
    # KL special case (tests.py:21-52): beta = 1 gradient equals the KL gradient
    op = orc.CircularPsf(np.full((8, 8), 1 / 64.0))
    den = np.random.default_rng(0).uniform(1, 2, 64)
    gnv = np.random.default_rng(1).uniform(1, 2, 64)
    np.testing.assert_allclose(orc.beta_divergence_grad(op.adjoint, den, gnv, 1), 1 - op.adjoint(gnv / den))


def test_oracle_error_behaviour(fixtures):
    gn, psf = fixtures["ngc/gn"], fixtures["ngc/psf"]
    with pytest.raises(ValueError, match="PSF is not normalized"):
        orc.solve(gn, psf * 1.01, np.float64(1.0), MAXIT=2)
    with pytest.raises(ValueError, match="errflag"):
        orc.solve(gn, psf, np.float64(1.0), MAXIT=2, errflag=True)
