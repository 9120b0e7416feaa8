"""Second-order tables of the cross-currency path on the CPU.  The reference gets them by nesting JAX transforms through
its bootstrap scan (xccy_curve.py:594-690) and its own XCCY GAMMA request never returns (engine.py:1936-1939 raises), so
there are no goldens: following SURVEY R8 the tables are validated by finite differences of the FIRST-order tables, which are
pinned to the reference's AD output (tests/test_xccy_cpu.py)."""
import copy

import numpy as np
import pytest

from adrates_b200.dual2 import D2, exp, interp, log
from adrates_b200.global_types import CurrencyTypes, CurveTypes
from adrates_b200.results import CrossGamma, Gamma, Risk
from tests.conftest import load_golden
from tests.util_xccy import build_xccy_model


@pytest.fixture(scope="module")
def model():
    return build_xccy_model(load_golden("ref_xccy.json"))


def test_dual2_algebra_against_finite_differences():
    rng = np.random.default_rng(0)
    x0 = rng.uniform(0.5, 1.5, 4)

    def f(x):
        a, b, c, d = x
        return exp(a * b - c / d) * log(a + 2.0 * d) / (1.0 + b * b) - interp(1.3, np.array([0.0, 1.0, 2.0]), [a, b * c, d])
    val = f([D2.var(v, i, 4) for i, v in enumerate(x0)])
    h = 1e-5
    for i in range(4):
        e = np.zeros(4)
        e[i] = h
        assert abs((f(x0 + e) - f(x0 - e)) / (2 * h) - val.g[i]) < 1e-8
        for j in range(4):
            ej = np.zeros(4)
            ej[j] = h
            fd = (f(x0 + e + ej) - f(x0 + e - ej) - f(x0 - e + ej) + f(x0 - e - ej)) / (4 * h * h)
            assert abs(fd - val.h[i, j]) < 1e-5
    assert abs(f(x0) - val.v) < 1e-15


def _rebuilt(xc, spreads):
    c = copy.copy(xc)
    c.basis_spreads = list(spreads)
    c._bootstrap()
    return c


def test_hess_basis_is_the_derivative_of_the_pinned_jacobian(model):
    xc = model.curves.GBP_USD_BASIS
    H = xc._hess_basis
    nb = len(xc.basis_spreads)
    assert H.shape == (len(xc._dfs), nb, nb)
    assert np.allclose(H, np.swapaxes(H, 1, 2), rtol=0, atol=1e-9 * np.abs(H).max())
    h = 1e-6
    for k in range(nb):
        up, dn = list(xc.basis_spreads), list(xc.basis_spreads)
        up[k] += h
        dn[k] -= h
        fd = (_rebuilt(xc, up)._jac_basis - _rebuilt(xc, dn)._jac_basis) / (2 * h)       # d J[:, :] / d spread_k
        assert np.max(np.abs(fd - H[:, :, k])) < 1e-6 * max(np.abs(H).max(), 1.0), k


def test_foreign_node_tables_against_finite_differences(model):
    """_jac_foreign_curve_dfs and _mixed_hess_foreign_basis: the scan on plain floats with payment-time foreign DFs taken by
    log-linear interpolation of bumped node DFs (the function the reference differentiates, xccy_curve.py:640-660)."""
    xc = model.curves.GBP_USD_BASIS
    fx = np.asarray(xc._foreign_curve._times)
    fd0 = np.asarray(xc._foreign_curve._dfs)
    idx = xc._node_idx

    def dfs(node_dfs, spreads):
        lg = np.log(node_dfs)
        out = xc._scan(list(spreads), [float(np.exp(np.interp(p["time"], fx, lg))) for p in xc._pts])
        return np.array([1.0] + [out[i] for i in idx])
    Jf, M = xc._jac_foreign_curve_dfs, xc._mixed_hess_foreign_basis
    assert Jf.shape == (len(xc._dfs), len(fd0)) and M.shape == (len(xc._dfs), len(xc.basis_spreads), len(fd0))
    h, hb = 1e-6, 1e-6
    b0 = np.array(xc.basis_spreads)
    for j in (5, 20, 40, len(fd0) - 1):
        e = np.zeros_like(fd0)
        e[j] = h
        fdj = (dfs(fd0 + e, b0) - dfs(fd0 - e, b0)) / (2 * h)
        assert np.max(np.abs(fdj - Jf[:, j])) < 1e-7 * max(np.abs(Jf).max(), 1.0), j
        for k in (0, len(b0) - 1):
            eb = np.zeros_like(b0)
            eb[k] = hb
            mixed = (dfs(fd0 + e, b0 + eb) - dfs(fd0 + e, b0 - eb) - dfs(fd0 - e, b0 + eb) + dfs(fd0 - e, b0 - eb)) / (4 * h * hb)
            assert np.max(np.abs(mixed - M[:, k, j])) < 2e-4 * max(np.abs(M).max(), 1.0), (k, j)


def test_path_a_jacobian_against_finite_differences(model):
    c = model.curves.GBP_OIS_SONIA
    J = c.path_a_jacobian()
    assert J.shape == (len(c._dfs), len(c.swap_rates)) and np.all(J[0] == 0.0)
    h = 1e-7
    for k in (0, 14, 20, 31):
        up, dn = copy.copy(c), copy.copy(c)
        up.swap_rates, dn.swap_rates = list(c.swap_rates), list(c.swap_rates)
        up.swap_rates[k] += h
        dn.swap_rates[k] -= h
        up._bootstrap_path_a()
        dn._bootstrap_path_a()
        assert np.max(np.abs((up._dfs - dn._dfs) / (2 * h) - J[:, k])) < 1e-6 * np.abs(J).max(), k


def test_cross_gamma_container():
    """cavour/requests/results.py:608-836, 839-942: validation, value, dict / frame / JSON export, addition, Risk access."""
    t1, t2 = ["1Y", "2Y", "5Y"], ["1Y", "10Y"]
    m = np.arange(6.0).reshape(3, 2)
    cg = CrossGamma(m, t1, t2, CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS, CurrencyTypes.USD)
    assert cg.value.amount == 15.0 and cg.value.currency is CurrencyTypes.USD
    assert cg.to_dict["5Y"]["10Y"] == 5.0 and list(cg.df.columns) == t2 and '"total": 15.0' in cg.to_json()
    assert (cg + cg).risk_matrix[2, 1] == 10.0 and "shape=[3, 2]" in repr(cg)
    with pytest.raises(ValueError):
        CrossGamma(m, t1, ["1Y"], CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS, CurrencyTypes.USD)
    with pytest.raises(ValueError):
        CrossGamma(np.zeros(3), t1, t2, CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS, CurrencyTypes.USD)
    with pytest.raises(TypeError):
        CrossGamma(m, t1, t2, "GBP_OIS_SONIA", CurveTypes.USD_GBP_BASIS, CurrencyTypes.USD)
    with pytest.raises(ValueError):
        cg + CrossGamma(m, t1, t2, CurveTypes.USD_OIS_SOFR, CurveTypes.USD_GBP_BASIS, CurrencyTypes.USD)
    g = Gamma(np.eye(2), t2, CurrencyTypes.USD, CurveTypes.USD_GBP_BASIS)
    r = Risk([g], cross_gammas=[cg])
    assert r.cross_gamma(CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS) is cg
    assert r.cross_gamma(CurveTypes.USD_GBP_BASIS, CurveTypes.GBP_OIS_SONIA) is None
    assert r.USD_GBP_BASIS is g and r(CurveTypes.USD_GBP_BASIS) is g
    with pytest.raises(ValueError):
        r(CurveTypes.GBP_OIS_SONIA)
    with pytest.raises(ValueError):
        Risk([g], cross_gammas=[cg, cg])
