"""XCCY GAMMA on the GPU: Risk([Gamma x 3], cross_gammas=[CrossGamma]) of Engine._compute_xccy (engine.py:1769-1967).
The reference's own request raises inside its cross-gamma contraction (reproduced when the goldens were generated), so there
is nothing to compare with; following SURVEY R8 every block is validated by central finite differences of the GPU delta
ladders - which ARE pinned to the reference engine (tests/test_gpu_parity.py) - under the reference's bump convention (the
other two curves held fixed), and the cross block against the explicit numpy contraction of its formula."""
import copy
from types import SimpleNamespace

import numpy as np
import pytest

from adrates_b200 import CurveTypes, RequestTypes
from adrates_b200.xccy_engine import compute_xccy, flatten_foreign_legs
from tests.conftest import load_golden
from tests.util_xccy import build_xccy_model, make_xccy_trade

pytestmark = pytest.mark.gpu
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
VD = [RequestTypes.VALUE, RequestTypes.DELTA]


def _bumped_ois(curve, k, h):
    """The same OISCurve with par rate k moved by h (decimal): same dates, new path-A nodes, new engine tables."""
    c = copy.copy(curve)
    c.swap_rates = list(curve.swap_rates)
    c.swap_rates[k] += h
    c._jac_path_a = None
    c._bootstrap_path_a()
    return c


def _bumped_basis(xc, k, h):
    c = copy.copy(xc)
    c.basis_spreads = list(xc.basis_spreads)
    c.basis_spreads[k] += h
    c._bootstrap()
    return c


def _model(m, **curves):
    ns = SimpleNamespace(USD_OIS_SOFR=m.curves.USD_OIS_SOFR, GBP_OIS_SONIA=m.curves.GBP_OIS_SONIA, GBP_USD_BASIS=m.curves.GBP_USD_BASIS)
    for k, v in curves.items():
        setattr(ns, k, v)
    return SimpleNamespace(curves=ns, value_dt=m.value_dt)


@pytest.fixture(scope="module")
def setup():
    g = load_golden("ref_xccy.json")
    m = build_xccy_model(g)
    trades = [make_xccy_trade(t) for t in g["trades"][:4]]
    return g, m, trades


def test_xccy_gamma_blocks_match_finite_differences_of_the_ladders(setup):
    g, m, trades = setup
    res = compute_xccy(trades, m, ALL)
    dom_i, for_i = trades[0]._domestic_floating_index, trades[0]._foreign_floating_index
    G_dom, G_for, G_bas = res.gamma(dom_i), res.gamma(for_i), res.gamma.USD_GBP_BASIS
    assert G_dom.risk_ladder.shape == (32, 32) and G_for.risk_ladder.shape == (32, 32)
    nb = len(m.curves.GBP_USD_BASIS.basis_spreads)
    assert G_bas.risk_ladder.shape == (nb, nb)
    for Gm in (G_dom, G_for, G_bas):
        assert np.allclose(Gm.risk_ladder, Gm.risk_ladder.T, rtol=0, atol=1e-9 * np.abs(Gm.risk_ladder).max())
        assert Gm.value.amount == float(np.sum(Gm.risk_ladder))
    # the ladders of the GAMMA request are those of the VALUE + DELTA request
    ref = compute_xccy(trades, m, VD)
    for ct in (dom_i, for_i, CurveTypes.USD_GBP_BASIS):
        assert np.allclose(res.risk(ct).risk_ladder, ref.risk(ct).risk_ladder, rtol=1e-12, atol=1e-9)
    assert abs(res.value.amount - ref.value.amount) <= 1e-9 * abs(ref.value.amount)
    h = 1e-6

    def fd(name, bump, ct, pillars):
        out = {}
        for k in pillars:
            up = compute_xccy(trades, _model(m, **{name: bump(k, +h)}), VD).risk(ct).risk_ladder
            dn = compute_xccy(trades, _model(m, **{name: bump(k, -h)}), VD).risk(ct).risk_ladder
            out[k] = (up - dn) / (2 * h) * 1e-4          # d(per-bp ladder)/d(rate k), per bp
        return out
    dom_c, for_c, xc = m.curves.USD_OIS_SOFR, m.curves.GBP_OIS_SONIA, m.curves.GBP_USD_BASIS
    dname, fname = dom_i.name, for_i.name
    for k, col in fd(dname, lambda k, e: _bumped_ois(getattr(m.curves, dname), k, e), dom_i, (14, 19, 24)).items():
        assert np.max(np.abs(col - G_dom.risk_ladder[:, k])) <= 2e-5 * np.abs(G_dom.risk_ladder).max(), k
    for k, col in fd(fname, lambda k, e: _bumped_ois(getattr(m.curves, fname), k, e), for_i, (14, 19, 24)).items():
        assert np.max(np.abs(col - G_for.risk_ladder[:, k])) <= 2e-5 * np.abs(G_for.risk_ladder).max(), k
    for k, col in fd("GBP_USD_BASIS", lambda k, e: _bumped_basis(xc, k, e), CurveTypes.USD_GBP_BASIS, (0, nb // 2, nb - 1)).items():
        assert np.max(np.abs(col - G_bas.risk_ladder[:, k])) <= 2e-5 * np.abs(G_bas.risk_ladder).max(), k
    assert dom_c is not for_c


def test_xccy_cross_gamma_is_the_contraction_of_its_formula(setup):
    """result[l, k] = 1e-8 * sum_i dPV/d(xccy node DF i) * sum_j mixed[i, k, j] * d(foreign path-A node DF j)/d(rate l)."""
    g, m, trades = setup
    res = compute_xccy(trades, m, ALL)
    for_i = trades[0]._foreign_floating_index
    cg = res.gamma.cross_gamma(for_i, CurveTypes.USD_GBP_BASIS)
    forn, xc = getattr(m.curves, for_i.name), m.curves.GBP_USD_BASIS
    assert cg is not None and cg.risk_matrix.shape == (32, len(xc.basis_spreads))
    assert res.gamma.cross_gamma(CurveTypes.USD_GBP_BASIS, for_i) is None
    assert "reference raises" in res.gamma.beyond_reference
    # gradient of the foreign-leg PV w.r.t. the XCCY node DFs from the flat terms: PV = sum amt exp(sum w ln d[node])
    flat = flatten_foreign_legs(trades, m.value_dt, forn, xc)
    from oracle import cavour_oracle as orc
    d_f, _, _ = orc.bootstrap_tables(forn.swap_rates, orc.plan_path_b(forn.swap_times, forn.year_fracs))
    d = np.concatenate([d_f, xc._dfs])
    w, nd = flat.weight.reshape(-1, 6), flat.node.reshape(-1, 6)
    p = flat.amt * np.exp(np.sum(w * np.log(d[nd]), axis=1))
    grad = np.zeros(d.shape[0])
    np.add.at(grad, nd.reshape(-1), (p[:, None] * w / d[nd]).reshape(-1))
    grad_x = grad[len(d_f):]
    A = np.einsum("ikj,jl->ilk", xc._mixed_hess_foreign_basis, forn.path_a_jacobian())
    want = 1e-8 * np.einsum("i,ilk->lk", grad_x, A)
    assert np.max(np.abs(cg.risk_matrix - want)) <= 1e-10 * max(np.abs(want).max(), 1e-30)
    assert np.abs(want).max() > 0.0
