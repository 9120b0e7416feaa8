"""Cross-currency path on CPU: XccyCurve bootstrap + Jacobian against the reference's AD tables, and the flattened
legs (evaluated with the numpy restatement of the kernels) against the reference engine's VALUE and three deltas."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200.flatten import Flattener
from adrates_b200.xccy_engine import domestic_leg_unit, flatten_foreign_legs
from tests.conftest import load_golden
from tests.flat_eval import eval_flat
from tests.util_xccy import build_xccy_model, make_xccy_trade

TOL = 1e-10


@pytest.fixture(scope="module")
def xg():
    return load_golden("ref_xccy.json")


def test_xccy_curve_matches_reference(xg):
    m = build_xccy_model(xg)
    xc = m.curves.GBP_USD_BASIS
    assert np.array_equal(xc._times, np.array(xg["xccy_times"]))
    assert np.max(np.abs(xc._dfs - np.array(xg["xccy_dfs"]))) < 1e-14
    J = np.array(xg["xccy_jac_basis"])
    assert np.max(np.abs(xc._jac_basis - J)) < 1e-12 * np.abs(J).max()
    assert xc.swap_times == xg["xccy_swap_times"] and xc._spot_fx == xg["xccy_spot_fx_internal"]
    assert xc._interp_type.name == xg["xccy_interp"]


def test_calibration_basis_swaps_reprice(xg):
    """tests/test_xccy_curve.py:213 / test_xccy_simple.py:131 of the reference: abs(PV / N) < 1e-8 through the non-AD value(),
    and the reference's refit check at its SWAP_TOL.  Holds when the OIS curves interpolate flat-forward (the bootstrap projects
    forwards log-linearly); with LINEAR_ZERO_RATES curves the check raises, as the reference's does."""
    from adrates_b200 import InterpTypes, LibError
    import tests.util_xccy as ux
    g = dict(xg)
    m = ux.build_xccy_model(g, ois_interp=InterpTypes.FLAT_FWD_RATES)
    xc = m.curves.GBP_USD_BASIS
    assert max(abs(r) for r in xc.par_residuals()) < 1e-12
    xc._check_refits(1e-10)
    lz = ux.build_xccy_model(dict(xg), ois_interp=InterpTypes.LINEAR_ZERO_RATES).curves.GBP_USD_BASIS
    with pytest.raises(LibError, match="XCCY swap with maturity .* not repriced. Difference is"):
        lz._check_refits(1e-10)
    sw = xc._used_swaps[0]
    kw = dict(value_dt=m.value_dt, domestic_discount_curve=xc._domestic_curve, foreign_discount_curve=xc._foreign_curve)
    v = sw.value(xccy_discount_curve=xc, spot_fx=xc._spot_fx, **kw)
    pv_dom = sw._domestic_leg.value(m.value_dt, xc._domestic_curve, xc._domestic_curve)
    pv_for = sw._foreign_leg.value(m.value_dt, xc, xc._foreign_curve)
    assert v == pv_dom + pv_for / xc._spot_fx
    with pytest.raises(ValueError, match="xccy_discount_curve required for domestic collateral"):
        sw.value(spot_fx=xc._spot_fx, **kw)
    from adrates_b200 import CollateralType
    foreign = CollateralType[sw._foreign_currency.name]
    with pytest.raises(ValueError, match="xccy_discount_curve_inverted required for foreign collateral"):
        sw.value(xccy_discount_curve=xc, spot_fx=xc._spot_fx, collateral_type=foreign, **kw)
    vf = sw.value(xccy_discount_curve_inverted=xc, spot_fx=xc._spot_fx, collateral_type=foreign, **kw)
    assert vf == sw._domestic_leg.value(m.value_dt, xc, xc._domestic_curve) * xc._spot_fx + \
        sw._foreign_leg.value(m.value_dt, xc._foreign_curve, xc._foreign_curve)


def test_flattened_xccy_trades_match_reference_engine(xg):
    m = build_xccy_model(xg)
    dom, forn, xc = m.curves.USD_OIS_SOFR, m.curves.GBP_OIS_SONIA, m.curves.GBP_USD_BASIS
    vd = m.value_dt
    tabs = {}
    for c in (dom, forn):
        plan = orc.plan_path_b(c.swap_times, c.year_fracs)
        tabs[id(c)] = orc.bootstrap_tables(c.swap_rates, plan)
    d_d, J_d, _ = tabs[id(dom)]
    d_f, J_f, _ = tabs[id(forn)]
    d_x, J_b = xc._dfs, xc._jac_basis
    d_s = np.concatenate([d_f, d_x])
    Gf, Gx, Rb = len(d_f), len(d_x), J_b.shape[1]
    J_for = np.vstack([J_f, np.zeros((Gx, 32))])
    J_bas = np.vstack([np.zeros((Gf, Rb)), J_b])
    z = lambda G, R: np.zeros((G, R, R))  # noqa: E731
    for t in xg["trades"]:
        sw = make_xccy_trade(t)
        fl = Flattener(dom)
        fl.add_components([(("XD", 0), domestic_leg_unit(sw, vd), 1.0)])
        pv_d, dl_d, _ = eval_flat(fl.finalize(dedup=False), d_d, J_d, z(len(d_d), 32))
        flat = flatten_foreign_legs([sw], vd, forn, xc)
        pv_f, dl_f, _ = eval_flat(flat, d_s, J_for, z(Gf + Gx, 32))
        _, dl_b, _ = eval_flat(flat, d_s, J_bas, z(Gf + Gx, Rb))
        N = t["domestic_notional"]
        T = float(t["tenor"][:-1])
        assert abs(pv_d[0] + pv_f[0] - t["value"]) <= TOL * max(abs(t["value"]), N), t["id"]
        ref = t["deltas"]
        for got, key in ((dl_d[0], "USD_OIS_SOFR"), (dl_f[0], "GBP_OIS_SONIA"), (dl_b[0], "USD_GBP_BASIS")):
            r = np.array(ref[key]["ladder"])
            assert np.max(np.abs(got[:len(r)] - r) / np.maximum(np.abs(r), N * 1e-4 * T)) < TOL, (t["id"], key)


def test_ois_with_usd_collateral_matches_reference_engine(xg):
    """Engine._compute_ois_xccy_collateral (engine.py:217-503): VALUE in collateral currency + OIS and basis ladders."""
    from adrates_b200 import Date
    from adrates_b200.xccy_engine import ois_collateral_terms, _flatten_stacked
    from tests.util_trades import make_trade
    cg = load_golden("ref_collateral.json")
    m = build_xccy_model(xg, xccy_name="GBP_USD_XCCY")
    ois, xc = m.curves.GBP_OIS_SONIA, m.curves.GBP_USD_XCCY
    plan = orc.plan_path_b(ois.swap_times, ois.year_fracs)
    d_o, J_o, _ = orc.bootstrap_tables(ois.swap_rates, plan)
    d_s = np.concatenate([d_o, xc._dfs])
    Go, Gx, Rb = len(d_o), len(xc._dfs), xc._jac_basis.shape[1]
    cvspec = {"name": "GBP_OIS_SONIA", "dc": "ACT_365F"}
    for t in cg["trades"]:
        sw = make_trade(dict(t, payment_lag=0), cvspec)
        flat = _flatten_stacked([ois_collateral_terms(sw, m.value_dt, xc._spot_fx)], ois, xc)
        pv, dl_o, _ = eval_flat(flat, d_s, np.vstack([J_o, np.zeros((Gx, 32))]), np.zeros((Go + Gx, 32, 32)))
        _, dl_b, _ = eval_flat(flat, d_s, np.vstack([np.zeros((Go, Rb)), xc._jac_basis]), np.zeros((Go + Gx, Rb, Rb)))
        N, T = t["notional"] / xc._spot_fx, float(t["tenor"][:-1])
        assert t["currency"] == "USD"
        assert abs(pv[0] - t["value"]) <= TOL * max(abs(t["value"]), N), t["id"]
        for got, key in ((dl_o[0], "GBP_OIS_SONIA"), (dl_b[0], "USD_GBP_BASIS")):
            r = np.array(t["deltas"][key]["ladder"])
            assert np.max(np.abs(got[:len(r)] - r) / np.maximum(np.abs(r), N * 1e-4 * T)) < TOL, (t["id"], key)
