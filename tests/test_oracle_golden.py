"""Pin the CPU oracle (oracle/cavour_oracle.py) to the reference: notebook values and
outputs of the unmodified reference engine (tests/golden/, see gen/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200.dates import Date
from tests.util_trades import METHOD, make_calibration_swaps, make_trade, leg_arrays, rel_err, trade_scales
from tests.conftest import GOLDEN

TOL = 1e-10  # north_star parity tolerance (relative, FP64)


def curve_inputs(cv):
    """swap_rates / swap_times / year_fracs rebuilt by the host layer, checked against
    the reference's (ois_curve.py:141-152)."""
    from adrates_b200.curves import OISCurve
    from adrates_b200.global_types import InterpTypes
    vd, swaps = make_calibration_swaps(cv)
    return vd, OISCurve(vd, swaps, InterpTypes[cv["interp"]])


_tables = {}


def tables_for(key, cv):
    if key not in _tables:
        plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
        d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
        _tables[key] = (plan["times"], d, J, C)
    return _tables[key]


def test_host_curve_inputs_match_reference(ref_curves):
    for key, cv in ref_curves.items():
        _, curve = curve_inputs(cv)
        assert curve.swap_rates == cv["swap_rates"], key
        assert curve.swap_times == cv["swap_times"], key
        assert curve.year_fracs == cv["year_fracs"], key


def test_path_b_dfs_bit_exact(ref_curves):
    for key, cv in ref_curves.items():
        plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
        assert np.array_equal(plan["times"], np.array(cv["pathB_times"])), key
        dfs = orc.bootstrap_dfs(cv["swap_rates"], plan)
        assert np.max(np.abs(dfs - np.array(cv["pathB_dfs"]))) <= 2e-16, key


@pytest.mark.parametrize("key", ["gbp_readme_lzr", "usd_dec24_lzr", "gbp_semi_lzr"])
def test_tangent_tables_match_reference_ad(ref_curves, key):
    ref = np.load(os.path.join(GOLDEN, f"ref_tables_{key}.npz"))
    _, d, J, C = tables_for(key, ref_curves[key])
    assert rel_err(J, ref["jac"], 1.0) < 1e-12
    assert rel_err(C, ref["hess"], 1.0) < 1e-12


def test_path_a_and_df_ad(ref_curves):
    for key, cv in ref_curves.items():
        t, d = orc.path_a_bootstrap(cv["swap_rates"], cv["swap_times"], cv["year_fracs"])
        assert np.array_equal(t, np.array(cv["pathA_times"])), key
        assert rel_err(d, cv["pathA_dfs"], 1.0) < 1e-15, key
        assert rel_err(orc.df_ad(np.array(cv["df_ad_t"]), t, d), cv["df_ad"], 1.0) < 1e-14, key


def test_notebook_golden_values(ref_curves, ref_trades):
    """notebooks/intro.ipynb cells 36/40/44: 1W PAY OIS at 5.2014% on the README curve."""
    cv = ref_curves["gbp_readme_lzr"]
    spec = next(t for t in ref_trades if t["id"] == "nb_1w_par")
    vd = Date(*cv["value_dt"])
    fixed, floating = leg_arrays(make_trade(spec, cv), vd)
    v, delta, gamma = orc.ois_analytics(tables_for("gbp_readme_lzr", cv), METHOD[cv["interp"]], fixed, floating)
    assert v == 4.672529030358419e-11                       # cell 36, bit-exact
    assert abs(delta[1] - 1.9158970567491282) < 1e-15        # cell 40 ('1W')
    assert np.all(np.delete(delta, 1) == 0.0)
    assert abs(gamma.sum() - (-7.34132e-06)) < 5e-12         # cell 44 (printed to 6 s.f.)


def test_trades_match_reference_engine(ref_curves, ref_trades):
    worst = {}
    for spec in ref_trades:
        cv = ref_curves[spec["curve"]]
        vd = Date(*cv["value_dt"])
        swap = make_trade(spec, cv)
        assert [[d.d(), d.m(), d.y()] for d in swap._fixed_leg._payment_dts] == spec["fixed_payment_dts"]
        assert swap._fixed_leg._payments == spec["fixed_payments"]
        assert swap._float_leg._year_fracs == spec["float_year_fracs"]
        fixed, floating = leg_arrays(swap, vd)
        v, delta, gamma = orc.ois_analytics(tables_for(spec["curve"], cv), METHOD[cv["interp"]], fixed, floating)
        s_pv, s_d, s_g = trade_scales(spec)
        e = (rel_err(v, spec["value"], s_pv), rel_err(delta, spec["delta"], s_d), rel_err(gamma, spec["gamma"], s_g))
        worst[spec["id"]] = e
        assert max(e) < TOL, (spec["id"], e)
        assert np.allclose(gamma, gamma.T, rtol=1e-10, atol=1e-14)
    assert len(worst) >= 27


def test_holiday_calendar_trades_match_reference_engine(ref_curves):
    """OIS rolled on UNITED_KINGDOM / TARGET / UNITED_STATES calendars (and the WEEKEND trades of the CASHFLOWS goldens), valued
    by the unmodified reference engine: the oracle on the host legs' dates pins holiday schedules through VALUE, delta, gamma."""
    from tests.util_cashflows import golden, make_cal_trade
    n = 0
    for spec in golden():
        if "error" in spec:
            continue
        cv = ref_curves[spec["curve"]]
        swap = make_cal_trade(spec, cv)
        assert [[d.d(), d.m(), d.y()] for d in swap._fixed_leg._payment_dts] == spec["fixed_payment_dts"]
        fixed, floating = leg_arrays(swap, Date(*cv["value_dt"]))
        v, delta, gamma = orc.ois_analytics(tables_for(spec["curve"], cv), METHOD[cv["interp"]], fixed, floating)
        s_pv, s_d, s_g = trade_scales(spec)
        e = (rel_err(v, spec["value"], s_pv), rel_err(delta, spec["delta"], s_d), rel_err(gamma, spec["gamma"], s_g))
        assert max(e) < TOL, (spec["id"], e)
        n += spec["cal"] != "WEEKEND"
    assert n == 4
