"""Non-AD host methods next to the path against known answers of the unmodified reference (tests/golden/ref_host_values.json,
tests/golden/gen/make_golden_host_values.py): YoYInflationSwap.value / breakeven_rate / pv01 and the per-payment lists of
SwapYoYInflationLeg.value (incl. the CPI look-up error of sub-annual / seasoned swaps), ZeroCouponInflationSwap.pv01,
OIS.value / pv01 / swap_rate / ir01 on the path-A curve."""
import numpy as np
import pytest

from adrates_b200 import (BusDayAdjustTypes, CurrencyTypes, CurveTypes, Date, DayCountTypes, FrequencyTypes, LibError, OIS, SwapTypes,
                          ZeroCouponInflationSwap)
from tests.conftest import load_golden
from tests.util_yoy import make_model, make_swap

TOL = 1e-12


def test_yoy_host_values_match_reference():
    g = load_golden("ref_yoy.json")
    ref = {r["id"]: r for r in load_golden("ref_host_values.json")["yoy"]}
    n_err = 0
    for name in g["inflation_curves"]:
        model, idx, ic = make_model(g, name)
        disc, vd = model.curves.GBP_OIS_SONIA, model.value_dt
        for c in (c for c in g["cases"] if c["index"] == name):
            r, sw, N = ref[c["id"]], make_swap(c, idx), c["notional"]
            assert abs(sw.pv01(vd, disc) - r["pv01"]) <= TOL * N
            if "error" in r:
                with pytest.raises(LibError) as ex:
                    sw.value(vd, disc, ic)
                assert "LibError: " + str(ex.value) == r["error"]
                n_err += 1
                continue
            assert abs(sw.value(vd, disc, ic) - r["value"]) <= TOL * N
            assert abs(sw._fixed_pv - r["fixed_pv"]) <= TOL * N and abs(sw._inflation_pv - r["inflation_pv"]) <= TOL * N
            leg = sw._inflation_leg
            for got, key, scale in ((leg._start_cpis, "start_cpis", 300.0), (leg._end_cpis, "end_cpis", 300.0),
                                    (leg._yoy_rates, "yoy_rates", 1.0), (leg._payments, "payments", N), (leg._dfs, "dfs", 1.0),
                                    (leg._pvs, "pvs", N)):
                assert np.max(np.abs(np.array(got) - np.array(r[key]))) <= TOL * scale, (c["id"], key)
            assert abs(sw.breakeven_rate(vd, disc, ic) - r["breakeven"]) <= TOL
    assert n_err == 8


def test_zcis_pv01_and_ois_host_values_match_reference(ref_curves):
    from tests.util_trades import build_model
    from tests.util_zcis import make_index
    g = load_golden("ref_host_values.json")
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve, vd = model.curves[cv["name"]], model.value_dt
    idx = make_index(load_golden("ref_zcis.json"), "rpi_linear")
    for r in g["zcis_pv01"]:
        if r["tenor"] == "matured":
            z = ZeroCouponInflationSwap(Date(30, 4, 2019), "5Y", SwapTypes.RECEIVE, 0.03, idx, 1_000_000)
            assert z.pv01(vd, curve) == 0.0 == r["pv01"]
            continue
        z = ZeroCouponInflationSwap(vd, r["tenor"], SwapTypes.PAY, r["rate"], idx, 2_500_000)
        assert abs(z.pv01(vd, curve) - r["pv01"]) <= TOL * 2_500_000
    for r in g["ois"]:
        sw = OIS(effective_dt=vd, term_dt_or_tenor=r["tenor"], fixed_leg_type=SwapTypes[r["side"]], fixed_coupon=r["coupon"],
                 fixed_freq_type=FrequencyTypes[r["fixed_freq"]], fixed_dc_type=DayCountTypes.ACT_365F,
                 floating_index=CurveTypes.GBP_OIS_SONIA, currency=CurrencyTypes.GBP, notional=5e6,
                 float_freq_type=FrequencyTypes[r["float_freq"]], float_dc_type=DayCountTypes.ACT_365F,
                 bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        assert abs(sw.value(vd, curve) - r["value"]) <= TOL * 5e6
        assert abs(sw.pv01(vd, curve) - r["pv01"]) <= TOL * 5e6
        assert abs(sw.swap_rate(vd, curve) - r["swap_rate"]) <= TOL
        assert abs(sw.ir01(vd, curve) - r["ir01"]) <= 1e-11 * 5e6
