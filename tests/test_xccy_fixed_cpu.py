"""XccyFixFloat.value / XccyFixFix.value (adrates_b200/trades.py; reference xccy_fix_float_swap.py:196-245,
xccy_fix_fix_swap.py:210-280) against known answers of the unmodified reference on the market of its own tests
(tests/golden/ref_xccy_fixed.json, tests/golden/gen/make_golden_xccy_fixed.py): spot-starting, forward-starting and seasoned
swaps, both sides, a known first fixing; the bootstrapped basis curve itself is compared node for node."""
import numpy as np
import pytest

from adrates_b200 import (BusDayAdjustTypes, CurrencyTypes, CurveTypes, Date, DayCountTypes, FrequencyTypes, InterpTypes, LibError, Model,
                          SwapTypes, XccyBasisSwap, XccyCurve, XccyFixFix, XccyFixFloat)
from tests.conftest import load_golden


@pytest.fixture(scope="module")
def market():
    g = load_golden("ref_xccy_fixed.json")
    vd = Date(*g["value_dt"])
    curves = {}
    for name, px, dc in (("GBP_OIS_SONIA", g["gbp"], DayCountTypes.ACT_365F), ("USD_OIS_SOFR", g["usd"], DayCountTypes.ACT_360)):
        m = Model(vd)
        m.build_curve(name=name, px_list=px, tenor_list=g["tenors"], spot_days=0, swap_type=SwapTypes.PAY, fixed_dcc_type=dc,
                      fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                      bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.FLAT_FWD_RATES)
        curves[name] = m.curves[name]
    gbp, usd = curves["GBP_OIS_SONIA"], curves["USD_OIS_SOFR"]
    basis = [XccyBasisSwap(effective_dt=vd, term_dt_or_tenor=t, domestic_notional=g["spot"] * 1_000_000, foreign_notional=1_000_000,
                           domestic_spread=0.0, foreign_spread=s, domestic_freq_type=FrequencyTypes.ANNUAL,
                           foreign_freq_type=FrequencyTypes.ANNUAL, domestic_dc_type=DayCountTypes.ACT_365F,
                           foreign_dc_type=DayCountTypes.ACT_360, domestic_floating_index=CurveTypes.GBP_OIS_SONIA,
                           foreign_floating_index=CurveTypes.USD_OIS_SOFR, domestic_currency=CurrencyTypes.GBP,
                           foreign_currency=CurrencyTypes.USD) for t, s in zip(g["tenors"], g["basis"])]
    xc = XccyCurve(value_dt=vd, basis_swaps=basis, domestic_curve=gbp, foreign_curve=usd, spot_fx=g["spot"],
                   interp_type=InterpTypes.FLAT_FWD_RATES, check_refit=True)      # flat-forward OIS curves: the refit check passes
    return g, vd, gbp, usd, xc


COMMON = dict(domestic_notional=790_000, foreign_notional=1_000_000, domestic_dc_type=DayCountTypes.ACT_365F,
              foreign_dc_type=DayCountTypes.ACT_360, domestic_floating_index=CurveTypes.GBP_OIS_SONIA,
              foreign_floating_index=CurveTypes.USD_OIS_SOFR, domestic_currency=CurrencyTypes.GBP, foreign_currency=CurrencyTypes.USD)


def test_basis_curve_matches_reference_nodes(market):
    g, _, _, _, xc = market
    assert np.max(np.abs(xc._times - np.array(g["xccy_times"]))) == 0.0
    assert np.max(np.abs(xc._dfs - np.array(g["xccy_dfs"]))) < 1e-14


def test_fix_float_and_fix_fix_values_match_reference(market, capsys):
    g, vd, gbp, usd, xc = market
    for r in g["fix_float"]:
        sw = XccyFixFloat(effective_dt=Date(*r["effective"]), term_dt_or_tenor=r["tenor"], domestic_leg_type=SwapTypes[r["side"]],
                          domestic_coupon=r["coupon"], foreign_spread=r["foreign"], domestic_freq_type=FrequencyTypes[r["dom_freq"]],
                          foreign_freq_type=FrequencyTypes[r["for_freq"]], **COMMON)
        assert abs(sw.value(vd, gbp, usd, xc, g["spot"]) - r["value"]) <= 1e-12 * 1_000_000, r["id"]
        if "value_fixing" in r:
            assert abs(sw.value(vd, gbp, usd, xc, g["spot"], 0.0525) - r["value_fixing"]) <= 1e-12 * 1_000_000, r["id"]
    for r in g["fix_fix"]:
        sw = XccyFixFix(effective_dt=Date(*r["effective"]), term_dt_or_tenor=r["tenor"], domestic_leg_type=SwapTypes[r["side"]],
                        domestic_coupon=r["coupon"], foreign_coupon=r["foreign"], domestic_freq_type=FrequencyTypes[r["dom_freq"]],
                        foreign_freq_type=FrequencyTypes[r["for_freq"]], **COMMON)
        assert abs(sw.value(vd, gbp, usd, xc, g["spot"]) - r["value"]) <= 1e-12 * 1_000_000, r["id"]
    assert len(g["fix_float"]) == 4 and len(g["fix_fix"]) == 4
    sw.print_valuation()                                  # the report of the last valuation: both legs, one row per payment
    shown = capsys.readouterr().out
    assert "DOMESTIC FIXED LEG VALUATION:" in shown and "FOREIGN FIXED LEG VALUATION:" in shown and shown.count("PAYMENTS VALUATION:") == 2
    with pytest.raises(LibError, match="Start date after maturity date"):
        XccyFixFix(effective_dt=vd, term_dt_or_tenor=vd.add_days(-30), domestic_leg_type=SwapTypes.PAY, domestic_coupon=0.04,
                   foreign_coupon=0.05, domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.ANNUAL, **COMMON)
