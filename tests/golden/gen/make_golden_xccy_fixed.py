#!/usr/bin/env python
"""Known answers of XccyFixFloat.value / XccyFixFix.value (cavour/trades/rates/xccy_fix_float_swap.py:196-245,
xccy_fix_fix_swap.py:210-280) from the UNMODIFIED reference on the market of its own tests (tests/test_xccy_fix_float.py:296-377:
GBP / USD OIS curves and a seven-pillar basis curve, all flat-forward).  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_xccy_fixed.py

Writes tests/golden/ref_xccy_fixed.json.
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.global_types import SwapTypes, CurveTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.models.models import Model
from cavour.trades.rates.xccy_basis_swap import XccyBasisSwap
from cavour.trades.rates.xccy_curve import XccyCurve
from cavour.trades.rates.xccy_fix_float_swap import XccyFixFloat
from cavour.trades.rates.xccy_fix_fix_swap import XccyFixFix

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
VD = (15, 6, 2023)
TENORS = ["1Y", "2Y", "3Y", "4Y", "5Y", "7Y", "10Y"]
GBP = [4.50, 4.55, 4.60, 4.65, 4.70, 4.74, 4.80]
USD = [5.20, 5.25, 5.30, 5.35, 5.40, 5.44, 5.50]
BASIS = [0.0025, 0.0028, 0.0030, 0.0032, 0.0034, 0.0036, 0.0039]
SPOT = 0.79
# id, effective, tenor, domestic side, domestic coupon, foreign spread / coupon, domestic freq, foreign freq
FIX_FLOAT = [("ff_5y_pay", VD, "5Y", "PAY", 0.047, 0.0034, "ANNUAL", "QUARTERLY"), ("ff_3y_rec", VD, "3Y", "RECEIVE", 0.045, 0.0, "SEMI_ANNUAL", "SEMI_ANNUAL"),
             ("ff_fwd_7y_pay", (15, 9, 2023), "7Y", "PAY", 0.05, 0.001, "ANNUAL", "ANNUAL"), ("ff_seasoned_4y", (15, 12, 2022), "4Y", "RECEIVE", 0.044, 0.002, "ANNUAL", "QUARTERLY")]
FIX_FIX = [("xx_5y_pay", VD, "5Y", "PAY", 0.047, 0.054, "ANNUAL", "QUARTERLY"), ("xx_10y_rec", VD, "10Y", "RECEIVE", 0.048, 0.055, "SEMI_ANNUAL", "ANNUAL"),
           ("xx_fwd_2y", (17, 7, 2023), "2Y", "PAY", 0.04, 0.05, "QUARTERLY", "QUARTERLY"), ("xx_seasoned_3y", (15, 3, 2023), "3Y", "RECEIVE", 0.046, 0.052, "ANNUAL", "ANNUAL")]


def nodes_to_numpy(c):
    c._times, c._dfs = np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)
    return c


def main():
    vd = Date(*VD)
    curves = {}
    for name, px, dc in (("GBP_OIS_SONIA", GBP, DayCountTypes.ACT_365F), ("USD_OIS_SOFR", USD, DayCountTypes.ACT_360)):
        m = Model(vd)
        m.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0, swap_type=SwapTypes.PAY, fixed_dcc_type=dc,
                      fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                      bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.FLAT_FWD_RATES)
        curves[name] = nodes_to_numpy(getattr(m.curves, name))
    gbp, usd = curves["GBP_OIS_SONIA"], curves["USD_OIS_SOFR"]
    basis = [XccyBasisSwap(effective_dt=vd, term_dt_or_tenor=t, domestic_notional=SPOT * 1_000_000, foreign_notional=1_000_000,
                           domestic_spread=0.0, foreign_spread=s, domestic_freq_type=FrequencyTypes.ANNUAL,
                           foreign_freq_type=FrequencyTypes.ANNUAL, domestic_dc_type=DayCountTypes.ACT_365F,
                           foreign_dc_type=DayCountTypes.ACT_360, domestic_floating_index=CurveTypes.GBP_OIS_SONIA,
                           foreign_floating_index=CurveTypes.USD_OIS_SOFR, domestic_currency=CurrencyTypes.GBP,
                           foreign_currency=CurrencyTypes.USD) for t, s in zip(TENORS, BASIS)]
    xc = nodes_to_numpy(XccyCurve(value_dt=vd, basis_swaps=basis, domestic_curve=gbp, foreign_curve=usd, spot_fx=SPOT,
                                  interp_type=InterpTypes.FLAT_FWD_RATES, check_refit=False))
    common = dict(domestic_notional=790_000, foreign_notional=1_000_000, domestic_dc_type=DayCountTypes.ACT_365F,
                  foreign_dc_type=DayCountTypes.ACT_360, domestic_floating_index=CurveTypes.GBP_OIS_SONIA,
                  foreign_floating_index=CurveTypes.USD_OIS_SOFR, domestic_currency=CurrencyTypes.GBP, foreign_currency=CurrencyTypes.USD)
    out = {"value_dt": VD, "tenors": TENORS, "gbp": GBP, "usd": USD, "basis": BASIS, "spot": SPOT,
           "xccy_times": [float(x) for x in xc._times], "xccy_dfs": [float(x) for x in xc._dfs], "fix_float": [], "fix_fix": []}
    for sid, eff, ten, side, cpn, spr, dfq, ffq in FIX_FLOAT:
        sw = XccyFixFloat(effective_dt=Date(*eff), term_dt_or_tenor=ten, domestic_leg_type=SwapTypes[side], domestic_coupon=cpn,
                          foreign_spread=spr, domestic_freq_type=FrequencyTypes[dfq], foreign_freq_type=FrequencyTypes[ffq], **common)
        rec = {"id": sid, "effective": eff, "tenor": ten, "side": side, "coupon": cpn, "foreign": spr, "dom_freq": dfq, "for_freq": ffq}
        try:
            rec["value"] = float(sw.value(vd, gbp, usd, xc, SPOT))
            if eff == VD:
                rec["value_fixing"] = float(sw.value(vd, gbp, usd, xc, SPOT, 0.0525))
        except Exception as ex:  # noqa: BLE001
            rec["error"] = type(ex).__name__ + ": " + str(ex)
        out["fix_float"].append(rec)
        print(rec, flush=True)
    for sid, eff, ten, side, cpn, fcpn, dfq, ffq in FIX_FIX:
        sw = XccyFixFix(effective_dt=Date(*eff), term_dt_or_tenor=ten, domestic_leg_type=SwapTypes[side], domestic_coupon=cpn,
                        foreign_coupon=fcpn, domestic_freq_type=FrequencyTypes[dfq], foreign_freq_type=FrequencyTypes[ffq], **common)
        rec = {"id": sid, "effective": eff, "tenor": ten, "side": side, "coupon": cpn, "foreign": fcpn, "dom_freq": dfq, "for_freq": ffq}
        try:
            rec["value"] = float(sw.value(vd, gbp, usd, xc, SPOT))
        except Exception as ex:  # noqa: BLE001
            rec["error"] = type(ex).__name__ + ": " + str(ex)
        out["fix_fix"].append(rec)
        print(rec, flush=True)
    with open(os.path.join(OUT, "ref_xccy_fixed.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
