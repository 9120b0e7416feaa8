#!/usr/bin/env python
"""Known answers of the reference's non-AD host methods next to the path: YoYInflationSwap.value / breakeven_rate / pv01 and the
lists SwapYoYInflationLeg.value keeps (yoy_inflation_swap.py:224-371, swap_yoy_inflation_leg.py:267-366), ZeroCouponInflationSwap
.pv01 (zcis.py:284-317), OIS.ir01 / pv01 / swap_rate (ois.py:250-330).  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_host_values.py

Writes tests/golden/ref_host_values.json.  Swaps / curves are those of make_golden_yoy.py and make_golden.py.
"""
import json
import os

import numpy as np

import make_golden as mg
from cavour.utils.date import Date
from cavour.utils.global_types import SwapTypes, InflationIndexTypes, InflationInterpTypes, CurveTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.curves.inflation_curve import InflationCurve
from cavour.trades.rates.zcis import ZeroCouponInflationSwap
from cavour.trades.rates.yoy_inflation_swap import YoYInflationSwap
from cavour.trades.rates.ois import OIS

from make_golden_yoy import SWAPS
from make_golden_zcis import CALIB, INDEX_SPECS, VALUE_DT, make_index

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    vd = Date(*VALUE_DT)
    model = mg.build_model("gbp_readme_lzr")
    c = model.curves.GBP_OIS_SONIA
    c._times, c._dfs = np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)
    out = {"yoy": [], "zcis_pv01": [], "ois": []}
    for iname in ("rpi_linear", "rpi_flat_lag2"):
        spec = INDEX_SPECS[iname]
        idx = make_index(spec)
        calib = [ZeroCouponInflationSwap(vd, ten, SwapTypes.PAY, r, idx, 1_000_000) for ten, r in CALIB]
        ic = InflationCurve(vd, calib, 293.8, CurrencyTypes.GBP, InflationIndexTypes.UK_RPI, discount_curve=c,
                            interp_type=InflationInterpTypes[spec["interp"]])
        for sid, eff, ten, side, rate, freq, notional, spread, dc, lag, bd in SWAPS:
            sw = YoYInflationSwap(Date(*eff), ten if isinstance(ten, str) else Date(*ten), SwapTypes[side], rate, idx,
                                  FrequencyTypes[freq], notional, spread, DayCountTypes[dc], lag, bd_type=BusDayAdjustTypes[bd])
            rec = {"id": f"{iname}_{sid}", "pv01": float(sw.pv01(vd, c))}
            try:
                rec["value"] = float(sw.value(vd, c, ic))
                leg = sw._inflation_leg
                rec.update(fixed_pv=float(sw._fixed_pv), inflation_pv=float(sw._inflation_pv),
                           start_cpis=[float(x) for x in leg._start_cpis], end_cpis=[float(x) for x in leg._end_cpis],
                           yoy_rates=[float(x) for x in leg._yoy_rates], payments=[float(x) for x in leg._payments],
                           dfs=[float(x) for x in leg._dfs], pvs=[float(x) for x in leg._pvs],
                           breakeven=float(sw.breakeven_rate(vd, c, ic)))
            except Exception as ex:  # noqa: BLE001
                rec["error"] = type(ex).__name__ + ": " + str(ex)
            out["yoy"].append(rec)
            print(rec["id"], rec.get("value"), rec.get("error"), flush=True)
        if iname == "rpi_linear":
            for ten, r in CALIB:
                z = ZeroCouponInflationSwap(vd, ten, SwapTypes.PAY, r, idx, 2_500_000)
                out["zcis_pv01"].append({"tenor": ten, "rate": r, "pv01": float(z.pv01(vd, c))})
            seasoned = ZeroCouponInflationSwap(Date(30, 4, 2019), "5Y", SwapTypes.RECEIVE, 0.03, idx, 1_000_000)
            out["zcis_pv01"].append({"tenor": "matured", "rate": 0.03, "pv01": float(seasoned.pv01(vd, c))})
    for tenor, side, cpn, ffreq, lfreq in (("2Y", "PAY", 0.045, "ANNUAL", "ANNUAL"), ("10Y", "RECEIVE", 0.04, "SEMI_ANNUAL", "QUARTERLY"),
                                           ("30Y", "PAY", 0.0375, "ANNUAL", "ANNUAL")):
        sw = OIS(effective_dt=vd, term_dt_or_tenor=tenor, fixed_leg_type=SwapTypes[side], fixed_coupon=cpn,
                 fixed_freq_type=FrequencyTypes[ffreq], fixed_dc_type=DayCountTypes.ACT_365F, floating_index=CurveTypes.GBP_OIS_SONIA,
                 currency=CurrencyTypes.GBP, notional=5e6, float_freq_type=FrequencyTypes[lfreq], float_dc_type=DayCountTypes.ACT_365F,
                 bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        out["ois"].append({"tenor": tenor, "side": side, "coupon": cpn, "fixed_freq": ffreq, "float_freq": lfreq,
                           "value": float(sw.value(vd, c)), "pv01": float(sw.pv01(vd, c)), "swap_rate": float(sw.swap_rate(vd, c)),
                           "ir01": float(sw.ir01(vd, c))})
        print(tenor, out["ois"][-1], flush=True)
    with open(os.path.join(OUT, "ref_host_values.json"), "w") as f:
        json.dump(out, f)
    print("done")


if __name__ == "__main__":
    main()
