#!/usr/bin/env python
"""CASHFLOWS-request goldens of year-on-year inflation swaps (Engine._compute_yoy_iis, engine.py:1355-1406) and of OIS with
cross-currency collateral (engine.py:497-501) from the UNMODIFIED reference.  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_cashflows_yoy.py

Writes tests/golden/ref_cashflows_yoy.json: per swap of make_golden_yoy.SWAPS (both index conventions) the rows of
`Position.compute([VALUE, CASHFLOWS]).cashflows`, or the error the reference raises.
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.global_types import SwapTypes, InflationIndexTypes, InflationInterpTypes, RequestTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.market.curves.inflation_curve import InflationCurve
from cavour.trades.rates.zcis import ZeroCouponInflationSwap
from cavour.trades.rates.yoy_inflation_swap import YoYInflationSwap
from cavour.models.models import Model
from cavour.market.position.position import Position

from make_golden import GBP_PX, TENORS, dmy
from make_golden_yoy import SWAPS
from make_golden_zcis import CALIB, INDEX_SPECS, VALUE_DT, make_index

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def rows_of(cf):
    return [{"payment_date": dmy(c.payment_date), "notional": c.notional, "payment_fraction": c.payment_fraction,
             "accrual_period": c.accrual_period, "amount": c.amount, "discount_factor": c.discount_factor,
             "discounted_amount": c.discounted_amount, "leg_type": c.leg_type} for c in cf.cashflows]


def main():
    vd = Date(*VALUE_DT)
    out = {"cases": []}
    for iname in ("rpi_linear", "rpi_flat_lag2"):
        spec = INDEX_SPECS[iname]
        model = Model(vd)
        model.build_curve(name="GBP_OIS_SONIA", px_list=GBP_PX, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
        c = model.curves.GBP_OIS_SONIA           # the non-AD legs query curve.df(): numpy node arrays
        c._times, c._dfs = np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)
        idx = make_index(spec)
        calib = [ZeroCouponInflationSwap(vd, ten, SwapTypes.PAY, r, idx, 1_000_000) for ten, r in CALIB]
        ic = InflationCurve(vd, calib, 293.8, CurrencyTypes.GBP, InflationIndexTypes.UK_RPI,
                            discount_curve=model.curves.GBP_OIS_SONIA, interp_type=InflationInterpTypes[spec["interp"]])
        model._curves_dict["GBP_RPI_INFLATION"] = ic
        for sid, eff, ten, side, rate, freq, notional, spread, dc, lag, bd in SWAPS:
            sw = YoYInflationSwap(Date(*eff), ten if isinstance(ten, str) else Date(*ten), SwapTypes[side], rate, idx,
                                  FrequencyTypes[freq], notional, spread, DayCountTypes[dc], lag,
                                  bd_type=BusDayAdjustTypes[bd])
            head = {"id": f"{iname}_{sid}", "index": iname}
            for reqs, tag in (([RequestTypes.CASHFLOWS], "cf_only"), ([RequestTypes.VALUE, RequestTypes.CASHFLOWS], "value_cf")):
                try:
                    res = Position(sw, model).compute(reqs)
                except Exception as ex:  # noqa: BLE001
                    head[tag] = {"error": type(ex).__name__ + ": " + str(ex)}
                    print(head["id"], tag, head[tag]["error"], flush=True)
                    continue
                cf = res.cashflows
                head[tag] = {"value": None if res.value is None else float(res.value.amount), "rows": rows_of(cf),
                             "total_amount": float(cf.total_amount), "total_pv": float(cf.total_pv), "repr": repr(cf),
                             "has_risk": res.risk is not None, "has_gamma": res.gamma is not None}
                print(head["id"], tag, head[tag]["value"], len(cf), sorted({r["leg_type"] for r in head[tag]["rows"]}), flush=True)
            out["cases"].append(head)
    with open(os.path.join(OUT, "ref_cashflows_yoy.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
