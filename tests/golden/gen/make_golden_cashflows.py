#!/usr/bin/env python
"""CASHFLOWS-request and holiday-calendar trade goldens from the UNMODIFIED reference.  TEST INFRASTRUCTURE, build
container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_cashflows.py

Writes tests/golden/ref_cashflows.json: per trade the rows of `Position.compute([VALUE, DELTA, GAMMA, CASHFLOWS]).cashflows`
(engine.py:190-213) next to VALUE / the delta ladder / gamma, for OIS on the WEEKEND calendar and on holiday calendars
(cal_type = UNITED_KINGDOM / TARGET / UNITED_STATES: pins schedules rolled on holiday calendars through the whole engine).
"""
import json
import os
import time

import numpy as np

import make_golden as mg
from cavour.utils.date import Date
from cavour.utils.calendar import BusDayAdjustTypes, CalendarTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.global_types import SwapTypes, CurveTypes, RequestTypes
from cavour.utils.currency import CurrencyTypes
from cavour.trades.rates.ois import OIS
from cavour.market.position.engine import Engine

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

# id, curve, effective, tenor, side, coupon, notional, fixed freq, float freq, spread, calendar
TRADES = [
    ("cf_readme_10y", "gbp_readme_lzr", None, "10Y", "PAY", 0.045, 1e7, "ANNUAL", "ANNUAL", 0.0, "WEEKEND"),
    ("cf_off_7y_rec", "gbp_readme_lzr", (17, 6, 2024), "7Y", "RECEIVE", 0.041, 1e6, "ANNUAL", "ANNUAL", 0.0, "WEEKEND"),
    ("cf_semi_quart_spread", "gbp_readme_lzr", None, "5Y", "PAY", 0.0425, 2e6, "SEMI_ANNUAL", "QUARTERLY", 0.0015, "WEEKEND"),
    ("cf_seasoned_3y", "gbp_readme_lzr", (15, 11, 2023), "3Y", "RECEIVE", 0.044, 5e6, "SEMI_ANNUAL", "SEMI_ANNUAL", 0.0, "WEEKEND"),
    ("cf_ff_off_4y", "gbp_readme_ff", (17, 6, 2024), "4Y", "PAY", 0.043, 1e6, "ANNUAL", "ANNUAL", 0.0, "WEEKEND"),
    ("cal_uk_xmas_6y", "gbp_readme_lzr", (24, 12, 2024), "6Y", "PAY", 0.0415, 3e6, "SEMI_ANNUAL", "SEMI_ANNUAL", 0.0, "UNITED_KINGDOM"),
    ("cal_uk_easter_12y", "gbp_readme_lzr", (17, 4, 2025), "12Y", "RECEIVE", 0.0412, 1e6, "ANNUAL", "QUARTERLY", 0.001, "UNITED_KINGDOM"),
    ("cal_target_may_9y", "gbp_readme_lzr", (30, 4, 2024), "9Y", "PAY", 0.041, 1e6, "ANNUAL", "ANNUAL", 0.0, "TARGET"),
    ("cal_us_july_5y", "gbp_readme_lzr", (3, 7, 2024), "5Y", "RECEIVE", 0.043, 4e6, "QUARTERLY", "QUARTERLY", 0.0, "UNITED_STATES"),
]


def main():
    t0 = time.time()
    models, caches = {}, {}
    out = []
    for (tid, ckey, eff, tenor, side, cpn, notl, ffreq, lfreq, spread, cal) in TRADES:
        name, vd, px, freq, dc, interp = mg.CURVES[ckey]
        if ckey not in models:
            models[ckey] = mg.build_model(ckey)
            curve = getattr(models[ckey].curves, name)
            ck = tuple(curve.swap_times)
            caches[ckey] = {ck: Engine(models[ckey])._cached_curve(ck, curve.swap_rates, curve.swap_times, curve.year_fracs,
                                                                    curve._interp_type)}
            # the non-AD legs query curve.df(): numpy node arrays (torch tensors have no `.size`)
            curve._times, curve._dfs = np.asarray(curve._times, dtype=np.float64), np.asarray(curve._dfs, dtype=np.float64)
            print("curve", ckey, time.time() - t0, flush=True)
        model = models[ckey]
        eff_dt = Date(*vd) if eff is None else Date(*eff)
        swap = OIS(effective_dt=eff_dt, term_dt_or_tenor=tenor, fixed_leg_type=SwapTypes[side], fixed_coupon=cpn,
                   fixed_freq_type=FrequencyTypes[ffreq], fixed_dc_type=DayCountTypes[dc], floating_index=CurveTypes[name],
                   currency=CurrencyTypes[name[:3]], notional=notl, float_spread=spread, float_freq_type=FrequencyTypes[lfreq],
                   float_dc_type=DayCountTypes[dc], cal_type=CalendarTypes[cal], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        pos = swap.position(model)
        pos._engine._curve_cache = caches[ckey]
        head = {"id": tid, "curve": ckey, "effective": mg.dmy(eff_dt), "tenor": tenor, "side": side, "coupon": cpn, "notional": notl,
                "fixed_freq": ffreq, "float_freq": lfreq, "spread": spread, "cal": cal}
        try:
            res = pos.compute([RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA, RequestTypes.CASHFLOWS])
        except Exception as ex:  # noqa: BLE001  (a seasoned swap: the non-AD float leg looks up a DF before the value date)
            out.append(dict(head, error=type(ex).__name__ + ": " + str(ex)))
            print(tid, out[-1]["error"], flush=True)
            continue
        cf = res.cashflows
        out.append({
            **head,
            "value": float(res.value.amount), "delta": [float(x) for x in np.asarray(res.risk.risk_ladder)],
            "gamma": np.asarray(res.gamma.risk_ladder, dtype=np.float64).tolist(),
            "fixed_payment_dts": [mg.dmy(d) for d in swap._fixed_leg._payment_dts],
            "float_payment_dts": [mg.dmy(d) for d in swap._float_leg._payment_dts],
            "rows": [{"payment_date": mg.dmy(c.payment_date), "notional": c.notional, "payment_fraction": c.payment_fraction,
                      "accrual_period": c.accrual_period, "amount": c.amount, "discount_factor": c.discount_factor,
                      "discounted_amount": c.discounted_amount, "leg_type": c.leg_type} for c in cf.cashflows],
            "total_amount": float(cf.total_amount), "total_pv": float(cf.total_pv),
            "repr": repr(cf), "first_row_dict": cf.cashflows[0].to_dict(),
        })
        print(tid, out[-1]["value"], len(cf), out[-1]["total_pv"], time.time() - t0, flush=True)
    with open(os.path.join(OUT, "ref_cashflows.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
