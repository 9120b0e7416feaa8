#!/usr/bin/env python
"""Golden vectors for year-on-year inflation swaps from the UNMODIFIED reference (/root/reference):
Position(yoy, model).compute([VALUE, DELTA, GAMMA]) = Engine._compute_yoy_iis (engine.py:986-1408).

TEST INFRASTRUCTURE; build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_yoy.py

The reference has no Model method that registers an inflation curve (and no test of this route); the engine looks
it up as model.curves.GBP_RPI_INFLATION, so the curve is put into the model's curve dict directly.
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

from make_golden import GBP_PX, TENORS
from make_golden_zcis import CALIB, INDEX_SPECS, VALUE_DT, make_index

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

# id, effective, tenor, fixed side, fixed rate, freq, notional, spread, day count, payment lag, bd
SWAPS = [
    ("spot_5y_annual", (30, 4, 2024), "5Y", "PAY", 0.034, "ANNUAL", 10_000_000.0, 0.0, "ACT_365F", 0, "FOLLOWING"),
    ("spot_10y_semi_rec", (30, 4, 2024), "10Y", "RECEIVE", 0.031, "SEMI_ANNUAL", 2_500_000.0, 0.001, "ACT_365F", 0, "MODIFIED_FOLLOWING"),
    ("fwd_7y_quarterly", (17, 6, 2024), "7Y", "PAY", 0.0335, "QUARTERLY", 1_000_000.0, -0.0005, "ACT_365F", 0, "MODIFIED_FOLLOWING"),
    ("fwd_30y_annual_lag2", (2, 9, 2024), "30Y", "RECEIVE", 0.03, "ANNUAL", 50_000_000.0, 0.0, "ACT_365F", 2, "FOLLOWING"),
    ("seasoned_4y_annual", (15, 11, 2022), (15, 11, 2026), "PAY", 0.04, "ANNUAL", 3_000_000.0, 0.002, "ACT_365F", 0, "FOLLOWING"),
    ("spot_2y_act360", (30, 4, 2024), "2Y", "PAY", 0.0325, "SEMI_ANNUAL", 750_000.0, 0.0, "ACT_360", 0, "FOLLOWING"),
    ("long_40y_annual", (30, 4, 2024), "40Y", "PAY", 0.032, "ANNUAL", 5_000_000.0, 0.0, "ACT_365F", 0, "FOLLOWING"),
]


def main():
    vd = Date(*VALUE_DT)
    reqs = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
    out = {"value_dt": VALUE_DT, "gbp_px": GBP_PX, "tenors": TENORS, "calibration": CALIB, "index_specs": INDEX_SPECS,
           "base_cpi": 293.8, "cases": []}
    for iname in ("rpi_linear", "rpi_flat_lag2"):
        spec = INDEX_SPECS[iname]
        model = Model(vd)
        model.build_curve(name="GBP_OIS_SONIA", px_list=GBP_PX, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
        idx = make_index(spec)
        calib = [ZeroCouponInflationSwap(vd, ten, SwapTypes.PAY, r, idx, 1_000_000) for ten, r in CALIB]
        ic = InflationCurve(vd, calib, 293.8, CurrencyTypes.GBP, InflationIndexTypes.UK_RPI,
                            discount_curve=model.curves.GBP_OIS_SONIA, interp_type=InflationInterpTypes[spec["interp"]])
        model._curves_dict["GBP_RPI_INFLATION"] = ic
        # recorded before any valuation: Engine._compute_yoy_iis re-runs _build_curve_ad under jacrev, which leaves
        # traced arrays in ic._times / ic._dfs (the engine itself guards against that, engine.py:1062-1075)
        out.setdefault("inflation_curves", {})[iname] = {
            "times": [float(x) for x in np.asarray(ic._times, dtype=np.float64)],
            "dfs": [float(x) for x in np.asarray(ic._dfs, dtype=np.float64)], "interp": ic._interp_type.name,
            "swap_times": [float(x) for x in ic.swap_times]}
        for sid, eff, ten, side, rate, freq, notional, spread, dc, lag, bd in SWAPS:
            sw = YoYInflationSwap(Date(*eff), ten if isinstance(ten, str) else Date(*ten), SwapTypes[side], rate, idx,
                                  FrequencyTypes[freq], notional, spread, DayCountTypes[dc], lag,
                                  bd_type=BusDayAdjustTypes[bd])
            res = Position(sw, model).compute(reqs)      # YoYInflationSwap has no .position() helper
            leg = sw._inflation_leg
            disc_d, infl_d = res.risk.GBP_OIS_SONIA, res.risk.GBP_RPI_INFLATION
            disc_g, infl_g = res.gamma.GBP_OIS_SONIA, res.gamma.GBP_RPI_INFLATION
            out["cases"].append({
                "id": f"{iname}_{sid}", "index": iname, "effective": eff, "tenor": ten, "fixed_leg": side,
                "fixed_rate": rate, "freq": freq, "notional": notional, "spread": spread, "dc": dc, "payment_lag": lag,
                "bd": bd,
                "payment_dts": [[d._d, d._m, d._y] for d in leg._payment_dts],
                "yoy_start_dts": [[d._d, d._m, d._y] for d in leg._yoy_start_dts],
                "year_fracs": [float(x) for x in leg._year_fracs],
                "value": float(res.value.amount),
                "disc_delta": [float(x) for x in np.asarray(disc_d.risk_ladder)],
                "infl_delta": [float(x) for x in np.asarray(infl_d.risk_ladder)],
                "infl_tenors": list(infl_d.tenors),
                "disc_gamma": np.asarray(disc_g.risk_ladder, dtype=np.float64).tolist(),
                "infl_gamma": np.asarray(infl_g.risk_ladder, dtype=np.float64).tolist(),
            })
    with open(os.path.join(OUT, "ref_yoy.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "YoY cases")


if __name__ == "__main__":
    main()
