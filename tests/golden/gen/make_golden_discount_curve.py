#!/usr/bin/env python
"""Known answers of the reference's DiscountCurve rate views (cavour/market/curves/discount_curve.py: zero_rate, cc_rate,
swap_rate, fwd, _fwd, fwd_rate, bump, _zero_to_df, _df_to_zero, survival_prob) on a plain DiscountCurve and on a bootstrapped
OIS curve.  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_discount_curve.py

Writes tests/golden/ref_discount_curve.json.
"""
import json
import os

import numpy as np

import make_golden as mg
from cavour.utils.date import Date
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.market.curves.discount_curve import DiscountCurve

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OFFSETS, VALUES = [0.5, 1.0, 2.0, 5.0, 10.0], [0.975, 0.95, 0.90, 0.78, 0.60]
TENORS = ["1M", "9M", "18M", "4Y", "10Y", "12Y"]


def flt(v):
    return [float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)]


def views(c, vd):
    dts = [vd.add_tenor(t) for t in TENORS]
    out = {"zero_cont_act360": flt(c.zero_rate(dts)), "zero_single": float(c.zero_rate(dts[3])),
           "cc_rate": flt(c.cc_rate(dts)), "survival": float(np.asarray(c.survival_prob(dts[2])).reshape(-1)[0]),
           "fwd": flt(c.fwd(dts)), "fwd_single": float(c.fwd(dts[1])), "_fwd": flt(c._fwd(np.array([0.0, 0.3, 1.0, 4.2, 11.0]))),
           "fwd_rate_3m": flt(c.fwd_rate(dts, "3M")), "fwd_rate_single": float(c.fwd_rate(dts[0], dts[4], DayCountTypes.ACT_365F)),
           "fwd_rate_lists": flt(c.fwd_rate(dts[:3], dts[3:])),
           "swap_rate": flt(c.swap_rate(vd, [dts[2], dts[3], dts[4]])), "swap_rate_single": flt(c.swap_rate(vd.add_tenor("6M"), dts[3], FrequencyTypes.SEMI_ANNUAL, DayCountTypes.ACT_360)),
           "zero": {}}
    for fq in ("CONTINUOUS", "SIMPLE", "ANNUAL", "SEMI_ANNUAL", "QUARTERLY", "MONTHLY"):
        for dc in ("ACT_360", "ACT_365F", "THIRTY_E_360"):
            out["zero"][f"{fq}/{dc}"] = flt(c.zero_rate(dts, FrequencyTypes[fq], DayCountTypes[dc]))
    b = c.bump(0.0025)
    out["bump"] = {"times": flt(b._times), "dfs": flt(b._dfs), "df": flt(b.df(dts))}
    out["zero_to_df"] = {fq: flt(c._zero_to_df(vd, np.array([0.01, 0.03, 0.05]), np.array([0.0, 1.5, 7.0]), FrequencyTypes[fq], DayCountTypes.ACT_360))
                         for fq in ("CONTINUOUS", "SIMPLE", "ANNUAL", "SEMI_ANNUAL", "QUARTERLY", "MONTHLY")}
    out["zero_to_df_scalar"] = flt(c._zero_to_df(vd, 0.04, 2.5, FrequencyTypes.ANNUAL, DayCountTypes.ACT_360))
    return out


def main():
    vd = Date(30, 4, 2024)
    out = {"value_dt": [30, 4, 2024], "offsets": OFFSETS, "values": VALUES, "tenors": TENORS, "plain": {}}
    for it in ("FLAT_FWD_RATES", "LINEAR_ZERO_RATES", "LINEAR_FWD_RATES", "PCHIP_ZERO_RATES"):
        out["plain"][it] = views(DiscountCurve(vd, OFFSETS, np.array(VALUES), InterpTypes[it]), vd)
    model = mg.build_model("gbp_readme_lzr")
    c = model.curves.GBP_OIS_SONIA
    c._times, c._dfs = np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)
    out["ois_gbp_readme_lzr"] = views(c, vd)
    with open(os.path.join(OUT, "ref_discount_curve.json"), "w") as f:
        json.dump(out, f)
    print("done")


if __name__ == "__main__":
    main()
