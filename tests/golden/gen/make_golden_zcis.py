#!/usr/bin/env python
"""Golden vectors for zero-coupon inflation swaps from the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE; build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_zcis.py

Runs the reference's InflationIndex / InflationCurve / ZeroCouponInflationSwap.value on two discount curves
(the hand-built FLAT_FWD curve of the reference's tests/test_zcis.py fixtures and the README SONIA OIS curve,
LINEAR_ZERO_RATES, path A) and writes tests/golden/ref_zcis.json.
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.global_types import SwapTypes, InflationIndexTypes, InflationInterpTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.indices.inflation_index import InflationIndex
from cavour.market.curves.discount_curve import DiscountCurve
from cavour.market.curves.inflation_curve import InflationCurve
from cavour.market.curves.interpolator import InterpTypes
from cavour.trades.rates.zcis import ZeroCouponInflationSwap
from cavour.models.models import Model

from make_golden import GBP_PX, TENORS

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
VALUE_DT = (30, 4, 2024)
FIXINGS = [((1, 11, 2023), 290.1), ((1, 12, 2023), 290.9), ((1, 1, 2024), 291.6), ((1, 2, 2024), 292.4),
           ((1, 3, 2024), 293.0), ((1, 4, 2024), 293.5), ((1, 5, 2024), 294.0)]
SEASONALITY = {1: 0.994, 2: 1.001, 3: 1.002, 4: 1.006, 5: 1.003, 6: 1.001, 7: 0.996, 8: 1.002, 9: 1.000, 10: 0.999,
               11: 0.997, 12: 0.999}
CALIB = [("1Y", 0.031), ("2Y", 0.0325), ("3Y", 0.0335), ("5Y", 0.034), ("7Y", 0.0345), ("10Y", 0.035),
         ("15Y", 0.0342), ("20Y", 0.0335), ("30Y", 0.032)]
FLAT_CURVE = ([0.25, 0.5, 1.0, 2.0, 5.0, 10.0, 30.0], [0.9875, 0.975, 0.95, 0.90, 0.78, 0.61, 0.22])

INDEX_SPECS = {
    "rpi_linear": dict(interp="LINEAR", lag=3, seasonality=None),
    "rpi_flat_lag2": dict(interp="FLAT", lag=2, seasonality=None),
    "rpi_compound_seasonal": dict(interp="COMPOUND", lag=3, seasonality=SEASONALITY),
}


def make_index(spec):
    idx = InflationIndex(InflationIndexTypes.UK_RPI, Date(*FIXINGS[0][0]), FIXINGS[0][1], CurrencyTypes.GBP,
                         lag_months=spec["lag"], interp_type=InflationInterpTypes[spec["interp"]],
                         seasonality_factors=spec["seasonality"])
    for d, v in FIXINGS:
        idx.add_fixing(Date(*d), v)
    return idx


def main():
    vd = Date(*VALUE_DT)
    model = Model(vd)
    model.build_curve(name="GBP_OIS_SONIA", px_list=GBP_PX, tenor_list=TENORS, spot_days=0,
                      interp_type=InterpTypes.LINEAR_ZERO_RATES)
    ois = model.curves.GBP_OIS_SONIA
    # shim artefact: under refshim the node arrays are torch tensors (`.size` is a method there, an int on the
    # jax / numpy arrays the reference expects); same values, plain float64 numpy arrays
    ois._times = np.asarray(ois._times, dtype=np.float64)
    ois._dfs = np.asarray(ois._dfs, dtype=np.float64)
    flat = DiscountCurve(vd, FLAT_CURVE[0], np.array(FLAT_CURVE[1]), InterpTypes.FLAT_FWD_RATES)
    dcurves = {"gbp_ois_lzr": ois, "flat_ff": flat}
    rng = np.random.Generator(np.random.PCG64(11))
    trades = []
    out = {"value_dt": VALUE_DT, "fixings": FIXINGS, "seasonality": SEASONALITY, "calibration": CALIB,
           "flat_curve": FLAT_CURVE, "index_specs": INDEX_SPECS, "base_cpi": 293.8,
           "discount_curves": {k: {"times": [float(x) for x in c._times], "dfs": [float(x) for x in c._dfs],
                                   "interp": c._interp_type.name} for k, c in dcurves.items()}}
    infl_curves = {}
    for iname, spec in INDEX_SPECS.items():
        idx = make_index(spec)
        calib = [ZeroCouponInflationSwap(vd, ten, SwapTypes.PAY, r, idx, 1_000_000) for ten, r in CALIB]
        ic = InflationCurve(vd, calib, 293.8, CurrencyTypes.GBP, InflationIndexTypes.UK_RPI, discount_curve=flat,
                            interp_type=InflationInterpTypes[spec["interp"]])
        infl_curves[iname] = {"times": [float(x) for x in ic._times], "dfs": [float(x) for x in ic._dfs],
                              "interp": ic._interp_type.name,
                              "forward_index": {f"{y}": float(ic.forward_index(vd.add_years(y))) for y in (0.5, 1, 2.5, 7, 12, 30, 35)}}
        n = 0
        for dname, dc in dcurves.items():
            for k in range(8):
                tenor = int(rng.integers(1, 31))
                start_off = int(rng.integers(0, 120)) if k % 2 else 0        # forward-starting half of the time
                eff = vd.add_days(start_off)
                rate = float(np.round(rng.normal(0.033, 0.006), 5))
                notional = float(np.round(np.exp(rng.uniform(np.log(1e5), np.log(1e8))), 2))
                side = "PAY" if rng.random() < 0.5 else "RECEIVE"
                lag = int(rng.integers(0, 3)) if k % 3 == 0 else 0
                z = ZeroCouponInflationSwap(eff, f"{tenor}Y", SwapTypes[side], rate, idx, notional, payment_lag=lag,
                                            dc_type=DayCountTypes.ACT_365F, bd_type=BusDayAdjustTypes.FOLLOWING)
                pv = z.value(vd, dc, ic)
                trades.append({"id": f"{iname}_{dname}_{n}", "index": iname, "discount": dname,
                               "effective": [eff._d, eff._m, eff._y], "tenor": f"{tenor}Y", "fixed_leg": side,
                               "fixed_rate": rate, "notional": notional, "payment_lag": lag, "value": float(pv),
                               "fixed_pv": float(z._fixed_pv), "inflation_pv": float(z._inflation_pv),
                               "payment_df": float(z._payment_df), "fixed_return": float(z._fixed_return),
                               "base_index": float(z._inflation_leg._base_index),
                               "final_index": float(z._inflation_leg._final_index),
                               "payment_dt": [z._payment_dt._d, z._payment_dt._m, z._payment_dt._y],
                               "breakeven": float(z.breakeven_inflation_rate(vd, dc, ic))})
                n += 1
    out["inflation_curves"] = infl_curves
    out["trades"] = trades
    # direct DF queries on both discount curves (DiscountCurve.df with ACT_365F dates)
    q = [vd.add_days(int(x)) for x in (0, 1, 45, 200, 365, 1000, 4000, 9000, 12000, 20000)]
    out["df_queries"] = {"dates": [[d._d, d._m, d._y] for d in q],
                         "dfs": {k: [float(c.df(d, DayCountTypes.ACT_365F)) for d in q] for k, c in dcurves.items()}}
    with open(os.path.join(OUT, "ref_zcis.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(trades), "ZCIS trades")


if __name__ == "__main__":
    main()
