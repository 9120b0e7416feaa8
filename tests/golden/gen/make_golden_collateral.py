#!/usr/bin/env python
"""Golden vectors for OIS with cross-currency collateral (Engine._compute_ois_xccy_collateral,
engine.py:217-503), from the UNMODIFIED reference under the torch-backed jax stand-in:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_collateral.py

Same model as make_golden_xccy.py, with the basis curve additionally registered as GBP_USD_XCCY (the name the
engine looks up for a GBP swap with USD collateral).  Output: tests/golden/ref_collateral.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import GBP_PX, USD_PX, TENORS, dmy  # noqa: E402
from make_golden_xccy import VD, BASIS_TENORS, BASIS_BPS, SPOT  # noqa: E402

from cavour.models.models import Model  # noqa: E402
from cavour.utils import *  # noqa: F401,F403,E402
from cavour.utils.date import Date  # noqa: E402
from cavour.trades.rates.ois import OIS  # noqa: E402
from cavour.utils.global_types import SwapTypes, CurveTypes, RequestTypes, CollateralType  # noqa: E402
from cavour.utils.currency import CurrencyTypes  # noqa: E402
from cavour.market.curves.interpolator import InterpTypes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
TRADES = [("c_5y_pay", None, "5Y", "PAY", 0.043, 2e6, "ANNUAL", "ANNUAL", 0.0),
          ("c_8y_rec_semi", None, "8Y", "RECEIVE", 0.041, 1e6, "SEMI_ANNUAL", "QUARTERLY", 0.0005),
          ("c_3y_fwd", ("bd", 15), "3Y", "PAY", 0.045, 5e6, "ANNUAL", "ANNUAL", 0.0)]


def main():
    vd = Date(*VD)
    model = Model(vd)
    for name, px, dc in (("GBP_OIS_SONIA", GBP_PX, DayCountTypes.ACT_365F), ("USD_OIS_SOFR", USD_PX, DayCountTypes.ACT_360)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0, swap_type=SwapTypes.PAY,
                          fixed_dcc_type=dc, fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          float_dc_type=dc, bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                          interp_type=InterpTypes.LINEAR_ZERO_RATES)
        c = getattr(model.curves, name)
        c._times = np.asarray(c._times, dtype=np.float64)
        c._dfs = np.asarray(c._dfs, dtype=np.float64)
    model.build_xccy_curve(name="GBP_USD_XCCY", domestic_curve_name="USD_OIS_SOFR", foreign_curve_name="GBP_OIS_SONIA",
                           basis_spreads=BASIS_BPS, tenor_list=BASIS_TENORS, spot_fx=SPOT,
                           domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.QUARTERLY)
    out = {"trades": []}
    for tid, eff, tenor, side, cpn, notl, ffreq, lfreq, spread in TRADES:
        eff_dt = vd if eff is None else vd.add_weekdays(eff[1])
        swap = OIS(effective_dt=eff_dt, term_dt_or_tenor=tenor, fixed_leg_type=SwapTypes[side], fixed_coupon=cpn,
                   fixed_freq_type=FrequencyTypes[ffreq], fixed_dc_type=DayCountTypes.ACT_365F,
                   floating_index=CurveTypes.GBP_OIS_SONIA, currency=CurrencyTypes.GBP, notional=notl,
                   float_spread=spread, float_freq_type=FrequencyTypes[lfreq], float_dc_type=DayCountTypes.ACT_365F,
                   bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        res = swap.position(model).compute([RequestTypes.VALUE, RequestTypes.DELTA], collateral_type=CollateralType.USD)
        deltas = {d.curve_type.name: {"ladder": [float(x) for x in np.asarray(d.risk_ladder)], "tenors": list(d.tenors)}
                  for d in res.risk._by_curve.values()}
        out["trades"].append({"id": tid, "effective": dmy(eff_dt), "tenor": tenor, "side": side, "coupon": cpn,
                              "notional": notl, "fixed_freq": ffreq, "float_freq": lfreq, "spread": spread,
                              "value": float(res.value.amount), "currency": res.value.currency.name, "deltas": deltas})
        print(tid, res.value.amount, flush=True)
    with open(os.path.join(OUT, "ref_collateral.json"), "w") as f:
        json.dump(out, f)
    print("done")


if __name__ == "__main__":
    main()
