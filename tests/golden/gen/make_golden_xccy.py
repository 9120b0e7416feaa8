#!/usr/bin/env python
"""Golden vectors for the cross-currency path, from the UNMODIFIED reference (see make_golden.py for how it is
run under the torch-backed jax stand-in):

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_xccy.py

Model with GBP SONIA (foreign) + USD SOFR (domestic) OIS curves and a GBP_USD_BASIS XccyCurve built by
Model.build_xccy_curve; XccyBasisSwap trades valued with Position(swap, model).compute([VALUE, DELTA]).
GAMMA is not generated: the reference raises in its cross-gamma einsum (engine.py:1936-1939 contracts the
65 path-A foreign nodes of XccyCurve._mixed_hess_foreign_basis with the 263-row engine Jacobian).
Output: tests/golden/ref_xccy.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import GBP_PX, USD_PX, TENORS, dmy  # noqa: E402

from cavour.models.models import Model  # noqa: E402
from cavour.utils import *  # noqa: F401,F403,E402
from cavour.utils.date import Date  # noqa: E402
from cavour.trades.rates.xccy_basis_swap import XccyBasisSwap  # noqa: E402
from cavour.utils.global_types import SwapTypes, CurveTypes, RequestTypes  # noqa: E402
from cavour.utils.currency import CurrencyTypes  # noqa: E402
from cavour.market.curves.interpolator import InterpTypes  # noqa: E402
from cavour.market.position.position import Position  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
VD = (30, 4, 2024)
BASIS_TENORS = ["1Y", "2Y", "3Y", "5Y", "7Y", "10Y"]
BASIS_BPS = [-5.0, -8.0, -10.0, -12.0, -13.5, -15.0]
SPOT = 1.25
TRADES = [
    # id, effective (None = value date, or ("bd", n)), tenor, dom notional, for notional, dom spread, for spread, dom freq, for freq
    ("x_4y_par_like", None, "4Y", 1.25e6, 1.0e6, 0.0, -0.0011, "ANNUAL", "QUARTERLY"),
    ("x_2y_calib", None, "2Y", 1.0e8, 8.0e7, 0.0, -0.0008, "ANNUAL", "QUARTERLY"),
    ("x_6y_spreads", None, "6Y", 2.5e6, 2.0e6, 0.0005, -0.0020, "ANNUAL", "QUARTERLY"),
    ("x_3y_fwd_start", ("bd", 20), "3Y", 1.25e6, 1.0e6, 0.0, -0.0010, "ANNUAL", "QUARTERLY"),
    ("x_9y_annual", None, "9Y", 5.0e6, 4.0e6, 0.0, -0.0014, "ANNUAL", "ANNUAL"),
]


def main():
    vd = Date(*VD)
    model = Model(vd)
    for name, px, dc in (("GBP_OIS_SONIA", GBP_PX, DayCountTypes.ACT_365F), ("USD_OIS_SOFR", USD_PX, DayCountTypes.ACT_360)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0, swap_type=SwapTypes.PAY,
                          fixed_dcc_type=dc, fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          float_dc_type=dc, bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                          interp_type=InterpTypes.LINEAR_ZERO_RATES)
        c = getattr(model.curves, name)   # non-AD df() needs numpy node arrays under the shim
        c._times = np.asarray(c._times, dtype=np.float64)
        c._dfs = np.asarray(c._dfs, dtype=np.float64)
    t0 = time.time()
    model.build_xccy_curve(name="GBP_USD_BASIS", domestic_curve_name="USD_OIS_SOFR", foreign_curve_name="GBP_OIS_SONIA",
                           basis_spreads=BASIS_BPS, tenor_list=BASIS_TENORS, spot_fx=SPOT,
                           domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.QUARTERLY)
    xc = model.curves.GBP_USD_BASIS
    print("xccy curve", len(xc._times), f"{time.time() - t0:.1f}s", flush=True)
    out = {
        "value_dt": list(VD), "gbp_px": GBP_PX, "usd_px": USD_PX, "tenors": TENORS,
        "basis_tenors": BASIS_TENORS, "basis_bps": BASIS_BPS, "spot_fx": SPOT,
        "xccy_times": np.asarray(xc._times).tolist(), "xccy_dfs": np.asarray(xc._dfs).tolist(),
        "xccy_swap_times": [float(x) for x in xc.swap_times],
        "xccy_jac_basis": np.asarray(xc._jac_basis).tolist(),
        "xccy_spot_fx_internal": float(xc._spot_fx),
        "xccy_interp": xc._interp_type.name,
        "trades": [],
    }
    for tid, eff, tenor, nd, nf, sd, sf, fd, ff in TRADES:
        t0 = time.time()
        eff_dt = vd if eff is None else vd.add_weekdays(eff[1])
        swap = XccyBasisSwap(effective_dt=eff_dt, term_dt_or_tenor=tenor, domestic_notional=nd, foreign_notional=nf,
                             domestic_spread=sd, foreign_spread=sf, domestic_freq_type=FrequencyTypes[fd],
                             foreign_freq_type=FrequencyTypes[ff], domestic_dc_type=DayCountTypes.ACT_360,
                             foreign_dc_type=DayCountTypes.ACT_365F, domestic_floating_index=CurveTypes.USD_OIS_SOFR,
                             foreign_floating_index=CurveTypes.GBP_OIS_SONIA, domestic_currency=CurrencyTypes.USD,
                             foreign_currency=CurrencyTypes.GBP)
        res = Position(swap, model).compute([RequestTypes.VALUE, RequestTypes.DELTA])
        deltas = {d.curve_type.name: {"ladder": [float(x) for x in np.asarray(d.risk_ladder)], "tenors": list(d.tenors)}
                  for d in res.risk._by_curve.values()}
        out["trades"].append({"id": tid, "effective": dmy(eff_dt), "tenor": tenor, "domestic_notional": nd,
                              "foreign_notional": nf, "domestic_spread": sd, "foreign_spread": sf, "domestic_freq": fd,
                              "foreign_freq": ff, "value": float(res.value.amount), "deltas": deltas})
        print(f"trade {tid}: pv={res.value.amount:.8g} {time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(OUT, "ref_xccy.json"), "w") as f:
        json.dump(out, f)
    print("done")


if __name__ == "__main__":
    main()
