#!/usr/bin/env python
"""Differential run of the ORACLE (oracle/cavour_oracle.py, the checker the GPU parity tests trust) against the unmodified
reference ENGINE on random OIS trades: `Position.compute([VALUE, DELTA, GAMMA])` of the reference (jax grad / hessian through the
torch-backed stand-in) beside `ois_analytics` of the oracle fed by THIS package's object layer (schedules, day counts, leg
arrays).  Build container only; TEST INFRASTRUCTURE.

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tools/reftests/differential_engine.py [seed] [n]

Trades: random effective date (on the value date, or days to years after it: off-grid cashflows), tenor 1Y-50Y, either side,
coupon, notional, fixed / floating frequencies, floating spread, payment lag, holiday calendar, business-day rule; curves:
the README SONIA curve with LINEAR_ZERO_RATES and with FLAT_FWD_RATES.  Metric: |x - ref| / max(|ref|, natural scale) as in
the parity tests (PV ~ notional, delta ~ notional x 1e-4 x T, gamma ~ notional x 1e-8 x T^2).
"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import make_golden as mg                                                              # noqa: E402
from cavour.utils.date import Date as RDate                                            # noqa: E402
from cavour.utils.calendar import BusDayAdjustTypes as RBd, CalendarTypes as RCal      # noqa: E402
from cavour.utils.day_count import DayCountTypes as RDC                                # noqa: E402
from cavour.utils.frequency import FrequencyTypes as RFreq                             # noqa: E402
from cavour.utils.global_types import SwapTypes as RSwap, CurveTypes as RCurve, RequestTypes as RReq   # noqa: E402
from cavour.utils.currency import CurrencyTypes as RCcy                                # noqa: E402
from cavour.trades.rates.ois import OIS as ROIS                                        # noqa: E402
from cavour.market.position.engine import Engine as REngine                            # noqa: E402

import adrates_b200 as O                                                               # noqa: E402
from oracle import cavour_oracle as orc                                                # noqa: E402
from tests.util_trades import leg_arrays, rel_err                                      # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20240430
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    rng = random.Random(seed)
    worst = [0.0, 0.0, 0.0]
    t0 = time.time()
    for ckey in ("gbp_readme_lzr", "gbp_readme_ff"):
        name, vd, px, freq, dc, interp = mg.CURVES[ckey]
        rmodel = mg.build_model(ckey)
        rcurve = getattr(rmodel.curves, name)
        key = tuple(rcurve.swap_times)
        cache = {key: REngine(rmodel)._cached_curve(key, rcurve.swap_rates, rcurve.swap_times, rcurve.year_fracs, rcurve._interp_type)}
        omodel = O.Model(O.Date(*vd))
        omodel.build_curve(name=name, px_list=px, tenor_list=mg.TENORS, spot_days=0, swap_type=O.SwapTypes.PAY,
                           fixed_dcc_type=O.DayCountTypes[dc], fixed_freq_type=O.FrequencyTypes[freq], float_freq_type=O.FrequencyTypes[freq],
                           float_dc_type=O.DayCountTypes[dc], bus_day_type=O.BusDayAdjustTypes.MODIFIED_FOLLOWING,
                           interp_type=O.InterpTypes[interp])
        ocurve = omodel.curves[name]
        plan = orc.plan_path_b(ocurve.swap_times, ocurve.year_fracs)
        tables = (plan["times"],) + tuple(orc.bootstrap_tables(ocurve.swap_rates, plan))
        for _ in range(n // 2):
            off = rng.choice([0, 0, rng.randint(1, 30), rng.randint(31, 400), rng.randint(401, 2000)])
            tenor = rng.choice(["1Y", "2Y", "3Y", "5Y", "7Y", "10Y", "12Y", "18M", "20Y", "30Y", "50Y"])
            side, cpn, N = rng.choice(["PAY", "RECEIVE"]), round(rng.uniform(0.005, 0.09), 4), rng.choice([1e5, 1e6, 2.5e7])
            ff, lf = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"]), rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"])
            spr, lag = rng.choice([0.0, 0.0, 0.0012]), rng.choice([0, 0, 2])
            cal, bd = rng.choice(["WEEKEND", "WEEKEND", "UNITED_KINGDOM", "TARGET"]), rng.choice(["MODIFIED_FOLLOWING", "FOLLOWING"])
            reff, oeff = RDate(*vd).add_weekdays(off) if off else RDate(*vd), O.Date(*vd).add_weekdays(off) if off else O.Date(*vd)
            r = ROIS(reff, tenor, RSwap[side], cpn, RFreq[ff], RDC[dc], RCurve[name], RCcy.GBP, N, lag, spr, RFreq[lf], RDC[dc], RCal[cal], RBd[bd])
            o = O.OIS(oeff, tenor, O.SwapTypes[side], cpn, O.FrequencyTypes[ff], O.DayCountTypes[dc], O.CurveTypes[name], O.CurrencyTypes.GBP, N,
                      lag, spr, O.FrequencyTypes[lf], O.DayCountTypes[dc], O.CalendarTypes[cal], O.BusDayAdjustTypes[bd])
            pos = r.position(rmodel)
            pos._engine._curve_cache = cache
            res = pos.compute([RReq.VALUE, RReq.DELTA, RReq.GAMMA])
            fixed, floating = leg_arrays(o, omodel.value_dt)
            pv, dl, gm = orc.ois_analytics(tables, ocurve._interp_type.value, fixed, floating)
            T = max(1.0, float(tenor[:-1]) / (12.0 if tenor.endswith("M") else 1.0))
            e = (rel_err(pv, float(res.value.amount), N), rel_err(dl, np.asarray(res.risk.risk_ladder, dtype=np.float64), N * 1e-4 * T),
                 rel_err(gm, np.asarray(res.gamma.risk_ladder, dtype=np.float64), N * 1e-8 * T * T))
            worst = [max(a, b) for a, b in zip(worst, e)]
            print(f"{ckey} +{off}bd {tenor} {side} {ff}/{lf} lag{lag} {cal} {bd}: pv {e[0]:.1e} delta {e[1]:.1e} gamma {e[2]:.1e}", flush=True)
    print(f"seed {seed}: {n} trades, worst pv / delta / gamma error {worst[0]:.1e} / {worst[1]:.1e} / {worst[2]:.1e} (gate 1e-10), {time.time() - t0:.0f} s")
    return 0 if max(worst) < 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main())
