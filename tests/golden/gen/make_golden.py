#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no
/root/reference):

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden.py

JAX is not installed here, so `refshim/jax` maps the jax/jnp calls the reference makes
onto torch float64 + torch.func (see its docstring).  Everything else executed is the
reference's own code: Model.build_curve, OISCurve (path A), Engine.build_curve_ad /
_cached_curve (path B + AD Jacobian/Hessian), Position.compute([VALUE, DELTA, GAMMA]),
Schedule / DayCount / Date / to_tenor.

Outputs (committed): tests/golden/ref_curves.json, ref_tables_*.npz, ref_trades.json,
ref_schedules.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch

from cavour.models.models import Model
from cavour.utils import *  # noqa: F401,F403
from cavour.utils.date import Date
from cavour.utils.helpers import to_tenor, times_from_dates
from cavour.utils.schedule import Schedule
from cavour.utils.day_count import DayCount, DayCountTypes
from cavour.utils.calendar import Calendar, CalendarTypes, BusDayAdjustTypes, DateGenRuleTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.global_types import SwapTypes, CurveTypes, RequestTypes
from cavour.utils.currency import CurrencyTypes
from cavour.trades.rates.ois import OIS
from cavour.market.curves.interpolator import InterpTypes
from cavour.market.position.engine import Engine

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

GBP_PX = [5.1998, 5.2014, 5.2003, 5.2027, 5.2023, 5.19281, 5.1656, 5.1482, 5.1342, 5.1173, 5.1013,
          5.0862, 5.0701, 5.054, 5.0394, 4.8707, 4.75483, 4.532, 4.3628, 4.2428, 4.16225, 4.1132,
          4.08505, 4.0762, 4.078, 4.0961, 4.12195, 4.1315, 4.113, 4.07724, 3.984, 3.88]
USD_PX = [5.3500, 5.3200, 5.3100, 5.2900, 5.2700, 5.2500, 5.2300, 5.2100, 5.1900, 5.1700, 5.1500,
          5.1300, 5.1100, 5.0900, 5.0700, 4.9500, 4.8500, 4.7000, 4.5800, 4.4800, 4.4100, 4.3600,
          4.3200, 4.2900, 4.2700, 4.2800, 4.3000, 4.3200, 4.3100, 4.2900, 4.2400, 4.1800]
TENORS = "1D 1W 2W 1M 2M 3M 4M 5M 6M 7M 8M 9M 10M 11M 1Y 18M 2Y 3Y 4Y 5Y 6Y 7Y 8Y 9Y 10Y 12Y 15Y 20Y 25Y 30Y 40Y 50Y".split()

CURVES = {
    # key: (curve name, value date dmy, px, freq, day count, interp)
    "gbp_readme_lzr": ("GBP_OIS_SONIA", (30, 4, 2024), GBP_PX, "ANNUAL", "ACT_365F", "LINEAR_ZERO_RATES"),
    "gbp_readme_ff": ("GBP_OIS_SONIA", (30, 4, 2024), GBP_PX, "ANNUAL", "ACT_365F", "FLAT_FWD_RATES"),
    "gbp_dec24_lzr": ("GBP_OIS_SONIA", (17, 12, 2024), GBP_PX, "ANNUAL", "ACT_365F", "LINEAR_ZERO_RATES"),
    "gbp_semi_lzr": ("GBP_OIS_SONIA", (17, 12, 2024), GBP_PX, "SEMI_ANNUAL", "ACT_365F", "LINEAR_ZERO_RATES"),
    "gbp_quarterly_lzr": ("GBP_OIS_SONIA", (17, 12, 2024), GBP_PX, "QUARTERLY", "ACT_365F", "LINEAR_ZERO_RATES"),
    "usd_dec24_lzr": ("USD_OIS_SOFR", (17, 12, 2024), USD_PX, "ANNUAL", "ACT_360", "LINEAR_ZERO_RATES"),
}
TABLE_CURVES = {"gbp_readme_lzr", "usd_dec24_lzr", "gbp_semi_lzr"}

# trade specs: curve key, effective (dmy or ("bd", n) business days after value date), tenor or
# termination dmy, PAY/RECEIVE, coupon, notional, fixed freq, float freq, spread, payment lag
TRADES = [
    ("nb_1w_par", "gbp_readme_lzr", None, "1W", "PAY", 0.052014, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("readme_10y", "gbp_readme_lzr", None, "10Y", "PAY", 0.045, 1e7, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_2y_rec", "gbp_readme_lzr", None, "2Y", "RECEIVE", 0.047, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_18m", "gbp_readme_lzr", None, "18M", "PAY", 0.049, 2.5e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_3m", "gbp_readme_lzr", None, "3M", "RECEIVE", 0.052, 5e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_30y_rec", "gbp_readme_lzr", None, "30Y", "RECEIVE", 0.0405, 3e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_50y_pay", "gbp_readme_lzr", None, "50Y", "PAY", 0.0388, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("on_13y_pay", "gbp_readme_lzr", None, "13Y", "PAY", 0.041, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("off_7y_rec", "gbp_readme_lzr", (17, 6, 2024), "7Y", "RECEIVE", 0.041, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("off_12y_pay", "gbp_readme_lzr", ("bd", 3), "12Y", "PAY", 0.0412, 4e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("off_1y_pay", "gbp_readme_lzr", ("bd", 100), "1Y", "PAY", 0.05, 4e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("off_49y_rec", "gbp_readme_lzr", ("bd", 200), "49Y", "RECEIVE", 0.039, 2e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("beyond_55y", "gbp_readme_lzr", None, "55Y", "RECEIVE", 0.039, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("semi_quart_5y", "gbp_readme_lzr", None, "5Y", "PAY", 0.0425, 1e6, "SEMI_ANNUAL", "QUARTERLY", 0.0, 0),
    ("spread_4y", "gbp_readme_lzr", None, "4Y", "RECEIVE", 0.044, 1e6, "ANNUAL", "ANNUAL", 0.0025, 0),
    ("lag2_6y", "gbp_readme_lzr", None, "6Y", "PAY", 0.0417, 1e6, "ANNUAL", "ANNUAL", 0.0, 2),
    ("ff_on_10y", "gbp_readme_ff", None, "10Y", "PAY", 0.045, 1e7, "ANNUAL", "ANNUAL", 0.0, 0),
    ("ff_off_7y", "gbp_readme_ff", (17, 6, 2024), "7Y", "RECEIVE", 0.041, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("ff_beyond_55y", "gbp_readme_ff", None, "55Y", "RECEIVE", 0.039, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("ff_off_quart_3y", "gbp_readme_ff", ("bd", 30), "3Y", "PAY", 0.045, 1e6, "QUARTERLY", "QUARTERLY", 0.001, 0),
    ("dec_on_5y", "gbp_dec24_lzr", None, "5Y", "PAY", 0.0424, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("dec_off_20y", "gbp_dec24_lzr", ("bd", 41), "20Y", "RECEIVE", 0.0413, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("semi_on_10y", "gbp_semi_lzr", None, "10Y", "PAY", 0.0408, 1e6, "SEMI_ANNUAL", "SEMI_ANNUAL", 0.0, 0),
    ("semi_off_3y", "gbp_semi_lzr", ("bd", 17), "3Y", "RECEIVE", 0.045, 1e6, "SEMI_ANNUAL", "SEMI_ANNUAL", 0.0, 0),
    ("quart_on_2y", "gbp_quarterly_lzr", None, "2Y", "PAY", 0.0475, 1e6, "QUARTERLY", "QUARTERLY", 0.0, 0),
    ("usd_on_5y", "usd_dec24_lzr", None, "5Y", "PAY", 0.0448, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
    ("usd_off_10y", "usd_dec24_lzr", ("bd", 7), "10Y", "RECEIVE", 0.0427, 1e6, "ANNUAL", "ANNUAL", 0.0, 0),
]


def dmy(dt):
    return [dt.d(), dt.m(), dt.y()]


def build_model(key):
    name, vd, px, freq, dc, interp = CURVES[key]
    model = Model(Date(*vd))
    model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0, swap_type=SwapTypes.PAY,
                      fixed_dcc_type=DayCountTypes[dc], fixed_freq_type=FrequencyTypes[freq],
                      float_freq_type=FrequencyTypes[freq], float_dc_type=DayCountTypes[dc],
                      bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes[interp])
    return model


def main():
    t00 = time.time()
    curves_out, caches, models = {}, {}, {}
    for key, (name, vd, px, freq, dc, interp) in CURVES.items():
        t0 = time.time()
        model = build_model(key)
        models[key] = model
        curve = getattr(model.curves, name)
        eng = Engine(model)
        ck = tuple(curve.swap_times)
        cache = eng._cached_curve(ck, curve.swap_rates, curve.swap_times, curve.year_fracs, curve._interp_type)
        caches[key] = {ck: cache}
        pa_t = np.asarray(curve._times, dtype=np.float64)
        pa_d = np.asarray(curve._dfs, dtype=np.float64)
        ad_q = [0.0, 0.01, 0.25, 1.0, 1.3, 5.0, 7.77, 10.0, 33.3, 50.0]
        df_ad = [float(curve.df_ad(q)) for q in ad_q]
        # non-AD df() needs numpy node arrays (torch tensors have no `.size` attribute)
        curve._times, curve._dfs = pa_t, pa_d
        q_dts = [Date(*vd).add_tenor(t) for t in ["1M", "9M", "2Y", "7Y", "11Y", "31Y", "60Y"]]
        df_dates = [float(curve.df(q, DayCountTypes[dc])) for q in q_dts]
        curves_out[key] = {
            "name": name, "value_dt": list(vd), "px": px, "tenors": TENORS, "freq": freq, "dc": dc, "interp": interp,
            "swap_rates": [float(x) for x in curve.swap_rates],
            "swap_times": [float(x) for x in curve.swap_times],
            "year_fracs": [[float(y) for y in yf] for yf in curve.year_fracs],
            "tenor_labels": to_tenor(curve.swap_times),
            "pathA_times": pa_t.tolist(), "pathA_dfs": pa_d.tolist(),
            "pathB_times": np.asarray(cache["times"]).tolist(), "pathB_dfs": np.asarray(cache["dfs"]).tolist(),
            "df_ad_t": ad_q, "df_ad": df_ad,
            "df_dates": [dmy(q) for q in q_dts], "df": df_dates,
        }
        if key in TABLE_CURVES:
            np.savez_compressed(os.path.join(OUT, f"ref_tables_{key}.npz"),
                                jac=np.asarray(cache["jac"]), hess=np.asarray(cache["hess"]))
        print(f"curve {key}: G={len(cache['times'])} pathA={len(pa_t)} {time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(OUT, "ref_curves.json"), "w") as f:
        json.dump(curves_out, f)

    trades_out = []
    for (tid, ckey, eff, tenor, side, cpn, notl, ffreq, lfreq, spread, lag) in TRADES:
        t0 = time.time()
        name, vd, px, freq, dc, interp = CURVES[ckey]
        model = models[ckey]
        value_dt = Date(*vd)
        if eff is None:
            eff_dt = value_dt
        elif eff[0] == "bd":
            eff_dt = value_dt.add_weekdays(eff[1])
        else:
            eff_dt = Date(*eff)
        swap = OIS(effective_dt=eff_dt, term_dt_or_tenor=tenor, fixed_leg_type=SwapTypes[side], fixed_coupon=cpn,
                   fixed_freq_type=FrequencyTypes[ffreq], fixed_dc_type=DayCountTypes[dc],
                   floating_index=CurveTypes[name], currency=CurrencyTypes[name[:3]], notional=notl,
                   payment_lag=lag, float_spread=spread, float_freq_type=FrequencyTypes[lfreq],
                   float_dc_type=DayCountTypes[dc], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        pos = swap.position(model)
        pos._engine._curve_cache = caches[ckey]  # the engine's own cache object, built above by the engine itself
        res = pos.compute([RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA])
        trades_out.append({
            "id": tid, "curve": ckey, "effective": dmy(eff_dt), "tenor": tenor, "side": side, "coupon": cpn,
            "notional": notl, "fixed_freq": ffreq, "float_freq": lfreq, "spread": spread, "payment_lag": lag,
            "fixed_payment_dts": [dmy(d) for d in swap._fixed_leg._payment_dts],
            "fixed_payments": [float(x) for x in swap._fixed_leg._payments],
            "float_year_fracs": [float(x) for x in swap._float_leg._year_fracs],
            "value": float(res.value.amount),
            "delta": [float(x) for x in np.asarray(res.risk.risk_ladder)],
            "delta_tenors": list(res.risk.tenors),
            "gamma": np.asarray(res.gamma.risk_ladder, dtype=np.float64).tolist(),
        })
        print(f"trade {tid}: pv={trades_out[-1]['value']:.10g} {time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(OUT, "ref_trades.json"), "w") as f:
        json.dump(trades_out, f)

    # ---------------- host-side fixtures: dates, schedules, day counts ----------------
    sched = []
    effs = [(30, 4, 2024), (17, 12, 2024), (29, 2, 2024), (31, 8, 2023), (31, 1, 2025), (15, 6, 2024), (28, 2, 2025)]
    tens = ["1D", "1W", "2W", "1M", "3M", "11M", "1Y", "18M", "2Y", "5Y", "13Y", "30Y", "50Y"]
    for e in effs:
        for tn in tens:
            for fq in ["ANNUAL", "SEMI_ANNUAL", "QUARTERLY", "MONTHLY"]:
                for bd in ["MODIFIED_FOLLOWING", "FOLLOWING", "PRECEDING", "NONE"]:
                    for dg in ["BACKWARD", "FORWARD"]:
                        if fq == "MONTHLY" and tn in ("30Y", "50Y"):
                            continue
                        if dg == "FORWARD" and bd != "MODIFIED_FOLLOWING":
                            continue
                        ed = Date(*e)
                        td = ed.add_tenor(tn)
                        try:
                            s = Schedule(ed, td, FrequencyTypes[fq], CalendarTypes.WEEKEND, BusDayAdjustTypes[bd],
                                         DateGenRuleTypes[dg])
                            dts = [dmy(d) for d in s._adjusted_dts]
                        except Exception as ex:  # noqa: BLE001
                            dts = "ERR:" + type(ex).__name__
                        sched.append({"eff": list(e), "tenor": tn, "term": dmy(td), "freq": fq, "bd": bd, "dg": dg,
                                      "dates": dts})
    dcs = []
    rng = np.random.default_rng(5)
    for _ in range(300):
        d1 = Date(1, 1, 2020).add_days(int(rng.integers(0, 12000)))
        d2 = d1.add_days(int(rng.integers(0, 4000)))
        row = {"d1": dmy(d1), "d2": dmy(d2), "serial_diff": int(d2 - d1), "wd1": int(d1.weekday())}
        for dc in ["ACT_365F", "ACT_360", "THIRTY_E_360", "THIRTY_360_BOND", "THIRTY_E_360_ISDA", "ACT_ACT_ISDA",
                   "THIRTY_E_PLUS_360", "SIMPLE"]:
            row[dc] = float(DayCount(DayCountTypes[dc]).year_frac(d1, d2)[0])
        row["add_wd_5"] = dmy(d1.add_weekdays(5))
        row["add_wd_m3"] = dmy(d1.add_weekdays(-3))
        row["add_months_m7"] = dmy(d1.add_months(-7))
        row["adj_mf"] = dmy(Calendar(CalendarTypes.WEEKEND).adjust(d1, BusDayAdjustTypes.MODIFIED_FOLLOWING))
        row["adj_mp"] = dmy(Calendar(CalendarTypes.WEEKEND).adjust(d1, BusDayAdjustTypes.MODIFIED_PRECEDING))
        dcs.append(row)
    tt = [0.001, 0.0027, 0.019, 0.02, 0.08, 0.0833, 0.0834, 0.1, 0.25, 0.5, 0.9167, 0.96, 0.999, 1.0, 1.04, 1.4583, 1.5,
          1.96, 2.0, 9.99, 10.0, 12.0082, 50.0329]
    with open(os.path.join(OUT, "ref_schedules.json"), "w") as f:
        json.dump({"schedules": sched, "daycounts": dcs, "to_tenor_in": tt, "to_tenor_out": to_tenor(tt)}, f)
    print("done", time.time() - t00)


if __name__ == "__main__":
    main()
