#!/usr/bin/env python
"""Golden vectors for floating-rate notes from the UNMODIFIED reference (/root/reference): Position(frn, model)
.compute([VALUE, DELTA, GAMMA]) = Engine._compute_frn, single-curve case (engine.py:700-925).

TEST INFRASTRUCTURE; build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_frn.py
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.global_types import RequestTypes, CurveTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.trades.credit.frn import FRN
from cavour.models.models import Model

from make_golden import GBP_PX, USD_PX, TENORS

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
VALUE_DT = (30, 4, 2024)

# id, issue, maturity, quoted margin, freq, day count, currency, index, face, payment lag, first fixing
FRNS = [
    ("gbp_5y_quarterly", (30, 4, 2024), "5Y", 0.005, "QUARTERLY", "ACT_365F", "GBP", "GBP_OIS_SONIA", 100.0, 0, None),
    ("gbp_3y_semi_fixing", (30, 4, 2024), "3Y", 0.0025, "SEMI_ANNUAL", "ACT_365F", "GBP", "GBP_OIS_SONIA", 1_000_000.0, 0, 0.0519),
    ("gbp_seasoned_fixing", (15, 2, 2023), (15, 2, 2030), 0.0075, "QUARTERLY", "ACT_365F", "GBP", "GBP_OIS_SONIA", 100.0, 0, 0.051),
    ("gbp_forward_issue_lag2", (16, 9, 2024), "10Y", 0.004, "SEMI_ANNUAL", "ACT_360", "GBP", "GBP_OIS_SONIA", 100.0, 2, None),
    ("gbp_zero_margin_annual", (30, 4, 2024), "7Y", 0.0, "ANNUAL", "ACT_365F", "GBP", "GBP_OIS_SONIA", 100.0, 0, None),
    ("usd_2y_quarterly", (30, 4, 2024), "2Y", 0.0035, "QUARTERLY", "ACT_360", "USD", "USD_OIS_SOFR", 100.0, 0, None),
]


# discounted on the currency's OIS curve, projected on the other one
DUAL = [
    ("gbp_on_sofr_5y_quarterly", (30, 4, 2024), "5Y", 0.005, "QUARTERLY", "ACT_365F", "GBP", "USD_OIS_SOFR", 100.0, 0, None),
    ("usd_on_sonia_3y_semi_fixing", (30, 4, 2024), "3Y", 0.0025, "SEMI_ANNUAL", "ACT_360", "USD", "GBP_OIS_SONIA", 1_000_000.0, 0, 0.0519),
    ("gbp_on_sofr_seasoned_lag2", (15, 2, 2023), (15, 2, 2031), 0.0075, "QUARTERLY", "ACT_365F", "GBP", "USD_OIS_SOFR", 100.0, 2, None),
    ("gbp_on_sofr_forward_10y", (16, 9, 2024), "10Y", 0.0, "ANNUAL", "ACT_365F", "GBP", "USD_OIS_SOFR", 250_000.0, 0, None),
]


def main():
    vd = Date(*VALUE_DT)
    model = Model(vd)
    for name, px in (("GBP_OIS_SONIA", GBP_PX), ("USD_OIS_SOFR", USD_PX)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    out = {"value_dt": VALUE_DT, "gbp_px": GBP_PX, "usd_px": USD_PX, "tenors": TENORS, "frns": []}
    reqs = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
    for fid, issue, mat, margin, freq, dc, ccy, index, face, lag, fixing in FRNS:
        f = FRN(Date(*issue), mat if isinstance(mat, str) else Date(*mat), margin, FrequencyTypes[freq], DayCountTypes[dc],
                CurrencyTypes[ccy], CurveTypes[index], face_value=face, payment_lag=lag, first_fixing_rate=fixing)
        res = f.position(model).compute(reqs)
        out["frns"].append({
            "id": fid, "issue": issue, "maturity": mat, "margin": margin, "freq": freq, "dc": dc, "currency": ccy,
            "index": index, "face": face, "payment_lag": lag, "first_fixing": fixing,
            "payment_dts": [[d._d, d._m, d._y] for d in f._payment_dts],
            "year_fracs": [float(x) for x in f._year_fracs],
            "value": float(res.value.amount),
            "delta": [float(x) for x in np.asarray(res.risk.risk_ladder)],
            "tenors": list(res.risk.tenors),
            "gamma": np.asarray(res.gamma.risk_ladder, dtype=np.float64).tolist()})
        print(fid, out["frns"][-1]["value"])
    # dual-curve notes (index curve != the currency's discount curve): the reference values them (VALUE only;
    # DELTA / GAMMA raise LibError, engine.py:921-924)
    out["dual"] = []
    for fid, issue, mat, margin, freq, dc, ccy, index, face, lag, fixing in DUAL:
        f = FRN(Date(*issue), mat if isinstance(mat, str) else Date(*mat), margin, FrequencyTypes[freq], DayCountTypes[dc],
                CurrencyTypes[ccy], CurveTypes[index], face_value=face, payment_lag=lag, first_fixing_rate=fixing)
        res = f.position(model).compute([RequestTypes.VALUE])
        raised = False
        try:
            f.position(model).compute([RequestTypes.VALUE, RequestTypes.DELTA])
        except Exception as ex:          # noqa: BLE001
            raised = type(ex).__name__
        out["dual"].append({"id": fid, "issue": issue, "maturity": mat, "margin": margin, "freq": freq, "dc": dc,
                            "currency": ccy, "index": index, "face": face, "payment_lag": lag, "first_fixing": fixing,
                            "value": float(res.value.amount), "delta_raises": raised})
        print(fid, out["dual"][-1]["value"], raised)
    # The engine caches curve tables by tuple(swap_times) only (engine.py:2362-2376), and a Position's Engine builds the
    # discount curve first: when both curves quote the same pillar dates the "index" lookup HITS the discount curve's
    # entry, so the notes above are in fact valued single-curve.  A model whose SOFR curve quotes fewer pillars has a
    # different key and exercises the genuine dual-curve arithmetic:
    model2 = Model(vd)
    model2.build_curve(name="GBP_OIS_SONIA", px_list=GBP_PX, tenor_list=TENORS, spot_days=0,
                       fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                       fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                       bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    keep = [i for i in range(len(TENORS)) if i not in (0, 2, 5)]
    model2.build_curve(name="USD_OIS_SOFR", px_list=[USD_PX[i] for i in keep], tenor_list=[TENORS[i] for i in keep],
                       spot_days=0, fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                       fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                       bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    out["dual_distinct_usd_keep"] = keep
    out["dual_distinct"] = []
    for fid, issue, mat, margin, freq, dc, ccy, index, face, lag, fixing in DUAL:
        if ccy != "GBP":
            continue
        f = FRN(Date(*issue), mat if isinstance(mat, str) else Date(*mat), margin, FrequencyTypes[freq], DayCountTypes[dc],
                CurrencyTypes[ccy], CurveTypes[index], face_value=face, payment_lag=lag, first_fixing_rate=fixing)
        res = f.position(model2).compute([RequestTypes.VALUE])
        out["dual_distinct"].append({"id": fid, "issue": issue, "maturity": mat, "margin": margin, "freq": freq, "dc": dc,
                                     "currency": ccy, "index": index, "face": face, "payment_lag": lag,
                                     "first_fixing": fixing, "value": float(res.value.amount)})
        print("distinct", fid, out["dual_distinct"][-1]["value"])
    with open(os.path.join(OUT, "ref_frn.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
