#!/usr/bin/env python
"""Golden vectors for bonds from the UNMODIFIED reference (/root/reference): Position(bond, model).compute(
[VALUE, DELTA, GAMMA]) = Engine._compute_bond (engine.py:505-640).

TEST INFRASTRUCTURE; build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_bonds.py
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.global_types import RequestTypes
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.trades.credit.bond import Bond
from cavour.models.models import Model

from make_golden import GBP_PX, USD_PX, TENORS

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
VALUE_DT = (30, 4, 2024)

# id, issue (d, m, y), maturity (tenor or date), coupon, freq, day count, currency, face, payment lag, amortisation
BONDS = [
    ("gbp_5y_semi", (30, 4, 2024), "5Y", 0.04, "SEMI_ANNUAL", "ACT_365F", "GBP", 100.0, 0, None),
    ("gbp_10y_annual_1m", (30, 4, 2024), "10Y", 0.0425, "ANNUAL", "ACT_365F", "GBP", 1_000_000.0, 0, None),
    ("gbp_seasoned_quarterly", (17, 1, 2022), (17, 1, 2031), 0.035, "QUARTERLY", "ACT_365F", "GBP", 100.0, 0, None),
    ("gbp_forward_issue", (15, 8, 2024), "7Y", 0.045, "SEMI_ANNUAL", "THIRTY_E_360", "GBP", 100.0, 0, None),
    ("gbp_zero_coupon", (30, 4, 2024), "12Y", 0.0, "ANNUAL", "ACT_365F", "GBP", 100.0, 0, None),
    ("gbp_30y_lag2", (30, 4, 2024), "30Y", 0.05, "SEMI_ANNUAL", "ACT_ACT_ISDA", "GBP", 100.0, 2, None),
    ("gbp_amortising", (30, 4, 2024), "5Y", 0.04, "ANNUAL", "ACT_365F", "GBP", 100.0, 0, [80.0, 60.0, 40.0, 20.0, 0.0]),
    ("usd_3y_semi", (30, 4, 2024), "3Y", 0.0475, "SEMI_ANNUAL", "ACT_360", "USD", 100.0, 0, None),
    ("usd_odd_dates", (11, 3, 2023), (23, 9, 2041), 0.039, "SEMI_ANNUAL", "THIRTY_E_360", "USD", 5_000_000.0, 0, None),
]


def main():
    vd = Date(*VALUE_DT)
    model = Model(vd)
    for name, px in (("GBP_OIS_SONIA", GBP_PX), ("USD_OIS_SOFR", USD_PX)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    out = {"value_dt": VALUE_DT, "gbp_px": GBP_PX, "usd_px": USD_PX, "tenors": TENORS, "bonds": []}
    reqs = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
    for bid, issue, mat, cpn, freq, dc, ccy, face, lag, amort in BONDS:
        b = Bond(Date(*issue), mat if isinstance(mat, str) else Date(*mat), cpn, FrequencyTypes[freq], DayCountTypes[dc],
                 CurrencyTypes[ccy], face_value=face, payment_lag=lag, amortization_schedule=amort)
        res = b.position(model).compute(reqs)
        out["bonds"].append({
            "id": bid, "issue": issue, "maturity": mat, "coupon": cpn, "freq": freq, "dc": dc, "currency": ccy,
            "face": face, "payment_lag": lag, "amortization": amort,
            "payment_dts": [[d._d, d._m, d._y] for d in b._payment_dts],
            "coupon_payments": [float(x) for x in b._coupon_payments],
            "value": float(res.value.amount),
            "delta": [float(x) for x in np.asarray(res.risk.risk_ladder)],
            "tenors": list(res.risk.tenors),
            "gamma": np.asarray(res.gamma.risk_ladder, dtype=np.float64).tolist()})
        print(bid, out["bonds"][-1]["value"])
    with open(os.path.join(OUT, "ref_bonds.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
