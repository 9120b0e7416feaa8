#!/usr/bin/env python
"""CASHFLOWS-request goldens for bonds and floating-rate notes from the UNMODIFIED reference (Engine._compute_bond :648-696,
Engine._compute_frn :930-983; Bond.value / FRN.value on the path-A curves).  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_cashflows_credit.py

The instruments are those of make_golden_bonds.py / make_golden_frn.py (incl. the dual-curve notes); writes
tests/golden/ref_cashflows_credit.json.
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
from cavour.trades.credit.bond import Bond
from cavour.trades.credit.frn import FRN
from cavour.models.models import Model

from make_golden import GBP_PX, USD_PX, TENORS
from make_golden_bonds import BONDS, VALUE_DT
from make_golden_frn import FRNS, DUAL

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def rows_of(cf):
    return [{"payment_date": [c.payment_date._d, c.payment_date._m, c.payment_date._y], "notional": float(c.notional),
             "payment_fraction": float(c.payment_fraction), "accrual_period": float(c.accrual_period), "amount": float(c.amount),
             "discount_factor": float(c.discount_factor), "discounted_amount": float(c.discounted_amount), "leg_type": c.leg_type}
            for c in cf.cashflows]


def main():
    vd = Date(*VALUE_DT)
    model = Model(vd)
    for name, px in (("GBP_OIS_SONIA", GBP_PX), ("USD_OIS_SOFR", USD_PX)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
        curve = getattr(model.curves, name)
        curve._times, curve._dfs = np.asarray(curve._times, dtype=np.float64), np.asarray(curve._dfs, dtype=np.float64)
    out = {"value_dt": VALUE_DT, "gbp_px": GBP_PX, "usd_px": USD_PX, "tenors": TENORS, "bonds": [], "frns": []}
    for bid, issue, mat, cpn, freq, dc, ccy, face, lag, amort in BONDS:
        b = Bond(Date(*issue), mat if isinstance(mat, str) else Date(*mat), cpn, FrequencyTypes[freq], DayCountTypes[dc],
                 CurrencyTypes[ccy], face_value=face, payment_lag=lag, amortization_schedule=amort)
        rec = {"id": bid, "issue": issue, "maturity": mat, "coupon": cpn, "freq": freq, "dc": dc, "currency": ccy, "face": face,
               "payment_lag": lag, "amortization": amort}
        try:
            cf = b.position(model).compute([RequestTypes.CASHFLOWS]).cashflows
            rec.update(rows=rows_of(cf), total_amount=float(cf.total_amount), total_pv=float(cf.total_pv), repr=repr(cf))
        except Exception as ex:  # noqa: BLE001
            rec["error"] = type(ex).__name__ + ": " + str(ex)
        out["bonds"].append(rec)
        print(bid, rec.get("total_pv"), rec.get("error"))
    for fid, issue, mat, margin, freq, dc, ccy, index, face, lag, fixing in FRNS + DUAL:
        f = FRN(Date(*issue), mat if isinstance(mat, str) else Date(*mat), margin, FrequencyTypes[freq], DayCountTypes[dc],
                CurrencyTypes[ccy], CurveTypes[index], face_value=face, payment_lag=lag, first_fixing_rate=fixing)
        rec = {"id": fid, "issue": issue, "maturity": mat, "margin": margin, "freq": freq, "dc": dc, "currency": ccy, "index": index,
               "face": face, "payment_lag": lag, "first_fixing": fixing}
        try:
            cf = f.position(model).compute([RequestTypes.CASHFLOWS]).cashflows
            rec.update(rows=rows_of(cf), total_amount=float(cf.total_amount), total_pv=float(cf.total_pv), repr=repr(cf))
        except Exception as ex:  # noqa: BLE001
            rec["error"] = type(ex).__name__ + ": " + str(ex)
        out["frns"].append(rec)
        print(fid, rec.get("total_pv"), rec.get("error"))
    with open(os.path.join(OUT, "ref_cashflows_credit.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
