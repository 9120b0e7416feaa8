#!/usr/bin/env python
"""Known answers of the reference's stand-alone bond / FRN analytics (cavour/trades/credit/bond.py:264-875, 1027-1110;
frn.py:225-573) from the UNMODIFIED reference, for the bonds of make_golden_bonds.py and the notes of make_golden_frn.py on the
path-A curves of the same model.  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:tests/golden/gen:/root/reference python tests/golden/gen/make_golden_credit_analytics.py

Writes tests/golden/ref_credit_analytics.json.
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.utils.currency import CurrencyTypes
from cavour.utils.day_count import DayCountTypes
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.calendar import BusDayAdjustTypes
from cavour.utils.global_types import CurveTypes
from cavour.market.curves.interpolator import InterpTypes
from cavour.trades.credit.bond import Bond
from cavour.trades.credit.frn import FRN
from cavour.models.models import Model

from make_golden import GBP_PX, USD_PX, TENORS
from make_golden_bonds import BONDS, VALUE_DT
from make_golden_frn import FRNS, DUAL

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SETTLE = (7, 5, 2024)          # a settlement date a week after the value date (accruals, relative DFs)


def main():
    vd, settle = Date(*VALUE_DT), Date(*SETTLE)
    model = Model(vd)
    for name, px in (("GBP_OIS_SONIA", GBP_PX), ("USD_OIS_SOFR", USD_PX)):
        model.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0,
                          fixed_dcc_type=DayCountTypes.ACT_365F, float_dc_type=DayCountTypes.ACT_365F,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
        c = getattr(model.curves, name)          # the non-AD methods query curve.df(): numpy node arrays
        c._times, c._dfs = np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)
    ois = {"GBP": model.curves.GBP_OIS_SONIA, "USD": model.curves.USD_OIS_SOFR}
    out = {"settle": SETTLE, "bonds": [], "frns": [],
           "equal_principal": {f"{face}/{n}": Bond.generate_equal_principal_schedule(face, n) for face, n in ((100.0, 4), (1e6, 7))},
           "annuity": {f"{face}/{n}/{cpn}/{fq}": Bond.generate_annuity_schedule(face, n, cpn, FrequencyTypes[fq])
                       for face, n, cpn, fq in ((100.0, 5, 0.04, "ANNUAL"), (1e6, 12, 0.05, "QUARTERLY"), (100.0, 3, 0.0, "ANNUAL"))}}
    for bid, issue, mat, cpn, freq, dc, ccy, face, lag, amort in BONDS:
        b = Bond(Date(*issue), mat if isinstance(mat, str) else Date(*mat), cpn, FrequencyTypes[freq], DayCountTypes[dc],
                 CurrencyTypes[ccy], face_value=face, payment_lag=lag, amortization_schedule=amort)
        c = ois[ccy]
        rec = {"id": bid}
        for tag, s in (("vd", vd), ("settle", settle)):
            clean = float(b.clean_price(vd, c, 0.0, s))
            rec[tag] = {
                "value": float(b.value(vd, c, 0.0, s)), "value_z75": float(b.value(vd, c, 0.0075, s)),
                "payment_dfs": [float(x) for x in b._payment_dfs], "coupon_pvs": [float(x) for x in b._coupon_pvs],
                "principal_pvs": [float(x) for x in b._principal_pvs],
                "accrued": float(b.accrued_interest(s)), "dirty": float(b.dirty_price(vd, c, 0.0, s)), "clean": clean,
                "clean_z75": float(b.clean_price(vd, c, 0.0075, s)),
                "ytm": float(b.yield_to_maturity(s, clean)), "ytm_minus2": float(b.yield_to_maturity(s, clean - 2.0)),
                "z_spread_minus2": float(b.z_spread(s, c, clean - 2.0)),
                "duration": float(b.duration(s, c)), "macaulay_z75": float(b.duration(s, c, "macaulay", 0.0075)),
                "convexity": float(b.convexity(s, c)), "dv01": float(b.dv01(s, c)), "cs01_z75": float(b.cs01(s, c, 0.0075)),
            }
        rec["current_yield"] = float(b.current_yield())
        if not b._is_zero_coupon:          # a zero-coupon bond's frequency has no zero-rate compounding
            rec["g_spread"] = float(b.g_spread(settle, ois["USD" if ccy == "GBP" else "GBP"], rec["settle"]["clean"] - 1.0))
            rec["i_spread"] = float(b.i_spread(settle, c, rec["settle"]["clean"] - 1.0))
        out["bonds"].append(rec)
        print(bid, rec["vd"]["clean"], rec["vd"]["ytm"], rec["settle"]["duration"], flush=True)
    for fid, issue, mat, margin, freq, dc, ccy, index, face, lag, fixing in FRNS + DUAL:
        f = FRN(Date(*issue), mat if isinstance(mat, str) else Date(*mat), margin, FrequencyTypes[freq], DayCountTypes[dc],
                CurrencyTypes[ccy], CurveTypes[index], face_value=face, payment_lag=lag, first_fixing_rate=fixing)
        disc, idx = ois[ccy], getattr(model.curves, index)
        rec = {"id": fid}
        for tag, s in (("vd", vd), ("settle", settle)):
            try:
                clean = float(f.clean_price(vd, disc, idx, 0.0, s))
                rec[tag] = {
                    "value": float(f.value(vd, disc, idx, 0.0, s)), "value_dm20": float(f.value(vd, disc, idx, 0.002, s)),
                    "rates": [float(x) for x in f._rates], "coupon_payments": [float(x) for x in f._coupon_payments],
                    "payment_dfs": [float(x) for x in f._payment_dfs], "payment_pvs": [float(x) for x in f._payment_pvs],
                    "dirty": float(f.dirty_price(vd, disc, idx, 0.0, s)), "accrued": float(f.accrued_interest(s)), "clean": clean,
                    "dm_minus_half": float(f.discount_margin(s, disc, idx, clean - 0.5)),
                    "mod_duration": float(f.modified_duration(vd, disc, idx, 0.0, s)),
                    "dv01_dm20": float(f.dv01(vd, disc, idx, 0.002, s)),
                }
            except Exception as ex:  # noqa: BLE001  (seasoned notes without a fixing look up a DF before the value date)
                rec[tag] = {"error": type(ex).__name__ + ": " + str(ex)}
        out["frns"].append(rec)
        print(fid, rec["vd"].get("clean"), rec["vd"].get("error"), flush=True)
    # capped / floored note, same curve for both roles
    f = FRN(Date(*VALUE_DT), "4Y", 0.003, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP,
            CurveTypes.GBP_OIS_SONIA, cap_rate=0.045, floor_rate=0.04)
    out["collar"] = {"value": float(f.value(vd, ois["GBP"])), "rates": [float(x) for x in f._rates],
                     "clean": float(f.clean_price(vd, ois["GBP"]))}
    with open(os.path.join(OUT, "ref_credit_analytics.json"), "w") as fh:
        json.dump(out, fh)
    print("done")


if __name__ == "__main__":
    main()
