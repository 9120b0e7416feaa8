"""Builders for the bond goldens (tests/golden/ref_bonds.json, produced by the unmodified reference engine)."""
from adrates_b200 import (Date, DayCountTypes, FrequencyTypes, BusDayAdjustTypes, SwapTypes, InterpTypes, CurrencyTypes,
                          Bond, FRN, CurveTypes)
from adrates_b200.models import Model


def build_bond_model(g):
    m = Model(Date(*g["value_dt"]))
    for name, px in (("GBP_OIS_SONIA", g["gbp_px"]), ("USD_OIS_SOFR", g["usd_px"])):
        m.build_curve(name=name, px_list=px, tenor_list=g["tenors"], spot_days=0, swap_type=SwapTypes.PAY,
                      fixed_dcc_type=DayCountTypes.ACT_365F, fixed_freq_type=FrequencyTypes.ANNUAL,
                      float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=DayCountTypes.ACT_365F,
                      bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    return m


def make_bond(b):
    mat = b["maturity"] if isinstance(b["maturity"], str) else Date(*b["maturity"])
    return Bond(Date(*b["issue"]), mat, b["coupon"], FrequencyTypes[b["freq"]], DayCountTypes[b["dc"]],
                CurrencyTypes[b["currency"]], face_value=b["face"], payment_lag=b["payment_lag"],
                amortization_schedule=b["amortization"])


def make_frn(f):
    mat = f["maturity"] if isinstance(f["maturity"], str) else Date(*f["maturity"])
    return FRN(Date(*f["issue"]), mat, f["margin"], FrequencyTypes[f["freq"]], DayCountTypes[f["dc"]],
               CurrencyTypes[f["currency"]], CurveTypes[f["index"]], face_value=f["face"], payment_lag=f["payment_lag"],
               first_fixing_rate=f["first_fixing"])
