"""Shared by the CPU and GPU CASHFLOWS tests: golden loader, trade construction, row comparison."""
import numpy as np

from adrates_b200 import RequestTypes
from adrates_b200.dates import BusDayAdjustTypes, CalendarTypes, Date, DayCountTypes, FrequencyTypes
from adrates_b200.global_types import CurrencyTypes, CurveTypes, SwapTypes
from adrates_b200.trades import OIS
from tests.conftest import load_golden

ALL4 = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA, RequestTypes.CASHFLOWS]


def golden():
    return load_golden("ref_cashflows.json")


def make_cal_trade(spec, cv):
    dc = DayCountTypes[cv["dc"]]
    return OIS(effective_dt=Date(*spec["effective"]), term_dt_or_tenor=spec["tenor"], fixed_leg_type=SwapTypes[spec["side"]],
               fixed_coupon=spec["coupon"], fixed_freq_type=FrequencyTypes[spec["fixed_freq"]], fixed_dc_type=dc,
               floating_index=CurveTypes[cv["name"]], currency=CurrencyTypes[cv["name"][:3]], notional=spec["notional"],
               float_spread=spec["spread"], float_freq_type=FrequencyTypes[spec["float_freq"]], float_dc_type=dc,
               cal_type=CalendarTypes[spec["cal"]], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)


def assert_rows_match(cf, spec, tol=1e-12):
    """Every field of every row; amounts and PVs to `tol` of the notional, DFs to `tol` absolute."""
    rows = spec["rows"]
    assert len(cf) == len(rows)
    n = spec["notional"]
    for got, ref in zip(cf.cashflows, rows):
        assert [got.payment_date.d(), got.payment_date.m(), got.payment_date.y()] == ref["payment_date"]
        assert got.leg_type == ref["leg_type"]
        assert got.notional == ref["notional"]
        assert got.accrual_period == ref["accrual_period"]                       # day-count fractions are bit-exact
        assert abs(got.payment_fraction - ref["payment_fraction"]) <= tol
        assert abs(got.amount - ref["amount"]) <= tol * n
        assert abs(got.discount_factor - ref["discount_factor"]) <= tol
        assert abs(got.discounted_amount - ref["discounted_amount"]) <= tol * n
    assert abs(cf.total_amount - spec["total_amount"]) <= 10 * tol * n
    assert abs(cf.total_pv - spec["total_pv"]) <= 10 * tol * n
    assert abs(cf.sum().amount - spec["total_pv"]) <= 10 * tol * n
    assert len(cf.fixed()) + len(cf.floating()) == len(cf) and len(cf.pay()) + len(cf.receive()) == len(cf)
    assert len(cf.notional_exchange()) == 0
    assert np.isclose(cf.fixed().total_pv + cf.floating().total_pv, cf.total_pv)


def credit_golden():
    return load_golden("ref_cashflows_credit.json")


def assert_credit_rows_match(cf, rec, scale, tol=1e-12):
    rows = rec["rows"]
    assert [c.leg_type for c in cf.cashflows] == [r["leg_type"] for r in rows]
    for got, ref in zip(cf.cashflows, rows):
        assert [got.payment_date.d(), got.payment_date.m(), got.payment_date.y()] == ref["payment_date"]
        assert abs(got.notional - ref["notional"]) <= tol * scale
        assert got.accrual_period == ref["accrual_period"]
        assert abs(got.payment_fraction - ref["payment_fraction"]) <= tol
        assert abs(got.amount - ref["amount"]) <= tol * scale
        assert abs(got.discount_factor - ref["discount_factor"]) <= tol
        assert abs(got.discounted_amount - ref["discounted_amount"]) <= tol * scale
    assert abs(cf.total_amount - rec["total_amount"]) <= 10 * tol * scale
    assert abs(cf.total_pv - rec["total_pv"]) <= 10 * tol * scale
    assert repr(cf) == rec["repr"]
