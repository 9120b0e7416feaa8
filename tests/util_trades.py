"""Shared helpers: rebuild golden-vector trades/curves with the host layer and extract the
leg arrays the reference engine extracts (engine.py:2519-2527, 2858-2877)."""
import numpy as np

from adrates_b200.dates import (Date, BusDayAdjustTypes, DayCountTypes, FrequencyTypes, times_from_dates)
from adrates_b200.global_types import SwapTypes, CurveTypes, CurrencyTypes, InterpTypes
from adrates_b200.trades import OIS

METHOD = {"LINEAR_ZERO_RATES": 4, "FLAT_FWD_RATES": 1}


def make_calibration_swaps(cv):
    vd = Date(*cv["value_dt"])
    dc = DayCountTypes[cv["dc"]]
    fq = FrequencyTypes[cv["freq"]]
    return vd, [OIS(effective_dt=vd, term_dt_or_tenor=t, fixed_leg_type=SwapTypes.PAY, fixed_coupon=px / 100,
                    fixed_freq_type=fq, fixed_dc_type=dc, floating_index=CurveTypes[cv["name"]],
                    currency=CurrencyTypes[cv["name"][:3]], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                    float_freq_type=fq, float_dc_type=dc) for t, px in zip(cv["tenors"], cv["px"])]


def make_trade(spec, cv):
    dc = DayCountTypes[cv["dc"]]
    return OIS(effective_dt=Date(*spec["effective"]), term_dt_or_tenor=spec["tenor"],
               fixed_leg_type=SwapTypes[spec["side"]], fixed_coupon=spec["coupon"],
               fixed_freq_type=FrequencyTypes[spec["fixed_freq"]], fixed_dc_type=dc,
               floating_index=CurveTypes[cv["name"]], currency=CurrencyTypes[cv["name"][:3]],
               notional=spec["notional"], payment_lag=spec["payment_lag"], float_spread=spec["spread"],
               float_freq_type=FrequencyTypes[spec["float_freq"]], float_dc_type=dc,
               bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)


def leg_arrays(swap, value_dt):
    """(fixed, floating) dicts in the shape oracle.ois_analytics takes."""
    fl, ft = swap._fixed_leg, swap._float_leg
    fixed = dict(
        payment_times=np.array([times_from_dates(d, value_dt, fl._dc_type) for d in fl._payment_dts]),
        payments=np.array(fl._payments), principal=fl._principal,
        leg_sign=+1.0 if fl._leg_type == SwapTypes.RECEIVE else -1.0,
        value_time=times_from_dates(value_dt, value_dt, fl._dc_type))
    n = len(ft._payment_dts)
    floating = dict(
        payment_times=np.array([times_from_dates(d, value_dt, ft._dc_type) for d in ft._payment_dts]),
        start_times=np.array([times_from_dates(d, value_dt, ft._dc_type) for d in ft._start_accrued_dts]),
        end_times=np.array([times_from_dates(d, value_dt, ft._dc_type) for d in ft._end_accrued_dts]),
        pay_alphas=np.array(ft._year_fracs), spreads=np.full(n, ft._spread),
        notionals=np.array(ft._notional_array or [ft._notional] * n), principal=ft._principal,
        leg_sign=+1.0 if ft._leg_type == SwapTypes.RECEIVE else -1.0,
        value_time=times_from_dates(value_dt, value_dt, ft._dc_type))
    return fixed, floating


def rel_err(got, ref, scale):
    """SURVEY R3 parity metric: |x - ref| <= tol * max(|ref|, scale)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale)))


def trade_scales(spec):
    """Natural magnitudes for the parity metric: PV ~ notional, delta ~ notional*1e-4*T,
    gamma ~ notional*1e-8*T^2 (T in years, at least 1)."""
    n = spec["notional"]
    ten = spec["tenor"].upper()
    T = float(ten[:-1]) * {"D": 1 / 365, "W": 7 / 365, "M": 1 / 12, "Y": 1.0}[ten[-1]]
    T = max(T, 1.0)
    return n, n * 1e-4 * T, n * 1e-8 * T * T


def build_model(cv):
    """Model with the golden curve built through the reference-facing API."""
    from adrates_b200.models import Model
    m = Model(Date(*cv["value_dt"]))
    m.build_curve(name=cv["name"], px_list=cv["px"], tenor_list=cv["tenors"], spot_days=0, swap_type=SwapTypes.PAY,
                  fixed_dcc_type=DayCountTypes[cv["dc"]], fixed_freq_type=FrequencyTypes[cv["freq"]],
                  float_freq_type=FrequencyTypes[cv["freq"]], float_dc_type=DayCountTypes[cv["dc"]],
                  bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes[cv["interp"]])
    return m
