"""Zero-coupon inflation swaps: host objects mirroring the reference and the batched PV on the GPU.

Reference: cavour/market/indices/inflation_index.py (InflationIndex), cavour/market/curves/inflation_curve.py
(InflationCurve), cavour/trades/rates/swap_inflation_leg.py (SwapInflationLeg), cavour/trades/rates/zcis.py
(ZeroCouponInflationSwap).  The reference offers a non-AD `value()` for these trades on the path-A discount
curve and no Engine route / Greeks; the same surface is kept here.

A ZCIS is two cashflows on one payment date: notional x [(1 + r)^T - 1] against notional x [I_final/I_base - 1].
The index arithmetic (lag, fixings, intra-month interpolation, seasonality, curve projection) is host logic on a
handful of dates per trade; the discounting of the whole book, PV_i = sum_j amt_ij * DF(t_ij) / DF(0) with DF on
the path-A nodes (Interpolator._uinterpolate), runs on the device (`cav_cashflow_pv`).  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .curves import DiscountCurve
from .dates import (BusDayAdjustTypes, Calendar, CalendarTypes, Date, DateGenRuleTypes, DayCount, DayCountTypes, FrequencyTypes,
                    times_from_dates)
from .argcheck import check_argument_types
from .error import LibError
from .global_types import (CurrencyTypes, InflationIndexTypes, InflationInterpTypes, InstrumentTypes, InterpTypes, ONE_MILLION,
                           SwapTypes)


class InflationIndex:
    """Monthly CPI index with publication lag, daily interpolation and optional seasonality
    (inflation_index.py:71-466)."""

    def __init__(self, index_type: InflationIndexTypes, base_date: Date, base_index: float, currency: CurrencyTypes,
                 lag_months: int = 3, interp_type: InflationInterpTypes = InflationInterpTypes.LINEAR,
                 seasonality_factors: Optional[Dict[int, float]] = None):
        check_argument_types(self.__init__, locals())
        if base_index <= 0.0:
            raise LibError("Base index must be positive")
        if lag_months < 0:
            raise LibError("Lag months must be non-negative")
        if seasonality_factors is not None:
            self._validate_seasonality_factors(seasonality_factors)
        self._index_type = index_type
        self._base_date = base_date
        self._base_index = base_index
        self._currency = currency
        self._lag_months = lag_months
        self._interp_type = interp_type
        self._seasonality_factors = seasonality_factors or {}
        self._use_seasonality = len(self._seasonality_factors) > 0
        self._fixings: Dict[int, tuple] = {base_date.serial(): (base_date, base_index)}
        self._inflation_curve = None

    @staticmethod
    def _validate_seasonality_factors(factors: Dict[int, float]):
        if set(factors.keys()) != set(range(1, 13)):
            raise LibError(f"Seasonality factors must include all months 1-12. Got: {sorted(factors.keys())}")
        for month, factor in factors.items():
            if factor <= 0:
                raise LibError(f"Seasonality factors must be positive. Month {month} has factor {factor}")
        avg = sum(factors.values()) / 12.0
        if abs(avg - 1.0) > 0.01:
            raise LibError(f"Seasonality factors should average to 1.0 (within 1% tolerance). Got average: {avg:.6f}")

    def _apply_seasonality(self, date: Date, cpi_value: float) -> float:
        if not self._use_seasonality:
            return cpi_value
        return cpi_value * self._seasonality_factors.get(date.m(), 1.0)

    def add_fixing(self, fixing_date: Date, index_value: float):
        if index_value <= 0.0:
            raise LibError(f"Index value must be positive, got {index_value}")
        self._fixings[fixing_date.serial()] = (fixing_date, index_value)

    def set_inflation_curve(self, inflation_curve):
        self._inflation_curve = inflation_curve

    def _apply_lag(self, ref_date: Date) -> Date:
        return ref_date.add_months(-self._lag_months)

    def get_index(self, ref_date: Date, apply_lag: bool = True) -> float:
        """inflation_index.py:240-287: fixings first (interpolated), then the curve projection."""
        lookup = self._apply_lag(ref_date) if apply_lag else ref_date
        value = self._get_historical_index(lookup)
        if value is not None:
            return self._apply_seasonality(lookup, value)
        if self._inflation_curve is not None:
            return self._apply_seasonality(lookup, self._inflation_curve.forward_index(lookup))
        raise LibError(f"No fixing available for {lookup} and no inflation curve set. "
                       f"Add fixings via add_fixing() or set curve via set_inflation_curve().")

    def inflation_ratio(self, start_dt: Date, end_dt: Date, apply_lag: bool = True) -> float:
        i0 = self.get_index(start_dt, apply_lag=apply_lag)
        i1 = self.get_index(end_dt, apply_lag=apply_lag)
        if i0 <= 0.0:
            raise LibError(f"Start index must be positive, got {i0}")
        return i1 / i0

    def _get_historical_index(self, lookup_date: Date) -> Optional[float]:
        lookup = lookup_date
        if not self._fixings:
            return None
        keys = sorted(self._fixings.keys())
        if lookup.serial() < keys[0] or lookup.serial() > keys[-1]:
            return None
        if lookup.serial() in self._fixings:
            return self._fixings[lookup.serial()][1]
        hi = int(np.searchsorted(keys, lookup.serial(), side="left"))     # first fixing after the date
        lo_d, lo_v = self._fixings[keys[hi - 1]]
        hi_d, hi_v = self._fixings[keys[hi]]
        return self._interpolate(lookup, lo_d, hi_d, lo_v, hi_v)

    def _interpolate(self, target_date: Date, lower_date: Date, upper_date: Date, lower_value: float, upper_value: float) -> float:
        target, lo_d, hi_d, lo_v, hi_v = target_date, lower_date, upper_date, lower_value, upper_value
        if self._interp_type == InflationInterpTypes.FLAT:
            return lo_v
        dc = DayCount(DayCountTypes.ACT_365F)
        total = dc.year_frac(lo_d, hi_d)[0]
        elapsed = dc.year_frac(lo_d, target)[0]
        if total == 0:
            return lo_v
        weight = elapsed / total
        if self._interp_type == InflationInterpTypes.LINEAR:
            return lo_v + weight * (hi_v - lo_v)
        if self._interp_type == InflationInterpTypes.COMPOUND:
            return lo_v * ((hi_v / lo_v) ** weight)
        raise LibError(f"Unknown interpolation type: {self._interp_type}")

    def get_all_fixings(self) -> list:
        return [self._fixings[k] for k in sorted(self._fixings)]


class InflationCurve(DiscountCurve):
    """Cumulative inflation factors I(T)/I(0) = (1 + r_k)^T_k at the ZCIS maturities, interpolated like a discount
    curve (inflation_curve.py:91-244, 353-424)."""

    def __init__(self, value_dt: Date, zcis_instruments: list, base_cpi: float, currency: CurrencyTypes,
                 index_type: InflationIndexTypes, discount_curve: DiscountCurve = None,
                 interp_type: InflationInterpTypes = InflationInterpTypes.LINEAR,
                 dc_type: DayCountTypes = DayCountTypes.ACT_365F, check_refit: bool = False):
        check_argument_types(self.__init__, locals())
        if base_cpi <= 0.0:
            raise LibError("Base CPI must be positive")
        if len(zcis_instruments) < 2:
            raise LibError("Need at least 2 ZCIS instruments to build a curve")
        self._value_dt = value_dt
        self._used_swaps = zcis_instruments
        self._base_cpi = base_cpi
        self._currency = currency
        self._index_type = index_type
        self._discount_curve = discount_curve
        self._interp_type_infl = interp_type
        self._dc_type = dc_type
        dc = DayCount(dc_type)
        self.swap_times, self.tenors, rates = [], [], []
        for z in zcis_instruments:
            rates.append(z._fixed_rate)
            yf = dc.year_frac(z._effective_dt, z._maturity_dt)[0]
            self.swap_times.append(yf)
            self.tenors.append(f"{int(round(yf))}Y" if abs(yf - round(yf)) < 0.1 else f"{yf:.2f}Y")
        self._interp_type = {InflationInterpTypes.LINEAR: InterpTypes.LINEAR_ZERO_RATES,
                             InflationInterpTypes.COMPOUND: InterpTypes.LINEAR_ZERO_RATES,
                             InflationInterpTypes.FLAT: InterpTypes.FLAT_FWD_RATES}.get(interp_type,
                                                                                       InterpTypes.LINEAR_ZERO_RATES)
        self._times = np.array([0.0] + list(self.swap_times))
        self._dfs = np.array([1.0] + [(1.0 + r) ** t for t, r in zip(self.swap_times, rates)])
        if not all(self._times[i] < self._times[i + 1] for i in range(len(self._times) - 1)):
            raise LibError("Pillar times must be strictly increasing")
        self._check_refit = check_refit
        if check_refit:
            self._check_refits(1e-10)

    def _df(self, t: float) -> float:
        return self._node_df(float(t))

    def _check_refits(self, zcis_tol):
        dc = DayCount(self._dc_type)
        for z in self._used_swaps:
            yf = dc.year_frac(z._effective_dt, z._maturity_dt)[0]
            implied = (self._df(yf) ** (1.0 / yf)) - 1.0 if yf > 0 else 0.0
            if abs(implied - z._fixed_rate) > zcis_tol:
                raise LibError(f"ZCIS with maturity {z._maturity_dt} not repriced. "
                               f"Difference is {abs(implied - z._fixed_rate) * 10000:.4f} bps")

    def forward_index(self, target_date: Date) -> float:
        if target_date < self._value_dt:
            raise LibError(f"Cannot project CPI before value date. Target: {target_date}, Value: {self._value_dt}")
        yf = DayCount(self._dc_type).year_frac(self._value_dt, target_date)[0]
        return self._base_cpi * self._df(yf)

    def inflation_rate(self, start_dt: Date, end_dt: Date) -> float:
        if end_dt <= start_dt:
            raise LibError("End date must be after start date")
        yf = DayCount(self._dc_type).year_frac(start_dt, end_dt)[0]
        if yf <= 0:
            raise LibError("Year fraction must be positive")
        return ((self.forward_index(end_dt) / self.forward_index(start_dt)) ** (1.0 / yf)) - 1.0


def _maturity_and_payment(effective_dt, end, payment_lag, cal_type, bd_type):
    termination = end if isinstance(end, Date) else effective_dt.add_tenor(end)
    cal = Calendar(cal_type)
    maturity = cal.adjust(termination, bd_type)
    if effective_dt > maturity:
        raise LibError("Start date after maturity date")
    payment = maturity if payment_lag == 0 else cal.add_business_days(maturity, payment_lag)
    return termination, maturity, payment


class SwapInflationLeg:
    """Single inflation-linked payment notional x [I(T - lag)/I(base - lag) - 1] (swap_inflation_leg.py:89-236)."""

    def __init__(self, effective_dt: Date, end_dt: (Date, str), leg_type: SwapTypes, inflation_index: InflationIndex,
                 notional: float = ONE_MILLION, payment_lag: int = 0, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING):
        check_argument_types(self.__init__, locals())
        self.instrument_type = InstrumentTypes.SWAP_INFLATION_LEG
        self._termination_dt, self._maturity_dt, self._payment_dt = _maturity_and_payment(
            effective_dt, end_dt, payment_lag, cal_type, bd_type)
        self._effective_dt = effective_dt
        self._leg_type = leg_type
        self._inflation_index = inflation_index
        self._notional = notional
        self._payment_lag = payment_lag
        self._cal_type, self._bd_type = cal_type, bd_type
        self._base_cpi_ref_dt = effective_dt
        self._final_cpi_ref_dt = self._maturity_dt
        self._base_index = self._final_index = self._inflation_return = self._payment_amount = None
        self._payment_df = self._payment_pv = None

    def payment(self, inflation_curve=None) -> float:
        """Signed undiscounted payment (host index arithmetic)."""
        if inflation_curve is not None:
            self._inflation_index.set_inflation_curve(inflation_curve)
        self._base_index = self._inflation_index.get_index(self._base_cpi_ref_dt, apply_lag=True)
        self._final_index = self._inflation_index.get_index(self._final_cpi_ref_dt, apply_lag=True)
        if self._base_index <= 0.0:
            raise LibError(f"Base index must be positive, got {self._base_index}")
        self._inflation_return = (self._final_index / self._base_index) - 1.0
        self._payment_amount = self._notional * self._inflation_return
        return self._payment_amount if self._leg_type == SwapTypes.RECEIVE else -self._payment_amount

    def value(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        amt = self.payment(inflation_curve)
        pv = cashflow_pv(discount_curve, value_dt, [[(self._payment_dt, amt)]])[0]
        self._payment_pv = abs(pv) if self._payment_dt > value_dt else 0.0
        return float(pv)


class ZeroCouponInflationSwap:
    """Fixed compounded return against cumulative inflation, one payment at maturity (zcis.py:79-238)."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), fixed_leg_type: SwapTypes, fixed_rate: float,
                 inflation_index: InflationIndex, notional: float = ONE_MILLION, payment_lag: int = 0,
                 dc_type: DayCountTypes = DayCountTypes.ACT_365F, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING):
        check_argument_types(self.__init__, locals())
        self.instrument_type = InstrumentTypes.ZCIS
        self.derivative_type = InstrumentTypes.ZCIS
        self._termination_dt, self._maturity_dt, self._payment_dt = _maturity_and_payment(
            effective_dt, term_dt_or_tenor, payment_lag, cal_type, bd_type)
        self._effective_dt = effective_dt
        self._fixed_leg_type = fixed_leg_type
        self._fixed_rate = fixed_rate
        self._inflation_index = inflation_index
        self._notional = notional
        self._payment_lag = payment_lag
        self._dc_type = dc_type
        self._cal_type = cal_type
        self._bd_type = bd_type
        infl_type = SwapTypes.RECEIVE if fixed_leg_type == SwapTypes.PAY else SwapTypes.PAY
        self._inflation_leg = SwapInflationLeg(effective_dt, self._termination_dt, infl_type, inflation_index, notional,
                                               payment_lag, cal_type, bd_type)
        self._fixed_return = self._fixed_payment = self._fixed_pv = self._inflation_pv = self._payment_df = None

    def cashflows(self, inflation_curve=None):
        """[(payment date, signed amount)] of the fixed and the inflation leg (zcis.py:203-232)."""
        yf = DayCount(self._dc_type).year_frac(self._effective_dt, self._maturity_dt)[0]
        self._fixed_return = ((1.0 + self._fixed_rate) ** yf) - 1.0
        self._fixed_payment = self._notional * self._fixed_return
        fixed = -self._fixed_payment if self._fixed_leg_type == SwapTypes.PAY else self._fixed_payment
        return [(self._payment_dt, fixed), (self._inflation_leg._payment_dt, self._inflation_leg.payment(inflation_curve))]

    def value(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        return float(value_zcis_book([self], value_dt, discount_curve, inflation_curve)[0])

    def pv01(self, value_dt: Date, discount_curve: DiscountCurve) -> float:
        """|dPV / d fixed rate| x 1 bp of the compounded fixed payment (zcis.py:284-317)."""
        yf = DayCount(self._dc_type).year_frac(self._effective_dt, self._maturity_dt)[0]
        df = 0.0
        if self._payment_dt > value_dt:
            df = discount_curve.df(self._payment_dt, DayCountTypes.ACT_365F) / discount_curve.df(value_dt, DayCountTypes.ACT_365F)
        return abs(self._notional * yf * ((1.0 + self._fixed_rate) ** (yf - 1.0)) * df) * 0.0001

    def breakeven_inflation_rate(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        """zcis.py:242-290: annual rate whose compounded return equals the projected inflation return."""
        self._inflation_leg.payment(inflation_curve)
        ret = self._inflation_leg._inflation_return
        yf = DayCount(self._dc_type).year_frac(self._effective_dt, self._maturity_dt)[0]
        if yf <= 0:
            raise LibError("Year fraction must be positive")
        return ((1.0 + ret) ** (1.0 / yf)) - 1.0


class SwapYoYInflationLeg:
    """Periodic payments notional x alpha_i x [I(end_i)/I(end_i - 12M) - 1 + spread]
    (cavour/trades/rates/swap_yoy_inflation_leg.py:95-262).  Schedule only; valuation and Greeks go through
    Position.compute -> yoy_engine (Engine._compute_yoy_iis)."""

    def __init__(self, effective_dt: Date, end_dt: (Date, str), leg_type: SwapTypes, inflation_index: InflationIndex,
                 freq_type: FrequencyTypes, dc_type: DayCountTypes, notional: float = ONE_MILLION, spread: float = 0.0,
                 payment_lag: int = 0, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self.instrument_type = InstrumentTypes.SWAP_YOY_INFLATION_LEG
        self._termination_dt = end_dt if isinstance(end_dt, Date) else effective_dt.add_tenor(end_dt)
        cal = Calendar(cal_type)
        self._maturity_dt = cal.adjust(self._termination_dt, bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError("Start date after maturity date")
        self._effective_dt, self._end_dt, self._leg_type = effective_dt, end_dt, leg_type
        self._inflation_index, self._freq_type, self._dc_type = inflation_index, freq_type, dc_type
        self._notional, self._spread, self._payment_lag = notional, spread, payment_lag
        self._cal_type, self._bd_type = cal_type, bd_type
        self._dg_type = dg_type or DateGenRuleTypes.BACKWARD
        self._end_of_month = end_of_month
        self.generate_payment_schedule()

    def generate_payment_schedule(self):
        """Accrual periods, payment dates and the two CPI reference dates of every coupon (swap_yoy_inflation_leg.py:193-262)"""
        from .dates import Schedule
        effective_dt, payment_lag, cal = self._effective_dt, self._payment_lag, Calendar(self._cal_type)
        dts = Schedule(effective_dt, self._termination_dt, self._freq_type, self._cal_type, self._bd_type, self._dg_type,
                       end_of_month=self._end_of_month)._adjusted_dts
        if len(dts) < 2:
            raise LibError("Schedule has none or only one date")
        dc = DayCount(self._dc_type)
        self._start_accrued_dts, self._end_accrued_dts, self._payment_dts = [], [], []
        self._year_fracs, self._accrued_days, self._yoy_start_dts, self._yoy_end_dts = [], [], [], []
        for start, end in zip(dts[:-1], dts[1:]):
            alpha, days, _ = dc.year_frac(start, end)
            self._start_accrued_dts.append(start)
            self._end_accrued_dts.append(end)
            self._payment_dts.append(end if payment_lag == 0 else cal.add_business_days(end, payment_lag))
            self._year_fracs.append(alpha)
            self._accrued_days.append(days)
            self._yoy_end_dts.append(end)                     # CPI reference dates: period end and one year before
            self._yoy_start_dts.append(end.add_months(-12))
        self._start_cpis, self._end_cpis, self._yoy_rates, self._payments, self._dfs, self._pvs = [], [], [], [], [], []

    def value(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        """Non-AD PV of the leg (swap_yoy_inflation_leg.py:267-366): per future payment the lagged CPI at both reference dates
        (fixings first, then the curve - a reference date before the value date without a fixing raises), the year-on-year
        rate plus spread on the accrual, discounted relative to the value date.  Host code as in the reference; Greeks and
        batched valuation go through Position.compute."""
        index = self._inflation_index
        if inflation_curve is not None:
            index.set_inflation_curve(inflation_curve)
        n = len(self._payment_dts)
        cols = [[0.0] * n for _ in range(6)]
        self._start_cpis, self._end_cpis, self._yoy_rates, self._payments, self._dfs, self._pvs = cols
        total = 0.0
        for i, pay in enumerate(self._payment_dts):
            if pay <= value_dt:
                continue
            start = index.get_index(self._yoy_start_dts[i], apply_lag=True)
            end = index.get_index(self._yoy_end_dts[i], apply_lag=True)
            if start <= 0.0:
                raise LibError(f"Start CPI must be positive, got {start}")
            rate = (end / start) - 1.0
            amount = self._notional * self._year_fracs[i] * (rate + self._spread)
            df = discount_curve.df(pay, self._dc_type) / discount_curve.df(value_dt, self._dc_type)
            for col, v in zip(cols, (start, end, rate, amount, df, amount * df)):
                col[i] = v
            total += amount * df
        return -total if self._leg_type == SwapTypes.PAY else total


class YoYInflationSwap:
    """Fixed coupons against year-on-year inflation coupons on one schedule
    (cavour/trades/rates/yoy_inflation_swap.py:85-220)."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), fixed_leg_type: SwapTypes, fixed_rate: float,
                 inflation_index: InflationIndex, freq_type: FrequencyTypes, notional: float = ONE_MILLION,
                 inflation_spread: float = 0.0, dc_type: DayCountTypes = DayCountTypes.ACT_365F, payment_lag: int = 0,
                 cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        from .global_types import CurveTypes
        from .trades import SwapFixedLeg
        dg_type = dg_type or DateGenRuleTypes.BACKWARD
        self.instrument_type = InstrumentTypes.YOY_INFLATION_SWAP
        self.derivative_type = InstrumentTypes.YOY_INFLATION_SWAP
        self._termination_dt = term_dt_or_tenor if isinstance(term_dt_or_tenor, Date) else effective_dt.add_tenor(term_dt_or_tenor)
        self._maturity_dt = Calendar(cal_type).adjust(self._termination_dt, bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError("Start date after maturity date")
        self._effective_dt, self._fixed_leg_type, self._fixed_rate = effective_dt, fixed_leg_type, fixed_rate
        self._inflation_index, self._freq_type, self._notional = inflation_index, freq_type, notional
        self._inflation_spread, self._dc_type, self._payment_lag = inflation_spread, dc_type, payment_lag
        self._cal_type, self._bd_type, self._dg_type, self._end_of_month = cal_type, bd_type, dg_type, end_of_month
        infl_type = SwapTypes.RECEIVE if fixed_leg_type == SwapTypes.PAY else SwapTypes.PAY
        currency = inflation_index._currency
        floating_index = {CurrencyTypes.GBP: CurveTypes.GBP_OIS_SONIA, CurrencyTypes.USD: CurveTypes.USD_OIS_SOFR,
                          CurrencyTypes.EUR: CurveTypes.EUR_OIS_ESTR}.get(currency, CurveTypes.USD_OIS_SOFR)
        self._fixed_leg = SwapFixedLeg(effective_dt, self._termination_dt, fixed_leg_type, fixed_rate, freq_type, dc_type,
                                       floating_index, currency, notional, 0.0, payment_lag, cal_type, bd_type, dg_type,
                                       end_of_month)
        self._inflation_leg = SwapYoYInflationLeg(effective_dt, self._termination_dt, infl_type, inflation_index,
                                                  freq_type, dc_type, notional, inflation_spread, payment_lag, cal_type,
                                                  bd_type, dg_type, end_of_month)
        self._fixed_pv = self._inflation_pv = None

    def position(self, model):
        """Convenience the reference lacks (its YoYInflationSwap has no .position); Position(swap, model) works too."""
        from .position import Position
        return Position(self, model)

    def value(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        """Non-AD net PV, fixed leg then inflation leg (yoy_inflation_swap.py:224-260)."""
        self._fixed_pv = self._fixed_leg.value(value_dt, discount_curve)
        self._inflation_pv = self._inflation_leg.value(value_dt, discount_curve, inflation_curve)
        return self._fixed_pv + self._inflation_pv

    def _annuity(self, value_dt: Date, discount_curve: DiscountCurve) -> float:
        df0 = discount_curve.df(value_dt, DayCountTypes.ACT_365F)
        return sum(yf * discount_curve.df(dt, DayCountTypes.ACT_365F) / df0
                   for dt, yf in zip(self._fixed_leg._payment_dts, self._fixed_leg._year_fracs) if dt > value_dt)

    def breakeven_rate(self, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None) -> float:
        """Fixed rate at which the swap is worth zero: inflation-leg PV over notional x annuity (yoy_inflation_swap.py:264-336)."""
        inflation_pv = self._inflation_leg.value(value_dt, discount_curve, inflation_curve)
        annuity = self._annuity(value_dt, discount_curve)
        if annuity <= 0:
            raise LibError("Annuity must be positive for breakeven calculation")
        sign = 1.0 if self._fixed_leg_type == SwapTypes.PAY else -1.0
        return sign * inflation_pv / (self._notional * annuity)

    def pv01(self, value_dt: Date, discount_curve: DiscountCurve) -> float:
        return abs(self._notional * self._annuity(value_dt, discount_curve) * 0.0001)


def cashflow_pv(discount_curve: DiscountCurve, value_dt: Date, trades, device: int = 0) -> np.ndarray:
    """PV per trade of explicit cashflows [(date, signed amount), ...] on the path-A discount curve:
    sum amt * DF(date)/DF(value_dt), cashflows on or before the value date are worth 0
    (zcis.py:210-218, swap_inflation_leg.py:211-221).  Evaluated by the CUDA library (cav_cashflow_pv)."""
    from . import _native
    offsets = np.zeros(len(trades) + 1, dtype=np.int64)
    t, amt = [], []
    for i, cfs in enumerate(trades):
        for dt, a in cfs:
            if dt > value_dt:
                t.append(times_from_dates(dt, discount_curve._value_dt, DayCountTypes.ACT_365F))
                amt.append(a)
        offsets[i + 1] = len(t)
    t0 = times_from_dates(value_dt, discount_curve._value_dt, DayCountTypes.ACT_365F)
    if t0 < 0.0 or (len(t) and min(t) < 0.0):
        raise LibError("Interpolate times must all be >= 0")
    pv, _ = _native.lib(device).cashflow_pv(discount_curve._interp_type.value, discount_curve._times,
                                            discount_curve._dfs, t0, offsets, np.array(t), np.array(amt))
    return pv


def value_zcis_book(swaps, value_dt: Date, discount_curve: DiscountCurve, inflation_curve=None, device: int = 0):
    """PV of every ZCIS of a book in one device call (replaces a Python loop over `zcis.value`)."""
    return cashflow_pv(discount_curve, value_dt, [s.cashflows(inflation_curve) for s in swaps], device)
