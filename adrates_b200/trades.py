"""Trade objects that feed the flattener: OIS and its two legs.

Interface mirror of the reference (constructor argument order, attribute names that the
engine reads, error behaviour):
    cavour/trades/rates/ois.py:100-205            OIS, OIS.position
    cavour/trades/rates/swap_fixed_leg.py:63-196  SwapFixedLeg.generate_payments
    cavour/trades/rates/swap_float_leg.py:65-186  SwapFloatLeg.generate_payment_dts
The legs only generate schedules (host integer date math, once per trade).  Valuation and
Greeks go through Position.compute -> the CUDA path; the non-AD `value()` helpers below
are the reference's path-A (curve.df) arithmetic kept for API parity.
"""
from __future__ import annotations

from enum import Enum

import numpy as np

from .dates import (Date, Calendar, CalendarTypes, BusDayAdjustTypes, DateGenRuleTypes, DayCount,
                    DayCountTypes, FrequencyTypes, Schedule)
from .argcheck import check_argument_types
from .error import LibError
from .global_types import SwapTypes, InstrumentTypes, CurveTypes, CurrencyTypes, ONE_MILLION


def _resolve_end(effective_dt: Date, end):
    return end if isinstance(end, Date) else effective_dt.add_tenor(end)


class _SwapLeg:
    """Schedule-bearing leg: accrual start/end, payment dates and accrual fractions."""

    def _init_common(self, effective_dt, end_dt, leg_type, freq_type, dc_type, floating_index, currency,
                     notional, payment_lag, cal_type, bd_type, dg_type, end_of_month, what):
        if not isinstance(effective_dt, Date):
            raise LibError("effective_dt must be a Date")
        self._termination_dt = _resolve_end(effective_dt, end_dt)
        self._maturity_dt = Calendar(cal_type).adjust(self._termination_dt, bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError(what + " date after maturity date")
        self._effective_dt = effective_dt
        self._end_dt = end_dt
        self._leg_type = leg_type
        self._freq_type = freq_type
        self._payment_lag = payment_lag
        self._notional = notional
        self._floating_index = floating_index
        self._currency = currency
        self._dc_type = dc_type
        self._cal_type = cal_type
        self._bd_type = bd_type
        self._dg_type = dg_type
        self._end_of_month = end_of_month

    # Rolled schedules are a pure function of these arguments and Dates are immutable: books hold many legs on the
    # same dates (both legs of a vanilla OIS, every trade of one maturity bucket), so the roll is memoised.
    _ROLL_CACHE: dict = {}
    _ROLL_CACHE_MAX = 200_000

    def _roll_schedule(self):
        key = (self._effective_dt._n, self._termination_dt._n, self._freq_type, self._cal_type, self._bd_type,
               self._dg_type, self._end_of_month, self._dc_type, self._payment_lag)
        hit = _SwapLeg._ROLL_CACHE.get(key)
        if hit is None:
            dts = Schedule(self._effective_dt, self._termination_dt, self._freq_type, self._cal_type,
                           self._bd_type, self._dg_type, end_of_month=self._end_of_month)._adjusted_dts
            if len(dts) < 2:
                raise LibError("Schedule has none or only one date")
            dc = DayCount(self._dc_type)
            cal = Calendar(self._cal_type)
            starts, ends, pays, pays_ad, fracs, days = [], [], [], [], [], []
            for start, end in zip(dts[:-1], dts[1:]):
                starts.append(start)
                ends.append(end)
                pays.append(end if self._payment_lag == 0 else cal.add_business_days(end, self._payment_lag))
                pays_ad.append(dc.year_frac(self._effective_dt, end)[0])
                alpha, num, _ = dc.year_frac(start, end)
                fracs.append(alpha)
                days.append(num)
            hit = (starts, ends, pays, pays_ad, fracs, days)
            if len(_SwapLeg._ROLL_CACHE) < _SwapLeg._ROLL_CACHE_MAX:
                _SwapLeg._ROLL_CACHE[key] = hit
        # per-leg lists (the reference's legs own theirs and some callers edit them)
        self._start_accrued_dts, self._end_accrued_dts = list(hit[0]), list(hit[1])
        self._payment_dts, self._payment_dts_ad = list(hit[2]), list(hit[3])
        self._year_fracs, self._accrued_days = list(hit[4]), list(hit[5])


class SwapFixedLeg(_SwapLeg):
    def __init__(self, effective_dt: Date, end_dt: (Date, str), leg_type: SwapTypes, coupon: float,
                 freq_type: FrequencyTypes, dc_type: DayCountTypes, floating_index: CurveTypes,
                 currency: CurrencyTypes, notional: float = ONE_MILLION, principal: float = 0.0,
                 payment_lag: int = 0, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self.intrument_type = InstrumentTypes.SWAP_FIXED_LEG
        self._init_common(effective_dt, end_dt, leg_type, freq_type, dc_type, floating_index, currency,
                          notional, payment_lag, cal_type, bd_type, dg_type, end_of_month, "Effective")
        self._principal = principal
        self._cpn = coupon
        self.generate_payments()

    def generate_payments(self):
        self._roll_schedule()
        self._adjusted_fixed_dts = list(self._payment_dts)
        self._rates = [self._cpn] * len(self._payment_dts)
        self._payments = [a * self._notional * self._cpn for a in self._year_fracs]

    def value(self, value_dt: Date, discount_curve):
        """Path-A (curve.df) PV of the fixed leg (swap_fixed_leg.py:200-245); keeps the per-payment discount factors, PVs and
        running PV of this valuation for print_valuation, as the reference does."""
        df0 = discount_curve.df(value_dt, self._dc_type)
        pv, df_p = 0.0, 0.0
        n = len(self._payment_dts)
        self._payment_dfs, self._payment_pvs, self._cumulative_pvs = [0.0] * n, [0.0] * n, [0.0] * n
        for i, (dt, amt) in enumerate(zip(self._payment_dts, self._payments)):
            if dt > value_dt:
                df_p = discount_curve.df(dt, self._dc_type) / df0
                pv += amt * df_p
                self._payment_dfs[i], self._payment_pvs[i], self._cumulative_pvs[i] = df_p, amt * df_p, pv
        if self._payment_dts[-1] > value_dt:
            pv += self._principal * df_p * self._notional
            self._payment_pvs[-1] += self._principal * df_p * self._notional
            self._cumulative_pvs[-1] = pv
        return -pv if self._leg_type == SwapTypes.PAY else pv

    def _header(self):
        print("START DATE:", self._effective_dt)
        print("MATURITY DATE:", self._maturity_dt)
        print("COUPON (%):", self._cpn * 100)
        print("FREQUENCY:", str(self._freq_type))
        print("DAY COUNT:", str(self._dc_type))

    def print_payments(self):
        """Schedule table: accrual dates, days, year fraction, rate, payment (swap_fixed_leg.py:250-285)"""
        self._header()
        rows = [[i + 1, self._payment_dts[i], self._start_accrued_dts[i], self._end_accrued_dts[i], self._accrued_days[i],
                 round(self._year_fracs[i], 4), round(self._rates[i] * 100.0, 4), round(self._payments[i], 2)]
                for i in range(len(self._payment_dts))]
        _print_table("PAYMENTS SCHEDULE:", ["PAY_NUM", "PAY_dt", "ACCR_START", "ACCR_END", "DAYS", "YEARFRAC", "RATE", "PMNT"], rows)

    def print_valuation(self):
        """Table of the last valuation: payment, discount factor, PV, running PV (swap_fixed_leg.py:289-325)"""
        self._header()
        if not getattr(self, "_payment_dfs", None):
            print("Payments not calculated.")
            return
        rows = [[i + 1, self._payment_dts[i], round(self._notional, 0), round(self._rates[i] * 100.0, 4), round(self._payments[i], 2),
                 round(self._payment_dfs[i], 4), round(self._payment_pvs[i], 2), round(self._cumulative_pvs[i], 2)]
                for i in range(len(self._payment_dts))]
        _print_table("PAYMENTS VALUATION:", ["PAY_NUM", "PAY_dt", "NOTIONAL", "RATE", "PMNT", "DF", "PV", "CUM_PV"], rows)


class SwapFloatLeg(_SwapLeg):
    def __init__(self, effective_dt: Date, end_dt: (Date, str), leg_type: SwapTypes, spread: float,
                 freq_type: FrequencyTypes, dc_type: DayCountTypes, floating_index: CurveTypes,
                 currency: CurrencyTypes, notional: float = ONE_MILLION, principal: float = 0.0,
                 payment_lag: int = 0, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False,
                 notional_exchange: bool = False):
        check_argument_types(self.__init__, locals())
        self.intrument_type = InstrumentTypes.SWAP_FLOAT_LEG
        self._init_common(effective_dt, end_dt, leg_type, freq_type, dc_type, floating_index, currency,
                          notional, payment_lag, cal_type, bd_type, dg_type, end_of_month, "Start")
        self._principal = 0.0  # the reference ignores the argument (swap_float_leg.py:106)
        self._notional_array = []
        self._spread = spread
        self._notional_exchange = notional_exchange
        self._payments = []
        self.generate_payment_dts()

    def generate_payment_dts(self):
        self._roll_schedule()

    def value(self, value_dt: Date, discount_curve, index_curve=None, first_fixing_rate=None):
        """Path-A PV of the floating leg (swap_float_leg.py:190-352)."""
        if discount_curve is None:
            raise LibError("Discount curve is None")
        index_curve = index_curve or discount_curve
        df0 = discount_curve.df(value_dt, self._dc_type)
        idx_dc = DayCount(index_curve._dc_type)
        pv, df_p, first = 0.0, 0.0, True
        n = len(self._payment_dts)
        # rows of this valuation for print_valuation (the coupon rows only: the exchanges are not inserted into the lists)
        self._valuation_rows = [(0.0, 0.0, 0.0, 0.0)] * n
        for i, dt in enumerate(self._payment_dts):
            if not dt > value_dt:
                continue
            s, e = self._start_accrued_dts[i], self._end_accrued_dts[i]
            if first and first_fixing_rate is not None:
                fwd = first_fixing_rate
            else:
                fwd = (index_curve.df(s, self._dc_type) / index_curve.df(e, self._dc_type) - 1.0) \
                    / idx_dc.year_frac(s, e)[0]
            first = False
            df_p = discount_curve.df(dt, self._dc_type) / df0
            pv += (fwd + self._spread) * self._year_fracs[i] * self._notional * df_p
            amt = (fwd + self._spread) * self._year_fracs[i] * self._notional
            self._valuation_rows[i] = (fwd, amt, df_p, amt * df_p)
        if self._notional_exchange:
            # -N at the effective date, +N at maturity, when not in the past (swap_float_leg.py:284-347);
            # unlike the reference this does not insert the exchange into the leg's date lists
            if self._effective_dt >= value_dt:
                pv -= self._notional * discount_curve.df(self._effective_dt, self._dc_type) / df0
            if self._maturity_dt >= value_dt and len(self._payment_dts) > 0:
                pv += self._notional * discount_curve.df(self._maturity_dt, self._dc_type) / df0
        return -pv if self._leg_type == SwapTypes.PAY else pv


    def _header(self):
        print("START DATE:", self._effective_dt)
        print("MATURITY DATE:", self._maturity_dt)
        print("SPREAD (bp):", self._spread * 10000)
        print("FREQUENCY:", str(self._freq_type))
        print("DAY COUNT:", str(self._dc_type))

    def print_payments(self):
        """Schedule table (swap_float_leg.py:356-390)"""
        self._header()
        rows = [[i + 1, self._payment_dts[i], self._start_accrued_dts[i], self._end_accrued_dts[i], self._accrued_days[i],
                 round(self._year_fracs[i], 4)] for i in range(len(self._payment_dts))]
        _print_table("PAYMENTS SCHEDULE:", ["PAY_NUM", "PAY_dt", "ACCR_START", "ACCR_END", "DAYS", "YEARFRAC"], rows)

    def print_valuation(self):
        """Table of the last valuation: forward, payment, discount factor, PV, running PV (swap_float_leg.py:394-440)"""
        self._header()
        if not getattr(self, "_valuation_rows", None):
            print("Rates not calculated.")
            return
        rows, run = [], 0.0
        for i, (fwd, amt, df_p, pv) in enumerate(self._valuation_rows):
            run += pv
            rows.append([i + 1, self._payment_dts[i], round(self._notional, 0), round(fwd * 100.0, 4), round(amt, 2), round(df_p, 4),
                         round(pv, 2), round(run, 2)])
        _print_table("PAYMENTS VALUATION:", ["PAY_NUM", "PAY_dt", "NOTIONAL", "IBOR", "PMNT", "DF", "PV", "CUM_PV"], rows)


def _print_table(title: str, header: list, rows: list):
    from tabulate import tabulate
    print("\n" + title)
    print(tabulate([[str(c) if isinstance(c, Date) else c for c in r] for r in rows], headers=header))


def _print_two_legs(first_title: str, first_leg, second_title: str, second_leg, method: str):
    for k, (title, leg) in enumerate(((first_title, first_leg), (second_title, second_leg))):
        print(("\n" if k else "") + "=" * 80)
        print(title)
        print("=" * 80)
        getattr(leg, method)()


class FinCompoundingTypes(Enum):
    """Overnight compounding conventions the reference names next to OIS (ois.py:62-66); the legs compound as COMPOUNDED."""
    COMPOUNDED = 1
    OVERNIGHT_COMPOUNDED_ANNUAL_RATE = 2
    AVERAGED = 3
    AVERAGED_DAILY = 4


class OIS:
    """Overnight index swap: fixed leg against compounded overnight floating leg."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), fixed_leg_type: SwapTypes, fixed_coupon: float,
                 fixed_freq_type: FrequencyTypes, fixed_dc_type: DayCountTypes, floating_index: CurveTypes,
                 currency: CurrencyTypes, notional: float = ONE_MILLION, payment_lag: int = 0,
                 float_spread: float = 0.0, float_freq_type: FrequencyTypes = FrequencyTypes.ANNUAL,
                 float_dc_type: DayCountTypes = DayCountTypes.THIRTY_E_360,
                 cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD):
        check_argument_types(self.__init__, locals())
        if not isinstance(fixed_leg_type, SwapTypes):
            raise LibError("fixed_leg_type must be a SwapTypes")
        self.derivative_type = InstrumentTypes.OIS_SWAP
        self._termination_dt = _resolve_end(effective_dt, term_dt_or_tenor)
        self._maturity_dt = Calendar(cal_type).adjust(self._termination_dt, bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError("Start date after maturity date")
        self._effective_dt = effective_dt
        float_leg_type = SwapTypes.RECEIVE if fixed_leg_type == SwapTypes.PAY else SwapTypes.PAY
        self._floating_index = floating_index
        self._currency = currency
        self._fixed_leg = SwapFixedLeg(effective_dt, self._termination_dt, fixed_leg_type, fixed_coupon,
                                       fixed_freq_type, fixed_dc_type, floating_index, currency, notional, 0.0,
                                       payment_lag, cal_type, bd_type, dg_type, False)
        self._float_leg = SwapFloatLeg(effective_dt, self._termination_dt, float_leg_type, float_spread,
                                       float_freq_type, float_dc_type, floating_index, currency, notional, 0.0,
                                       payment_lag, cal_type, bd_type, dg_type, False, False)
        self._adjusted_fixed_dts = self._fixed_leg._adjusted_fixed_dts
        self._fixed_coupon = self._fixed_leg._cpn
        self._fixed_year_fracs = self._fixed_leg._year_fracs
        self._start_dt = self._fixed_leg._effective_dt
        self._notional = notional

    def position(self, model):
        from .position import Position
        return Position(self, model)

    # --- non-AD path-A helpers (ois.py:209-320) ---
    def value(self, value_dt: Date, ois_curve=None, discount_curve=None, xccy_discount_curve=None, spot_fx: float = None,
              collateral_type=None, first_fixing_rate=None):
        """Non-AD PV (ois.py:209-272): both legs discounted on `discount_curve` (default: the projection curve `ois_curve`);
        with a collateral type in another currency the legs are discounted on `xccy_discount_curve` and the PV is converted
        into the collateral currency as `value / spot_fx`."""
        from .global_types import collateral_to_currency
        if discount_curve is None and collateral_type is None:
            discount_curve = ois_curve
        foreign_collateral = False
        if collateral_type is not None:
            collateral_ccy = collateral_to_currency(collateral_type)
            foreign_collateral = collateral_ccy != self._currency
            if foreign_collateral:
                if xccy_discount_curve is None or spot_fx is None:
                    raise ValueError(f"xccy_discount_curve and spot_fx required for {self._currency.name} swap with "
                                     f"{collateral_ccy.name} collateral")
                discount_curve = xccy_discount_curve
            else:
                discount_curve = ois_curve
        value = self._fixed_leg.value(value_dt, discount_curve) + \
            self._float_leg.value(value_dt, discount_curve, ois_curve, first_fixing_rate)
        return value / spot_fx if (foreign_collateral and spot_fx is not None) else value

    def pv01(self, value_dt, discount_curve):
        pv = self._fixed_leg.value(value_dt, discount_curve)
        return np.abs(pv / self._fixed_leg._cpn / self._fixed_leg._notional * 100)   # a numpy float, so that swap_rate of a matured swap is nan, not an error - as in the reference

    def print_fixed_leg_pv(self):
        self._fixed_leg.print_valuation()

    def print_float_leg_pv(self):
        self._float_leg.print_valuation()

    def print_payments(self):
        self._fixed_leg.print_payments()
        self._float_leg.print_payments()

    def ir01(self, value_dt, discount_curve):
        """PV change per basis point from revaluing on the curve bumped by -/+ 10 bp (ois.py:289-300)."""
        down = self.value(value_dt, discount_curve.bump(-0.001))
        up = self.value(value_dt, discount_curve.bump(0.001))
        return (up - down) / 10 / 2

    def swap_rate(self, value_dt, ois_curve, first_fixing_rate=None):
        pv01 = self.pv01(value_dt, ois_curve)
        return self._float_leg.value(value_dt, ois_curve, ois_curve, first_fixing_rate) / pv01 / self._fixed_leg._notional


class XccyBasisSwap:
    """Cross-currency basis swap: receive domestic floating, pay foreign floating + basis spread,
    notionals exchanged at start and maturity (cavour/trades/rates/xccy_basis_swap.py:67-205)."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), domestic_notional: float, foreign_notional: float,
                 domestic_spread: float, foreign_spread: float, domestic_freq_type: FrequencyTypes,
                 foreign_freq_type: FrequencyTypes, domestic_dc_type: DayCountTypes, foreign_dc_type: DayCountTypes,
                 domestic_floating_index: CurveTypes, foreign_floating_index: CurveTypes,
                 domestic_currency: CurrencyTypes, foreign_currency: CurrencyTypes,
                 domestic_payment_lag: int = 0, foreign_payment_lag: int = 0,
                 domestic_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 foreign_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 domestic_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 foreign_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 domestic_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 foreign_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 domestic_end_of_month: bool = False, foreign_end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self.derivative_type = InstrumentTypes.XCCY_SWAP
        self._termination_dt = _resolve_end(effective_dt, term_dt_or_tenor)
        self._maturity_dt = Calendar(domestic_cal_type).adjust(self._termination_dt, domestic_bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError("Start date after maturity date")
        self._effective_dt = effective_dt
        self._domestic_notional = domestic_notional
        self._foreign_notional = foreign_notional
        self._domestic_currency = domestic_currency
        self._foreign_currency = foreign_currency
        self._domestic_floating_index = domestic_floating_index
        self._foreign_floating_index = foreign_floating_index
        self._domestic_leg = SwapFloatLeg(effective_dt, self._termination_dt, SwapTypes.RECEIVE, domestic_spread,
                                          domestic_freq_type, domestic_dc_type, domestic_floating_index,
                                          domestic_currency, domestic_notional, 0.0, domestic_payment_lag,
                                          domestic_cal_type, domestic_bd_type, domestic_dg_type,
                                          domestic_end_of_month, True)
        self._foreign_leg = SwapFloatLeg(effective_dt, self._termination_dt, SwapTypes.PAY, foreign_spread,
                                         foreign_freq_type, foreign_dc_type, foreign_floating_index,
                                         foreign_currency, foreign_notional, 0.0, foreign_payment_lag,
                                         foreign_cal_type, foreign_bd_type, foreign_dg_type,
                                         foreign_end_of_month, True)
        self._domestic_spread = domestic_spread
        self._foreign_spread = foreign_spread
        self._adjusted_domestic_dts = self._domestic_leg._payment_dts
        self._adjusted_foreign_dts = self._foreign_leg._payment_dts

    def position(self, model):
        from .position import Position
        return Position(self, model)

    def print_payments(self):
        _print_two_legs("DOMESTIC LEG:", self._domestic_leg, "FOREIGN LEG:", self._foreign_leg, "print_payments")

    def print_valuation(self):
        _print_two_legs("DOMESTIC LEG VALUATION:", self._domestic_leg, "FOREIGN LEG VALUATION:", self._foreign_leg, "print_valuation")

    def value(self, value_dt: Date, domestic_discount_curve, foreign_discount_curve, xccy_discount_curve=None,
              xccy_discount_curve_inverted=None, spot_fx: float = None, collateral_type=None,
              first_fixing_rate_domestic: float = None, first_fixing_rate_foreign: float = None):
        """Non-AD PV in the collateral currency (xccy_basis_swap.py:209-301).  Domestic collateral (the default): domestic leg
        on its own OIS curve, foreign leg projected on the foreign OIS curve and discounted on the XCCY curve, converted as
        `foreign / spot_fx`; foreign collateral: domestic leg discounted on the inverted XCCY curve, converted as
        `domestic * spot_fx`.  (The XCCY bootstrap's own par condition is `domestic + spot_fx * foreign = 0`,
        xccy_curve.py:465-474; for a calibration swap both legs are then at par, so either conversion gives zero.)"""
        from .global_types import collateral_to_currency
        collateral_ccy = self._domestic_currency if collateral_type is None else collateral_to_currency(collateral_type)
        if collateral_ccy == self._domestic_currency:
            dom_disc, for_disc = domestic_discount_curve, xccy_discount_curve
            if for_disc is None:
                raise ValueError(f"xccy_discount_curve required for domestic collateral ({self._domestic_currency.name})")
        elif collateral_ccy == self._foreign_currency:
            dom_disc, for_disc = xccy_discount_curve_inverted, foreign_discount_curve
            if dom_disc is None:
                raise ValueError(f"xccy_discount_curve_inverted required for foreign collateral "
                                 f"({self._foreign_currency.name})")
        else:
            raise ValueError(f"Third-party collateral not yet supported: {collateral_type}. Only "
                             f"{self._domestic_currency.name} or {self._foreign_currency.name} collateral allowed.")
        pv_dom = self._domestic_leg.value(value_dt, dom_disc, domestic_discount_curve, first_fixing_rate_domestic)
        pv_for = self._foreign_leg.value(value_dt, for_disc, foreign_discount_curve, first_fixing_rate_foreign)
        if collateral_ccy == self._domestic_currency:
            return pv_dom + pv_for / spot_fx
        return pv_dom * spot_fx + pv_for


def _notional_exchange_pv(curve, value_dt: Date, effective_dt: Date, maturity_dt: Date, notional: float,
                          leg_type: SwapTypes) -> float:
    """PV of paying the notional away at the start and getting it back at maturity, from the receiver's side; the exchanges
    count while their date is on or after the value date (xccy_fix_float_swap.py:218-236, xccy_fix_fix_swap.py:232-275)."""
    pv = 0.0
    if effective_dt >= value_dt:
        pv -= notional * curve.df(effective_dt)
    if maturity_dt >= value_dt:
        pv += notional * curve.df(maturity_dt)
    return pv if leg_type == SwapTypes.RECEIVE else -pv


class _XccyFixedDomestic:
    """What XccyFixFloat and XccyFixFix share: dates, notionals, currencies and the fixed domestic leg without principal (the
    notional exchanges are added by value())."""

    def _init_domestic(self, effective_dt, term_dt_or_tenor, domestic_notional, foreign_notional, domestic_leg_type,
                       domestic_coupon, domestic_freq_type, domestic_dc_type, domestic_floating_index, foreign_floating_index,
                       domestic_currency, foreign_currency, domestic_payment_lag, domestic_cal_type, domestic_bd_type,
                       domestic_dg_type, domestic_end_of_month):
        self.derivative_type = InstrumentTypes.XCCY_SWAP
        self._termination_dt = _resolve_end(effective_dt, term_dt_or_tenor)
        self._maturity_dt = Calendar(domestic_cal_type).adjust(self._termination_dt, domestic_bd_type)
        if effective_dt > self._maturity_dt:
            raise LibError("Start date after maturity date")
        self._effective_dt = effective_dt
        self._domestic_notional, self._foreign_notional = domestic_notional, foreign_notional
        self._domestic_currency, self._foreign_currency = domestic_currency, foreign_currency
        self._domestic_floating_index, self._foreign_floating_index = domestic_floating_index, foreign_floating_index
        self._domestic_leg_type = domestic_leg_type
        self._domestic_leg = SwapFixedLeg(effective_dt, self._termination_dt, domestic_leg_type, domestic_coupon,
                                          domestic_freq_type, domestic_dc_type, domestic_floating_index, domestic_currency,
                                          domestic_notional, 0.0, domestic_payment_lag, domestic_cal_type, domestic_bd_type,
                                          domestic_dg_type, domestic_end_of_month)
        return SwapTypes.PAY if domestic_leg_type == SwapTypes.RECEIVE else SwapTypes.RECEIVE

    def print_valuation(self):
        kind = "FLOATING" if isinstance(self._foreign_leg, SwapFloatLeg) else "FIXED"
        _print_two_legs("DOMESTIC FIXED LEG VALUATION:", self._domestic_leg, f"FOREIGN {kind} LEG VALUATION:", self._foreign_leg,
                        "print_valuation")

    def _domestic_value(self, value_dt: Date, domestic_discount_curve) -> float:
        return self._domestic_leg.value(value_dt, domestic_discount_curve) + _notional_exchange_pv(
            domestic_discount_curve, value_dt, self._effective_dt, self._maturity_dt, self._domestic_notional,
            self._domestic_leg_type)


class XccyFixFloat(_XccyFixedDomestic):
    """Fixed domestic coupons against foreign floating + spread, notionals exchanged on both legs; non-AD valuation only
    (cavour/trades/rates/xccy_fix_float_swap.py:79-245).  The reference's engine route for cross-currency swaps reads
    floating-leg attributes off both legs (engine.py:1476-1511), so - there as here - Position.compute is for XccyBasisSwap."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), domestic_notional: float, foreign_notional: float,
                 domestic_leg_type: SwapTypes, domestic_coupon: float, foreign_spread: float,
                 domestic_freq_type: FrequencyTypes, foreign_freq_type: FrequencyTypes, domestic_dc_type: DayCountTypes,
                 foreign_dc_type: DayCountTypes, domestic_floating_index: CurveTypes, foreign_floating_index: CurveTypes,
                 domestic_currency: CurrencyTypes, foreign_currency: CurrencyTypes, domestic_payment_lag: int = 0,
                 foreign_payment_lag: int = 0, domestic_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 foreign_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 domestic_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 foreign_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 domestic_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 foreign_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 domestic_end_of_month: bool = False, foreign_end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        foreign_leg_type = self._init_domestic(
            effective_dt, term_dt_or_tenor, domestic_notional, foreign_notional, domestic_leg_type, domestic_coupon,
            domestic_freq_type, domestic_dc_type, domestic_floating_index, foreign_floating_index, domestic_currency,
            foreign_currency, domestic_payment_lag, domestic_cal_type, domestic_bd_type, domestic_dg_type, domestic_end_of_month)
        self._foreign_leg = SwapFloatLeg(effective_dt, self._termination_dt, foreign_leg_type, foreign_spread, foreign_freq_type,
                                         foreign_dc_type, foreign_floating_index, foreign_currency, foreign_notional, 0.0,
                                         foreign_payment_lag, foreign_cal_type, foreign_bd_type, foreign_dg_type,
                                         foreign_end_of_month, True)

    def value(self, value_dt: Date, domestic_discount_curve: DiscountCurve, foreign_discount_curve: DiscountCurve,
              xccy_discount_curve: DiscountCurve, spot_fx: float, first_fixing_rate_foreign=None) -> float:
        """Domestic-currency PV: fixed leg and its notional exchanges on the domestic OIS curve, foreign floating leg (which
        carries its own exchanges) projected on the foreign OIS curve and discounted on the XCCY curve, divided by spot."""
        check_argument_types(self.value, locals())
        foreign = self._foreign_leg.value(value_dt, xccy_discount_curve, foreign_discount_curve, first_fixing_rate_foreign)
        return self._domestic_value(value_dt, domestic_discount_curve) + foreign / spot_fx


class XccyFixFix(_XccyFixedDomestic):
    """Fixed coupons in both currencies, notionals exchanged on both legs; non-AD valuation only
    (cavour/trades/rates/xccy_fix_fix_swap.py:77-280)."""

    def __init__(self, effective_dt: Date, term_dt_or_tenor: (Date, str), domestic_notional: float, foreign_notional: float,
                 domestic_leg_type: SwapTypes, domestic_coupon: float, foreign_coupon: float,
                 domestic_freq_type: FrequencyTypes, foreign_freq_type: FrequencyTypes, domestic_dc_type: DayCountTypes,
                 foreign_dc_type: DayCountTypes, domestic_floating_index: CurveTypes, foreign_floating_index: CurveTypes,
                 domestic_currency: CurrencyTypes, foreign_currency: CurrencyTypes, domestic_payment_lag: int = 0,
                 foreign_payment_lag: int = 0, domestic_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 foreign_cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 domestic_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 foreign_bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 domestic_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 foreign_dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 domestic_end_of_month: bool = False, foreign_end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self._foreign_leg_type = self._init_domestic(
            effective_dt, term_dt_or_tenor, domestic_notional, foreign_notional, domestic_leg_type, domestic_coupon,
            domestic_freq_type, domestic_dc_type, domestic_floating_index, foreign_floating_index, domestic_currency,
            foreign_currency, domestic_payment_lag, domestic_cal_type, domestic_bd_type, domestic_dg_type, domestic_end_of_month)
        self._foreign_leg = SwapFixedLeg(effective_dt, self._termination_dt, self._foreign_leg_type, foreign_coupon,
                                         foreign_freq_type, foreign_dc_type, foreign_floating_index, foreign_currency,
                                         foreign_notional, 0.0, foreign_payment_lag, foreign_cal_type, foreign_bd_type,
                                         foreign_dg_type, foreign_end_of_month)

    def value(self, value_dt: Date, domestic_discount_curve: DiscountCurve, foreign_discount_curve: DiscountCurve,
              xccy_discount_curve: DiscountCurve, spot_fx: float) -> float:
        """Domestic-currency PV: each fixed leg with its notional exchanges, the domestic one on the domestic OIS curve, the
        foreign one on the XCCY curve (the foreign OIS curve is not used), foreign PV divided by spot."""
        check_argument_types(self.value, locals())
        foreign = self._foreign_leg.value(value_dt, xccy_discount_curve) + _notional_exchange_pv(
            xccy_discount_curve, value_dt, self._effective_dt, self._maturity_dt, self._foreign_notional, self._foreign_leg_type)
        return self._domestic_value(value_dt, domestic_discount_curve) + foreign / spot_fx
