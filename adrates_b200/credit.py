"""Bonds and floating-rate notes on the CUDA valuation path.

Reference: cavour/trades/credit/bond.py:80-247 (constructor and coupon schedule) and Engine._compute_bond
(cavour/market/position/engine.py:505-640): a bond is priced with the engine's fixed-leg pricer on the OIS curve of
its currency - coupons `year_frac x coupon x outstanding principal` on the payment dates plus the face value on the
last payment date, from the investor's side - so VALUE / DELTA / GAMMA come from the same kernels as an OIS fixed leg.
An FRN (cavour/trades/credit/frn.py:80-220, Engine._compute_frn engine.py:700-925) is the engine's floating-leg pricer
on the same curve: coupons (forward + quoted margin) x accrual x face, an optional known first fixing, the face value
at maturity; caps and floors are ignored by the engine (and here).  Like the reference, DELTA / GAMMA exist only when
the index curve is the discount curve of the currency.
The stand-alone price / yield / spread / duration methods of the reference's classes (bond.py:264-875, frn.py:222-573) are
non-AD host code; they are mirrored in credit_analytics.py and mixed in here.
"""
from __future__ import annotations

from .dates import (BusDayAdjustTypes, Calendar, CalendarTypes, Date, DateGenRuleTypes, DayCount, DayCountTypes,
                    FrequencyTypes, Schedule)
from .credit_analytics import BondAnalytics, FRNAnalytics
from .argcheck import check_argument_types
from .error import LibError
from .global_types import CurrencyTypes, CurveTypes, InstrumentTypes

# Engine._compute_bond: bonds discount on the OIS curve of their currency (engine.py:517-527, 603-608)
BOND_CURVE = {CurrencyTypes.GBP: CurveTypes.GBP_OIS_SONIA, CurrencyTypes.USD: CurveTypes.USD_OIS_SOFR,
              CurrencyTypes.EUR: CurveTypes.EUR_OIS_ESTR}


class Bond(BondAnalytics):
    def __init__(self, issue_dt: Date, maturity_dt_or_tenor: (Date, str), coupon: float, freq_type: FrequencyTypes,
                 dc_type: DayCountTypes, currency: CurrencyTypes, face_value: float = 100.0, payment_lag: int = 0,
                 amortization_schedule: (list, type(None)) = None, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self.derivative_type = InstrumentTypes.BOND
        self._maturity_dt = maturity_dt_or_tenor if isinstance(maturity_dt_or_tenor, Date) \
            else issue_dt.add_tenor(maturity_dt_or_tenor)
        if issue_dt >= self._maturity_dt:
            raise LibError("Issue date must be before maturity date")
        self._issue_dt = issue_dt
        self._coupon = coupon
        self._freq_type = freq_type
        self._dc_type = dc_type
        self._currency = currency
        self._face_value = face_value
        self._payment_lag = payment_lag
        self._cal_type = cal_type
        self._bd_type = bd_type
        self._dg_type = dg_type
        self._end_of_month = end_of_month
        self._amortization_schedule = amortization_schedule
        self._is_zero_coupon = (coupon == 0.0 or freq_type == FrequencyTypes.ZERO)
        if not self._is_zero_coupon:
            self._generate_coupon_schedule()
        else:
            self._payment_dts = [self._maturity_dt]
            self._year_fracs = [0.0]
            self._coupon_payments = [0.0]
            self._accrual_start_dts = [issue_dt]
            self._accrual_end_dts = [self._maturity_dt]
            self._num_coupons = 0
            self._principal_schedule = [self._face_value, 0.0]
            self._principal_payments = [self._face_value]

    def _generate_coupon_schedule(self):
        """bond.py:162-245."""
        calendar = Calendar(self._cal_type)
        dts = Schedule(self._issue_dt, self._maturity_dt, self._freq_type, self._cal_type, self._bd_type,
                       self._dg_type, end_of_month=self._end_of_month)._adjusted_dts
        n = len(dts) - 1
        if self._amortization_schedule is not None:
            if len(self._amortization_schedule) != n:
                raise LibError(f"Amortization schedule length ({len(self._amortization_schedule)}) "
                               f"must match number of payment periods ({n})")
            self._principal_schedule = [self._face_value] + list(self._amortization_schedule)
        else:
            self._principal_schedule = [self._face_value] * n + [0.0]
        dc = DayCount(self._dc_type)
        self._accrual_start_dts, self._accrual_end_dts, self._payment_dts = [], [], []
        self._year_fracs, self._coupon_payments, self._principal_payments = [], [], []
        prev = self._issue_dt
        for i, nxt in enumerate(dts[1:]):
            yf = dc.year_frac(prev, nxt)[0]
            self._accrual_start_dts.append(prev)
            self._accrual_end_dts.append(nxt)
            self._payment_dts.append(calendar.add_business_days(nxt, self._payment_lag))
            self._year_fracs.append(yf)
            self._coupon_payments.append(yf * self._coupon * self._principal_schedule[i])
            self._principal_payments.append(self._principal_schedule[i] - self._principal_schedule[i + 1])
            prev = nxt
        self._num_coupons = len(self._payment_dts)

    @property
    def _floating_index(self) -> CurveTypes:
        """Curve the engine prices this bond on (and labels its ladders with)."""
        if self._currency not in BOND_CURVE:
            raise LibError(f"No default OIS curve for currency {self._currency}")
        return BOND_CURVE[self._currency]

    def position(self, model):
        from .position import Position
        return Position(self, model)


class FRN(FRNAnalytics):
    def __init__(self, issue_dt: Date, maturity_dt_or_tenor: (Date, str), quoted_margin: float, freq_type: FrequencyTypes,
                 dc_type: DayCountTypes, currency: CurrencyTypes, floating_index: CurveTypes, face_value: float = 100.0,
                 payment_lag: int = 0, cap_rate: (float, type(None)) = None, floor_rate: (float, type(None)) = None,
                 first_fixing_rate: (float, type(None)) = None,
                 cal_type: CalendarTypes = CalendarTypes.WEEKEND, bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD, end_of_month: bool = False):
        check_argument_types(self.__init__, locals())
        self._issue_dt = issue_dt
        self._quoted_margin = quoted_margin
        self._freq_type = freq_type
        self._dc_type = dc_type
        self._currency = currency
        self._floating_index = floating_index
        self._face_value = face_value
        self._payment_lag = payment_lag
        self._cap_rate = cap_rate
        self._floor_rate = floor_rate
        self._first_fixing_rate = first_fixing_rate
        self._cal_type = cal_type
        self._bd_type = bd_type
        self._dg_type = dg_type
        self._end_of_month = end_of_month
        mat = maturity_dt_or_tenor if isinstance(maturity_dt_or_tenor, Date) else issue_dt.add_tenor(maturity_dt_or_tenor)
        calendar = Calendar(cal_type)
        self._maturity_dt = calendar.adjust(mat, bd_type)
        if issue_dt >= self._maturity_dt:
            raise LibError("Issue date must be before maturity date")
        self.derivative_type = InstrumentTypes.FRN
        self._generate_payment_schedule()

    def _generate_payment_schedule(self):
        """Accrual periods, payment dates (lagged by business days) and year fractions of the note (frn.py:173-220)"""
        calendar = Calendar(self._cal_type)
        dts = Schedule(self._issue_dt, self._maturity_dt, self._freq_type, self._cal_type, self._bd_type, self._dg_type,
                       end_of_month=self._end_of_month)._adjusted_dts
        if len(dts) < 2:
            raise LibError("Schedule must have at least two dates")
        dc = DayCount(self._dc_type)
        self._payment_dts, self._start_accrued_dts, self._end_accrued_dts = [], [], []
        self._year_fracs, self._accrued_days = [], []
        for prev, nxt in zip(dts[:-1], dts[1:]):
            self._start_accrued_dts.append(prev)
            self._end_accrued_dts.append(nxt)
            self._payment_dts.append(nxt if self._payment_lag == 0 else calendar.add_business_days(nxt, self._payment_lag))
            yf, days, _ = dc.year_frac(prev, nxt)
            self._year_fracs.append(yf)
            self._accrued_days.append(days)

    def position(self, model):
        from .position import Position
        return Position(self, model)
