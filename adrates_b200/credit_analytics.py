"""Stand-alone price / yield / spread / duration analytics of bonds and floating-rate notes - the reference's non-AD methods
(cavour/trades/credit/bond.py:264-875, 1027-1110; cavour/trades/credit/frn.py:225-573), mixed into `credit.Bond` / `credit.FRN`.

Host code, as in the reference: a handful of `DiscountCurve.df` look-ups on the path-A nodes per call (one vectorised look-up per
call here) and a scalar root search.  The valuation-and-Greeks path of these instruments - `Position.compute`, `BondBook` - runs on
the device and does not come through here; `key_rate_durations` is the one method that does (it is the engine's ladder rescaled).

Conventions the reference keeps and this file follows: yields and z-spreads compound continuously on ACT/365.25 from the
settlement date; the yield / duration / convexity sums treat the principal as one bullet of the face value at maturity whatever the
amortisation schedule; "modified" duration is returned equal to Macaulay duration (continuous compounding).
"""
from __future__ import annotations

import math

import numpy as np

from .dates import Date, DayCount, FrequencyTypes
from .error import LibError


def _root(f, lo: float, hi: float, guess: float, **brent_kw) -> float:
    """Bracketed root where the bracket holds, else Newton / secant from `guess` (the reference's brentq-then-newton)."""
    from scipy.optimize import brentq, newton
    try:
        return brentq(f, lo, hi, **brent_kw)
    except Exception:  # noqa: BLE001  (no sign change in the bracket)
        return newton(f, guess, maxiter=100)


class BondAnalytics:
    # ---- present value -------------------------------------------------------------------------------------------------------
    def value(self, value_dt: Date, discount_curve, z_spread: float = 0.0, settlement_dt: Date = None) -> float:
        """PV of the coupons and principal repayments after the settlement date, discount factors relative to it, optionally
        with a continuously compounded z-spread (bond.py:264-365).  Keeps the per-payment lists the reference keeps."""
        settle = value_dt if settlement_dt is None else settlement_dt
        dts = self._payment_dts
        live = np.array([d > settle for d in dts])
        rel = np.zeros(len(dts))
        if live.any():
            df = np.asarray(discount_curve.df([d for d, a in zip(dts, live) if a]), dtype=np.float64)
            if z_spread != 0.0:
                t = np.array([(d - settle) / 365.25 for d, a in zip(dts, live) if a])
                df = df * np.exp(-z_spread * t)
            rel[live] = df / discount_curve.df(settle)
        coupons = np.asarray(self._coupon_payments, dtype=np.float64)
        principal = np.asarray(self._principal_payments, dtype=np.float64)
        coupon_pv = coupons * rel
        principal_pv = np.where(principal > 0, principal * rel, 0.0)
        self._payment_dfs = [float(x) for x in rel]
        self._coupon_pvs = [float(x) for x in coupon_pv]
        self._principal_pvs = [float(x) for x in principal_pv]
        return float(coupon_pv.sum() + principal_pv.sum())

    def accrued_interest(self, settlement_dt: Date) -> float:
        """Coupon accrued since the start of the period the settlement date falls in, in currency units (bond.py:368-400)."""
        if self._is_zero_coupon:
            return 0.0
        last = self._issue_dt
        for i, pay in enumerate(self._payment_dts):
            if pay <= settlement_dt:
                last = self._accrual_end_dts[i]
            else:
                last = self._accrual_start_dts[i]
                break
        return DayCount(self._dc_type).year_frac(last, settlement_dt)[0] * self._coupon * self._face_value

    def dirty_price(self, value_dt: Date, discount_curve, z_spread: float = 0.0, settlement_dt: Date = None) -> float:
        settle = value_dt if settlement_dt is None else settlement_dt
        return self.value(value_dt, discount_curve, z_spread, settle) / self._face_value * 100.0

    def clean_price(self, value_dt: Date, discount_curve, z_spread: float = 0.0, settlement_dt: Date = None) -> float:
        settle = value_dt if settlement_dt is None else settlement_dt
        return self.dirty_price(value_dt, discount_curve, z_spread, settle) - self.accrued_interest(settle) / self._face_value * 100.0

    # ---- yields and spreads --------------------------------------------------------------------------------------------------
    def _bullet_flows(self, settlement_dt: Date):
        """(times in ACT/365.25 years, amounts) of the yield sums: live coupons, then the face value at maturity."""
        t = [(d - settlement_dt) / 365.25 for d in self._payment_dts if d > settlement_dt]
        a = [c for d, c in zip(self._payment_dts, self._coupon_payments) if d > settlement_dt]
        if self._maturity_dt > settlement_dt:
            t.append((self._maturity_dt - settlement_dt) / 365.25)
            a.append(self._face_value)
        return np.asarray(t, dtype=np.float64), np.asarray(a, dtype=np.float64)

    def _target_pv(self, settlement_dt: Date, clean_price: float) -> float:
        dirty = clean_price + self.accrued_interest(settlement_dt) / self._face_value * 100.0
        return dirty / 100.0 * self._face_value

    def yield_to_maturity(self, settlement_dt: Date, clean_price: float) -> float:
        """Continuously compounded yield that reprices the clean price (bond.py:463-513)."""
        t, a = self._bullet_flows(settlement_dt)
        target = self._target_pv(settlement_dt, clean_price)
        return _root(lambda y: float(np.sum(a * np.exp(-y * t))) - target, -0.5, 0.5, 0.05, maxiter=100)

    def current_yield(self) -> float:
        return 0.0 if self._is_zero_coupon else self._coupon

    def z_spread(self, settlement_dt: Date, discount_curve, clean_price: float) -> float:
        """Constant spread over the curve that reprices the clean price (bond.py:534-572)."""
        target = self._target_pv(settlement_dt, clean_price)
        return _root(lambda z: self.value(settlement_dt, discount_curve, z, settlement_dt) - target, -0.1, 0.5, 0.01, maxiter=100)

    def g_spread(self, settlement_dt: Date, govt_curve, clean_price: float) -> float:
        """Bond yield less the government curve's zero rate to maturity in the bond's frequency and day count (bond.py:576-609)."""
        return self.yield_to_maturity(settlement_dt, clean_price) - \
            govt_curve.zero_rate(self._maturity_dt, freq_type=self._freq_type, dc_type=self._dc_type)

    def i_spread(self, settlement_dt: Date, discount_curve, clean_price: float) -> float:
        """The same against the swap curve (bond.py:613-644)."""
        return self.g_spread(settlement_dt, discount_curve, clean_price)

    # ---- risk ----------------------------------------------------------------------------------------------------------------
    def _yield_moment(self, settlement_dt: Date, discount_curve, z_spread: float, power: int) -> float:
        ytm = self.yield_to_maturity(settlement_dt, self.clean_price(settlement_dt, discount_curve, z_spread, settlement_dt))
        t, a = self._bullet_flows(settlement_dt)
        pv = a * np.exp(-ytm * t)
        total = float(np.sum(pv))
        if total == 0.0:                    # nothing left to pay after the settlement date: the reference divides by zero here
            raise ZeroDivisionError("float division by zero")
        return float(np.sum(pv * t ** power)) / total

    def duration(self, settlement_dt: Date, discount_curve, duration_type: str = "modified", z_spread: float = 0.0) -> float:
        """PV-weighted mean time of the flows at the bond's own yield (bond.py:648-701)."""
        if duration_type.lower() not in ("macaulay", "modified"):
            raise ValueError(f"Unknown duration type: {duration_type}")
        return self._yield_moment(settlement_dt, discount_curve, z_spread, 1)

    def convexity(self, settlement_dt: Date, discount_curve, z_spread: float = 0.0) -> float:
        return self._yield_moment(settlement_dt, discount_curve, z_spread, 2)

    def dv01(self, settlement_dt: Date, discount_curve, z_spread: float = 0.0) -> float:
        """Central difference of the PV in a 1 bp parallel shift, applied through the z-spread (bond.py:752-781)."""
        bump = 0.0001
        return (self.value(settlement_dt, discount_curve, z_spread - bump, settlement_dt) -
                self.value(settlement_dt, discount_curve, z_spread + bump, settlement_dt)) / 2.0

    cs01 = dv01          # the reference computes both the same way (bond.py:834-872)

    def key_rate_durations(self, model) -> dict:
        """-delta_k / price x 1e4 per curve pillar from the engine's ladder (bond.py:785-830): this one runs on the device."""
        from .global_types import RequestTypes
        from .position import Engine
        res = Engine(model).compute(self, [RequestTypes.VALUE, RequestTypes.DELTA])
        price = res.value.amount
        return {ten: (-float(d) / price * 10000.0 if price != 0 else 0.0) for ten, d in zip(res.risk.tenors, res.risk.risk_ladder)}

    # ---- amortisation schedules (outstanding principal after each period) ----------------------------------------------------------
    @staticmethod
    def generate_equal_principal_schedule(face_value: float, num_periods: int) -> list:
        if num_periods <= 0:
            raise LibError("Number of periods must be positive")
        step = face_value / num_periods
        return [max(0.0, face_value - i * step) for i in range(1, num_periods + 1)]

    @staticmethod
    def generate_annuity_schedule(face_value: float, num_periods: int, coupon_rate: float, freq_type: FrequencyTypes) -> list:
        """Level total payment per period (bond.py:1059-1109)."""
        if num_periods <= 0:
            raise LibError("Number of periods must be positive")
        per_year = {FrequencyTypes.ANNUAL: 1, FrequencyTypes.SEMI_ANNUAL: 2, FrequencyTypes.QUARTERLY: 4,
                    FrequencyTypes.MONTHLY: 12}.get(freq_type, 1)
        rate = coupon_rate / per_year
        if rate == 0:
            return BondAnalytics.generate_equal_principal_schedule(face_value, num_periods)
        growth = (1 + rate) ** num_periods
        payment = face_value * (rate * growth) / (growth - 1)
        out, balance = [], face_value
        for _ in range(num_periods):
            balance -= payment - balance * rate
            out.append(max(0.0, balance))
        return out


class FRNAnalytics:
    def value(self, value_dt: Date, discount_curve, index_curve=None, discount_margin: float = 0.0,
              settlement_dt: Date = None) -> float:
        """PV of the projected coupons (forward off the index curve or the first fixing, plus the quoted margin, capped and
        floored) and of the face value at maturity, optionally with a continuously compounded discount margin in the note's
        day count (frn.py:225-342)."""
        if discount_curve is None:
            raise LibError("Discount curve is required")
        index_curve = discount_curve if index_curve is None else index_curve
        settle = value_dt if settlement_dt is None else settlement_dt
        dc, index_dc = DayCount(self._dc_type), DayCount(index_curve._dc_type)
        df_settle = discount_curve.df(settle, self._dc_type)
        n = len(self._payment_dts)
        self._rates, self._coupon_payments = [0.0] * n, [0.0] * n
        self._payment_dfs, self._payment_pvs = [0.0] * n, [0.0] * n
        pv, first = 0.0, True
        for i, pay in enumerate(self._payment_dts):
            if not pay > settle:
                continue
            start, end = self._start_accrued_dts[i], self._end_accrued_dts[i]
            if first and self._first_fixing_rate is not None:
                fwd = self._first_fixing_rate
                first = False
            else:
                fwd = (index_curve.df(start, self._dc_type) / index_curve.df(end, self._dc_type) - 1.0) / \
                    index_dc.year_frac(start, end)[0]
            rate = fwd + self._quoted_margin
            if self._cap_rate is not None:
                rate = min(rate, self._cap_rate)
            if self._floor_rate is not None:
                rate = max(rate, self._floor_rate)
            amount = rate * self._year_fracs[i] * self._face_value
            df = discount_curve.df(pay, self._dc_type) / df_settle
            if discount_margin != 0.0:
                df *= math.exp(-discount_margin * dc.year_frac(settle, pay)[0])
            self._rates[i], self._coupon_payments[i], self._payment_dfs[i], self._payment_pvs[i] = rate, amount, df, amount * df
            pv += amount * df
        if self._maturity_dt > settle:
            df = discount_curve.df(self._maturity_dt, self._dc_type) / df_settle
            if discount_margin != 0.0:
                df *= math.exp(-discount_margin * dc.year_frac(settle, self._maturity_dt)[0])
            pv += self._face_value * df
            if n > 0:
                self._payment_pvs[-1] += self._face_value * df
        return float(pv)

    def dirty_price(self, value_dt: Date, discount_curve, index_curve=None, discount_margin: float = 0.0,
                    settlement_dt: Date = None) -> float:
        return 100.0 * self.value(value_dt, discount_curve, index_curve, discount_margin, settlement_dt) / self._face_value

    def accrued_interest(self, settlement_dt: Date) -> float:
        """Accrued coupon per 100 of face in the current period, at the first fixing (if any) plus the margin (frn.py:371-416)."""
        dc = DayCount(self._dc_type)
        for i, pay in enumerate(self._payment_dts):
            if pay > settlement_dt and settlement_dt >= self._start_accrued_dts[i]:
                rate = self._quoted_margin + (self._first_fixing_rate if self._first_fixing_rate is not None else 0.0)
                if self._cap_rate is not None:
                    rate = min(rate, self._cap_rate)
                if self._floor_rate is not None:
                    rate = max(rate, self._floor_rate)
                return 100.0 * (rate * dc.year_frac(self._start_accrued_dts[i], settlement_dt)[0] * self._face_value) / self._face_value
        return 0.0

    def clean_price(self, value_dt: Date, discount_curve, index_curve=None, discount_margin: float = 0.0,
                    settlement_dt: Date = None) -> float:
        settle = value_dt if settlement_dt is None else settlement_dt
        return self.dirty_price(value_dt, discount_curve, index_curve, discount_margin, settle) - self.accrued_interest(settle)

    def discount_margin(self, settlement_dt: Date, discount_curve, index_curve, clean_price: float, dm_guess: float = 0.0) -> float:
        """Spread over the discount curve that reprices the clean price (frn.py:449-490)."""
        target = clean_price + self.accrued_interest(settlement_dt)
        from scipy.optimize import brentq, newton

        def err(dm):
            return self.dirty_price(settlement_dt, discount_curve, index_curve, dm, settlement_dt) - target
        try:
            return brentq(err, -0.10, 0.20, xtol=1e-8)
        except Exception:  # noqa: BLE001
            try:
                return newton(err, dm_guess, tol=1e-8, maxiter=50)
            except Exception:  # noqa: BLE001
                raise LibError(f"Failed to converge on discount margin for price {clean_price}")

    def modified_duration(self, value_dt: Date, discount_curve, index_curve=None, discount_margin: float = 0.0,
                          settlement_dt: Date = None) -> float:
        """-dP/P per unit of discount margin by a 1 bp central difference of the dirty price (frn.py:494-534)."""
        settle = value_dt if settlement_dt is None else settlement_dt
        bump = 0.0001
        p0 = self.dirty_price(value_dt, discount_curve, index_curve, discount_margin, settle)
        up = self.dirty_price(value_dt, discount_curve, index_curve, discount_margin + bump, settle)
        down = self.dirty_price(value_dt, discount_curve, index_curve, discount_margin - bump, settle)
        return -(up - down) / (2 * bump * p0)

    def dv01(self, value_dt: Date, discount_curve, index_curve=None, discount_margin: float = 0.0, settlement_dt: Date = None) -> float:
        settle = value_dt if settlement_dt is None else settlement_dt
        bump = 0.0001
        return abs(self.value(value_dt, discount_curve, index_curve, discount_margin + bump, settle) -
                   self.value(value_dt, discount_curve, index_curve, discount_margin, settle))
