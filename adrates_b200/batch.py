"""Array-based trade books: vectorised date / schedule / day-count generation and flattening.

SURVEY 8f rank 4.  The reference builds every trade as Python objects - one `Schedule`, two
legs and ~50 `Date`s per OIS, 1.24 ms per trade - before a single cashflow is valued:

    cavour/utils/date.py:529-653, 796-879     add_weekdays / add_months / add_tenor
    cavour/utils/calendar.py:139-217          Calendar.adjust (business-day roll)
    cavour/utils/schedule.py:163-270          Schedule.generate (backward / forward roll,
                                              termination adjust, duplicate filter)
    cavour/utils/day_count.py:122-330         DayCount.year_frac
    cavour/trades/rates/swap_fixed_leg.py:130-196, swap_float_leg.py:130-186   leg dates
    cavour/market/position/engine.py:2519-2539, 2858-2897                      host prep

Once the kernels value a million trades in under 2 ms that object layer IS the end-to-end
cost.  This module restates the same rules on int64 day-serial arrays (one numpy operation
per rule, whatever the number of trades) and produces the `FlatPortfolio` the CUDA library
uploads.  `OISBook.from_arrays(...)` is the array-based public entry point next to the
object-based `Portfolio.compute`; both give the same per-trade results
(tests/test_batch_cpu.py compares dates, accruals and the flat layouts exhaustively).

Calendars: WEEKEND and NONE, as in dates.py.  Day counts: every convention that needs only
the two dates (ACT/365F, ACT/360, SIMPLE, the four 30/360 variants, ACT/ACT ISDA).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .curves import OISCurve, plan_queries
from .dates import (BusDayAdjustTypes, Calendar, CalendarTypes, Date, DateGenRuleTypes, DayCountTypes, FrequencyTypes,
                    annual_frequency)
from .error import LibError
from .flatten import FlatPortfolio, group_trades

I64 = np.int64


# ======================================================================================
# dates as day serials (same epoch as dates.Date._n: days since 0000-03-01)
# ======================================================================================
def serials(dts) -> np.ndarray:
    """list[Date] | Date | int array -> int64 serials."""
    if isinstance(dts, Date):
        return np.array([dts._n], dtype=I64)
    if isinstance(dts, np.ndarray):
        return dts.astype(I64, copy=False)
    dts = list(dts)
    if dts and isinstance(dts[0], Date):
        return np.fromiter((d._n for d in dts), dtype=I64, count=len(dts))
    return np.asarray(dts, dtype=I64)


def to_dates(n) -> list:
    return [Date._of(int(x)) for x in np.asarray(n).reshape(-1)]


def ordinal(d, m, y) -> np.ndarray:
    d, m, y = (np.asarray(a, dtype=I64) for a in (d, m, y))
    yy = y - (m <= 2)
    era = yy // 400
    yoe = yy - era * 400
    mp = (m + 9) % 12
    doy = (153 * mp + 2) // 5 + d - 1
    return era * 146097 + yoe * 365 + yoe // 4 - yoe // 100 + doy


def ymd(n):
    n = np.asarray(n, dtype=I64)
    era = n // 146097
    doe = n - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = np.where(mp < 10, mp + 3, mp - 9)
    y = yoe + era * 400 + (m <= 2)
    return d, m, y


def is_leap(y):
    return ((y % 4 == 0) & (y % 100 != 0)) | (y % 400 == 0)


_MDAYS = np.array([0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=I64)


def days_in_month(m, y):
    return _MDAYS[m] + ((m == 2) & is_leap(y))


def weekday(n):
    return (np.asarray(n, dtype=I64) + 2) % 7      # 0 = Monday (0000-03-01 is a Wednesday)


def add_months(n, mm, eom: bool = False, day=None):
    """Date.add_months on arrays: same day of month (or `day`), clipped to the month length; eom=True -> month end."""
    d, m, y = ymd(n)
    k = y * 12 + (m - 1) + np.asarray(mm, dtype=I64)
    y2 = k // 12
    m2 = k - y2 * 12 + 1
    dim = days_in_month(m2, y2)
    d2 = dim if eom else np.minimum(d if day is None else day, dim)
    return ordinal(d2, m2, y2)


def add_weekdays(n, k):
    """Date.add_weekdays / Calendar.add_business_days (WEEKEND rules) in closed form: a weekend start behaves like
    the adjacent Friday (stepping forward) or Monday (stepping backward)."""
    n = np.asarray(n, dtype=I64)
    k = np.broadcast_to(np.asarray(k, dtype=I64), n.shape)
    w = weekday(n)
    nf = n - np.where(w == 5, 1, np.where(w == 6, 2, 0))
    wf = np.minimum(w, 4)
    fwd = nf + k + 2 * ((wf + k) // 5)
    nb = n + np.where(w == 5, 2, np.where(w == 6, 1, 0))
    wb = np.where(w >= 5, 0, w)
    kb = -k
    bwd = nb - kb - 2 * ((4 - wb + kb) // 5)
    return np.where(k > 0, fwd, np.where(k < 0, bwd, n))


def _cal_table(cal):
    """(CalendarTypes value, holidays table or None) of a CalendarTypes member or a dates.Calendar (INTERSECTION)."""
    if isinstance(cal, Calendar):
        return cal._cal_type, cal._table()
    if cal in (CalendarTypes.WEEKEND, CalendarTypes.NONE):
        return cal, None
    if cal == CalendarTypes.INTERSECTION:
        raise LibError("Pass the Calendar object made by create_calendar_intersection, not CalendarTypes.INTERSECTION")
    from . import holidays
    return cal, holidays.table(cal)


def adjust(n, bd_type: BusDayAdjustTypes, cal_type=CalendarTypes.WEEKEND):
    """Calendar.adjust on arrays (calendar.py:139-217): weekday arithmetic for the WEEKEND / NONE calendars, look-ups in
    the next / previous-business-day tables of adrates_b200.holidays for the holiday calendars (`cal_type` is a
    CalendarTypes member, or the Calendar object of an INTERSECTION)."""
    if type(bd_type) != BusDayAdjustTypes:
        raise LibError("Invalid type passed. Need Finbd_type")
    cal_type, tab = _cal_table(cal_type)
    n = np.asarray(n, dtype=I64)
    if cal_type == CalendarTypes.NONE or bd_type == BusDayAdjustTypes.NONE:
        return n
    following = bd_type in (BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.MODIFIED_FOLLOWING)
    modified = bd_type in (BusDayAdjustTypes.MODIFIED_FOLLOWING, BusDayAdjustTypes.MODIFIED_PRECEDING)
    if tab is not None:
        from . import holidays
        holidays.check_range(n)
        i = n - holidays.BASE
        fwd, bwd = tab.next_bd[i], tab.prev_bd[i]
        if np.any(fwd >= holidays.N_DAYS) or np.any(bwd < 0):
            holidays.check_range(np.asarray([holidays.BASE - 1]))
        first, other = (fwd, bwd) if following else (bwd, fwd)
        out = first + holidays.BASE
        if modified:
            crossed = ymd(out)[1] != ymd(n)[1]
            out = np.where(crossed, other + holidays.BASE, out)
        return out
    w = weekday(n)
    fwd = np.where(w == 5, 2, np.where(w == 6, 1, 0))
    bwd = -np.where(w == 5, 1, np.where(w == 6, 2, 0))
    first, other = (fwd, bwd) if following else (bwd, fwd)
    out = n + first
    if modified:
        crossed = ymd(out)[1] != ymd(n)[1]
        out = np.where(crossed, n + other, out)
    return out


def add_tenor(n, count, unit: str):
    """Date.add_tenor for 'D','W','M','Y' on arrays (date.py:796-879).  Year tenors are applied one year at a time
    by the reference, so a 29-Feb start drops to the 28th and stays there; month tenors keep the original day."""
    n = np.asarray(n, dtype=I64)
    c = np.broadcast_to(np.asarray(count, dtype=I64), n.shape)
    unit = unit.upper()
    if unit == "D":
        return n + c
    if unit == "W":
        return n + 7 * c
    if unit == "M":
        return add_months(n, c)
    if unit == "Y":
        d, m, _ = ymd(n)
        day = np.where((m == 2) & (d == 29) & (c != 0), 28, d)
        return add_months(n, 12 * c, day=day)
    raise LibError("Unknown tenor type in " + unit)


_FIXED_DEN = {DayCountTypes.ACT_365F: 365, DayCountTypes.ACT_360: 360, DayCountTypes.SIMPLE: 365.0}
_THIRTY = (DayCountTypes.THIRTY_360_BOND, DayCountTypes.THIRTY_E_360, DayCountTypes.THIRTY_E_360_ISDA,
           DayCountTypes.THIRTY_E_PLUS_360)


def year_frac(n1, n2, dc_type: DayCountTypes):
    """DayCount(dc_type).year_frac(dt1, dt2)[0] on arrays (two-date conventions; day_count.py:122-330)."""
    n1 = np.asarray(n1, dtype=I64)
    n2 = np.asarray(n2, dtype=I64)
    if dc_type in _FIXED_DEN:
        return (n2 - n1) / _FIXED_DEN[dc_type]
    T = DayCountTypes
    d1, m1, y1 = ymd(n1)
    d2, m2, y2 = ymd(n2)
    if dc_type in _THIRTY:
        feb1 = (m1 == 2) & (d1 == days_in_month(m1, y1))
        feb2 = (m2 == 2) & (d2 == days_in_month(m2, y2))
        d1 = np.where(d1 == 31, 30, d1)
        if dc_type == T.THIRTY_360_BOND:
            d2 = np.where((d2 == 31) & (d1 == 30), 30, d2)
        elif dc_type == T.THIRTY_E_360:
            d2 = np.where(d2 == 31, 30, d2)
        elif dc_type == T.THIRTY_E_360_ISDA:
            d1 = np.where(feb1, 30, d1)
            d2 = np.where((d2 == 31) | feb2, 30, d2)
        else:
            roll = d2 == 31
            m2 = np.where(roll, m2 + 1, m2)
            d2 = np.where(roll, 1, d2)
        return (360 * (y2 - y1) + 30 * (m2 - m1) + (d2 - d1)) / 360
    if dc_type in (T.ACT_ACT_ISDA, T.ZERO):
        den1 = np.where(is_leap(y1), 366, 365)
        den2 = np.where(is_leap(y2), 366, 365)
        same = (n2 - n1) / den1
        one = np.ones_like(y1)
        k1 = ordinal(one, one, y1 + 1) - n1
        k2 = n2 - ordinal(one, one, y2)
        return np.where(y1 == y2, same, k1 / den1 + k2 / den2 + (y2 - y1 - 1.0))
    raise LibError(f"{dc_type} needs a third date; use the object-based legs")


# ======================================================================================
# schedules
# ======================================================================================
@dataclass
class Schedules:
    """Ragged schedule dates: dates[offsets[i]:offsets[i+1]] = Schedule(...)._adjusted_dts of schedule i."""
    offsets: np.ndarray      # int64 [S+1]
    dates: np.ndarray        # int64 serials

    def __len__(self):
        return self.offsets.shape[0] - 1

    def of(self, i: int) -> np.ndarray:
        return self.dates[self.offsets[i]:self.offsets[i + 1]]


def _ragged(counts):
    """offsets, owner[row], position-in-owner[row] of a ragged layout with the given row counts."""
    counts = np.asarray(counts, dtype=I64)
    off = np.zeros(counts.shape[0] + 1, dtype=I64)
    np.cumsum(counts, out=off[1:])
    owner = np.repeat(np.arange(counts.shape[0], dtype=I64), counts)
    pos = np.arange(int(off[-1]), dtype=I64) - off[:-1][owner]
    return off, owner, pos


def roll_schedules(eff, term, freq_type: FrequencyTypes, cal_type=CalendarTypes.WEEKEND,
                   bd_type=BusDayAdjustTypes.FOLLOWING, dg_type=DateGenRuleTypes.BACKWARD,
                   adjust_termination_dt: bool = True, end_of_month: bool = False) -> Schedules:
    """`Schedule(eff, term, ...)._adjusted_dts` for arrays of (effective, termination) serials."""
    eff = np.asarray(eff, dtype=I64)
    term = np.asarray(term, dtype=I64)
    if np.any(eff >= term):
        raise LibError("Effective date must be before termination date.")
    step = int(12 / annual_frequency(freq_type))
    de, me, ye = ymd(eff)
    dt, mt, yt = ymd(term)
    M = (yt * 12 + mt) - (ye * 12 + me)            # months from the effective month to the termination month
    if dg_type == DateGenRuleTypes.BACKWARD:
        # rolls term - k*step for k = 0.. while > eff; the first one <= eff is the previous coupon date
        q, r = np.divmod(M, step)
        cnt = q + (r != 0)                          # k with k*step < M: strictly later month than eff
        same_month = np.where(q == 0, term, add_months(term, -step * q, eom=end_of_month))
        cnt = cnt + ((r == 0) & (same_month > eff))
        off, owner, pos = _ragged(cnt + 1)
        k = cnt[owner] - pos                        # pos 0 = previous coupon date, last = termination (k = 0)
        raw = add_months(term[owner], -step * k, eom=end_of_month)
        raw = np.where(k == 0, term[owner], raw)    # the termination date itself is never moved to a month end
        inner = (pos > 0) & (k > 0)
        dates = np.where(inner, adjust(raw, bd_type, cal_type), raw)
    elif dg_type == DateGenRuleTypes.FORWARD:
        # eff + k*step for k = 0.. while < term, each adjusted, then the termination date
        q, r = np.divmod(M, step)
        cnt = q + (r != 0)
        same_month = add_months(eff, step * q)
        cnt = cnt + ((r == 0) & (same_month < term))
        off, owner, pos = _ragged(cnt + 1)
        last = pos == cnt[owner]
        raw = np.where(last, term[owner], add_months(eff[owner], step * np.where(last, 0, pos)))
        dates = np.where(last, raw, adjust(raw, bd_type, cal_type))
    else:
        raise LibError("Unknown date generation rule")
    first = off[:-1]
    dates[first] = np.maximum(dates[first], eff)
    if adjust_termination_dt:
        dates[off[1:] - 1] = adjust(term, bd_type, cal_type)
    # the reference drops one head date per coinciding consecutive pair and rejects decreasing dates
    same_owner = owner[1:] == owner[:-1]
    step_d = dates[1:] - dates[:-1]
    if np.any(same_owner & (step_d < 0)):
        raise LibError("Dates are not monotonic")
    dup = np.zeros(eff.shape[0], dtype=I64)
    np.add.at(dup, owner[1:][same_owner & (step_d == 0)], 1)
    if np.any(dup):
        keep = pos >= dup[owner]
        dates = dates[keep]
        off, _, _ = _ragged(cnt + 1 - dup)
    return Schedules(off, dates)


@dataclass
class LegSchedules:
    """Accrual periods of many legs (SwapFixedLeg.generate_payments / SwapFloatLeg.generate_payment_dts)."""
    offsets: np.ndarray      # int64 [S+1] periods per leg
    start: np.ndarray        # int64 serials
    end: np.ndarray
    pay: np.ndarray
    alpha: np.ndarray        # f64 accrual fractions in the leg's day count


def leg_schedules(eff, term, freq_type, dc_type, payment_lag: int = 0, cal_type=CalendarTypes.WEEKEND,
                  bd_type=BusDayAdjustTypes.FOLLOWING, dg_type=DateGenRuleTypes.BACKWARD,
                  end_of_month: bool = False) -> LegSchedules:
    sch = roll_schedules(eff, term, freq_type, cal_type, bd_type, dg_type, True, end_of_month)
    n_dates = np.diff(sch.offsets)
    if np.any(n_dates < 2):
        raise LibError("Schedule has none or only one date")
    is_last = np.zeros(sch.dates.shape[0], dtype=bool)
    is_last[sch.offsets[1:] - 1] = True
    is_first = np.zeros(sch.dates.shape[0], dtype=bool)
    is_first[sch.offsets[:-1]] = True
    start = sch.dates[~is_last]
    end = sch.dates[~is_first]
    pay = end if payment_lag == 0 else add_weekdays(end, int(payment_lag))
    off = np.zeros(n_dates.shape[0] + 1, dtype=I64)
    np.cumsum(n_dates - 1, out=off[1:])
    return LegSchedules(off, start, end, pay, year_frac(start, end, dc_type))


# ======================================================================================
# OIS books
# ======================================================================================
def _merge_terms(owner, time, amt, n_units):
    """Sum the amounts of single-DF terms of one unit that hit the same time, drop exact zeros, order by time
    (flatten._merge_single_df_terms on ragged arrays).  Returns (unit_offsets, time, amt)."""
    if owner.shape[0] == 0:
        return np.zeros(n_units + 1, dtype=I64), time, amt
    order = np.lexsort((time, owner))
    owner, time, amt = owner[order], time[order], amt[order]
    new = np.ones(owner.shape[0], dtype=bool)
    new[1:] = (owner[1:] != owner[:-1]) | (time[1:] != time[:-1])
    head = np.flatnonzero(new)
    # left-to-right sums within each (unit, time) run, like the dict accumulation of the object path
    s = amt[head].copy()
    run = np.diff(np.append(head, owner.shape[0]))
    for j in range(1, int(run.max())):
        sel = run > j
        s[sel] += amt[head[sel] + j]
    keep = s != 0.0
    owner, time, s = owner[head][keep], time[head][keep], s[keep]
    off = np.zeros(n_units + 1, dtype=I64)
    np.cumsum(np.bincount(owner, minlength=n_units), out=off[1:])
    return off, time, s


EAGER_CHECK_MAX = 4096     # books up to this size are validated in from_arrays; larger ones by the device flattener


class OISBook:
    """A book of vanilla OIS on one curve, held as arrays (one entry per trade).

    Conventions (frequencies, day counts, calendar, roll rules, payment lag) are per book, as they are per
    currency in practice; effective date, termination date (or tenor), side, coupon, notional and spread are per trade.
    `termination` and `spread` are materialised on the host only when host code asks for them: `compute()` hands the
    arrays to the device flattener (cav_book_from_arrays), which applies the tenor rules itself."""

    def __init__(self, curve: OISCurve, effective, termination, fixed_sign, coupon, notional, spread=None,
                 fixed_freq_type: FrequencyTypes = FrequencyTypes.ANNUAL, fixed_dc_type: DayCountTypes = DayCountTypes.ACT_365F,
                 float_freq_type: FrequencyTypes = FrequencyTypes.ANNUAL, float_dc_type: DayCountTypes = DayCountTypes.THIRTY_E_360,
                 payment_lag: int = 0, cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING, dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 tenor=None, tenor_unit: str = "Y"):
        self.curve = curve
        # Narrow per-trade arrays (int32 serials, int8 sides) are kept as the caller passed them: the device flattener reads
        # them as they are (less to move over the host link); int64 / float64 views for host code are made on first use.
        self._eff_wire = self._sign_wire = self._term_wire = None
        if isinstance(effective, np.ndarray) and effective.dtype == np.int32:
            self._eff_wire, effective = effective, None
        if isinstance(termination, np.ndarray) and termination.dtype == np.int32:
            self._term_wire, termination = termination, None
        if isinstance(fixed_sign, np.ndarray) and fixed_sign.dtype == np.int8:
            self._sign_wire, fixed_sign = fixed_sign, None
        self._effective = effective           # int64 serials
        self._termination = termination       # int64 serials (unadjusted) or None while only the tenor is known
        self._tenor = tenor                   # int32 counts in `tenor_unit` ('Y' / 'M') or None
        self._tenor_unit = tenor_unit
        self._fixed_sign = fixed_sign         # f64: +1 receive fixed, -1 pay fixed
        self.coupon = coupon
        self.notional = notional
        self._spread = spread                 # f64 per trade, or None = no spreads
        self.fixed_freq_type = fixed_freq_type
        self.fixed_dc_type = fixed_dc_type
        self.float_freq_type = float_freq_type
        self.float_dc_type = float_dc_type
        self.payment_lag = payment_lag
        self.cal_type = cal_type
        self.bd_type = bd_type
        self.dg_type = dg_type
        self._validated = False

    @property
    def effective(self) -> np.ndarray:
        if self._effective is None:
            self._effective = self._eff_wire.astype(I64)
        return self._effective

    @effective.setter
    def effective(self, a):
        self._effective, self._eff_wire = a, None

    @property
    def fixed_sign(self) -> np.ndarray:
        if self._fixed_sign is None:
            self._fixed_sign = self._sign_wire.astype(np.float64)
        return self._fixed_sign

    @fixed_sign.setter
    def fixed_sign(self, a):
        self._fixed_sign, self._sign_wire = a, None

    @property
    def termination(self) -> np.ndarray:
        if self._termination is None:
            self._termination = self._term_wire.astype(I64) if self._term_wire is not None else \
                add_tenor(self.effective, self._tenor, self._tenor_unit)
        return self._termination

    @property
    def spread(self) -> np.ndarray:
        if self._spread is None:
            return np.zeros(self.n_trades)
        return self._spread

    @property
    def n_trades(self) -> int:
        return int((self._eff_wire if self._effective is None else self._effective).shape[0])

    @classmethod
    def from_arrays(cls, curve: OISCurve, effective, termination=None, tenor_years=None, tenor_months=None,
                    fixed_leg_type=None, fixed_sign=None, fixed_coupon=None, notional=1_000_000.0, float_spread=0.0,
                    **conventions) -> "OISBook":
        """effective: list[Date] or serials.  Maturity: `termination` (dates / serials), or `tenor_years` /
        `tenor_months` (integers, applied like Date.add_tenor('nY' / 'nM')).  Side: `fixed_leg_type` (SwapTypes per
        trade) or `fixed_sign` (+1 receive / -1 pay).  Arrays of the right dtype are kept as they are (no copies)."""
        from .global_types import SwapTypes
        narrow_dates = isinstance(effective, np.ndarray) and effective.dtype == np.int32 and effective.flags.c_contiguous and \
            (termination is None or (isinstance(termination, np.ndarray) and termination.dtype == np.int32))
        eff = effective if narrow_dates else serials(effective)
        n = eff.shape[0]
        term = tenor = None
        unit = "Y"
        if termination is not None:
            term = np.ascontiguousarray(termination) if narrow_dates else serials(termination)
            if term.shape[0] != n:
                raise LibError("effective and termination arrays differ in length")
        elif tenor_years is not None or tenor_months is not None:
            unit = "Y" if tenor_years is not None else "M"
            tenor = np.ascontiguousarray(np.broadcast_to(np.asarray(tenor_years if tenor_years is not None else tenor_months),
                                                         (n,)), dtype=np.int32)
        else:
            raise LibError("OISBook needs termination dates or tenors")
        if fixed_sign is None:
            if fixed_leg_type is None:
                raise LibError("fixed_leg_type must be a SwapTypes")
            if isinstance(fixed_leg_type, SwapTypes):
                fixed_leg_type = [fixed_leg_type] * n
            if any(not isinstance(s, SwapTypes) for s in fixed_leg_type):
                raise LibError("fixed_leg_type must be a SwapTypes")
            fixed_sign = np.array([+1.0 if s == SwapTypes.RECEIVE else -1.0 for s in fixed_leg_type])
        vec = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))  # noqa: E731
        if fixed_coupon is None:
            raise LibError("fixed_coupon is required")
        no_spread = np.ndim(float_spread) == 0 and float(float_spread) == 0.0
        narrow_sign = isinstance(fixed_sign, np.ndarray) and fixed_sign.dtype == np.int8 and fixed_sign.shape == (n,) and \
            fixed_sign.flags.c_contiguous
        book = cls(curve, eff, term, fixed_sign if narrow_sign else vec(fixed_sign), vec(fixed_coupon), vec(notional),
                   None if no_spread else vec(float_spread),
                   tenor=tenor, tenor_unit=unit, **conventions)
        if n <= EAGER_CHECK_MAX:
            book._validate()
        return book

    def _validate(self):
        """Start date after (adjusted) maturity date: the reference raises when the trade is built (ois.py:127-128).  Small
        books are checked in from_arrays, large ones where they are flattened (on the device: cav_book_from_arrays)."""
        if not self._validated:
            if np.any(self.effective > adjust(self.termination, self.bd_type, self.cal_type)):
                raise LibError("Start date after maturity date")
            self._validated = True

    # ---- schedule classes: trades that share (effective, termination) share both leg schedules
    def schedule_classes(self):
        span = self.termination - self.effective         # one int64 key per trade (a serial is < 2**22 until year 11000)
        if np.any(span < 0) or np.any(span >= (1 << 22)):
            raise LibError("Start date after maturity date")
        uniq, cls_of = np.unique((self.effective << 22) | span, return_inverse=True)
        eff = uniq >> 22
        return eff, eff + (uniq & ((1 << 22) - 1)), cls_of.reshape(-1).astype(I64)

    def _legs(self, eff, term):
        fixed = leg_schedules(eff, term, self.fixed_freq_type, self.fixed_dc_type, self.payment_lag, self.cal_type,
                              self.bd_type, self.dg_type)
        same = (self.float_freq_type == self.fixed_freq_type) and (self.float_dc_type == self.fixed_dc_type)
        flt = fixed if same else leg_schedules(eff, term, self.float_freq_type, self.float_dc_type, self.payment_lag,
                                               self.cal_type, self.bd_type, self.dg_type)
        return fixed, flt

    def flatten(self, dedup: bool = True, max_group: int = 256, tiles: bool = True) -> FlatPortfolio:
        """The FlatPortfolio of this book (flatten.ois_components + Flattener.finalize on arrays).

        dedup=True: one annuity / floating (/ spread-annuity) unit per schedule class, trades carry weights.
        dedup=False: one private unit per trade (payment_lag 0 only)."""
        self._validate()
        curve = self.curve
        vd = curve._value_dt._n
        eff, term, cls_of = self.schedule_classes()
        S = eff.shape[0]
        fixed, flt = self._legs(eff, term)
        vd1 = np.array([vd], dtype=I64)

        def times(n, dc):      # times_from_dates(dt, value_dt, dc) = dc.year_frac(value_dt, dt)
            return year_frac(np.broadcast_to(vd1, n.shape), n, dc)

        # fixed leg: sum_{t_i > 0} alpha_i DF(t_i)            (engine.py:2430: strictly after the value date)
        f_owner = np.repeat(np.arange(S, dtype=I64), np.diff(fixed.offsets))
        f_t = times(fixed.pay, self.fixed_dc_type)
        live = f_t > 0.0
        A = (f_owner[live], f_t[live], fixed.alpha[live])
        # floating leg: coupons with payment time >= 0 (engine.py:2695) and a positive accrual
        l_owner = np.repeat(np.arange(S, dtype=I64), np.diff(flt.offsets))
        l_tp = times(flt.pay, self.float_dc_type)
        l_ts = times(flt.start, self.float_dc_type)
        l_te = times(flt.end, self.float_dc_type)
        live = l_tp >= 0.0
        fwd = live & (flt.alpha > 0)
        has_spread = bool(np.any(self.spread != 0.0))
        # spread annuity only for classes that hold a trade with a spread (flatten.ois_components emits none otherwise)
        live = live & (np.bincount(cls_of[self.spread != 0.0], minlength=S) > 0)[l_owner]
        Sp = (l_owner[live], l_tp[live], flt.alpha[live])
        lagged = bool(np.any(l_tp[fwd] != l_te[fwd]))
        if lagged:
            return self._flatten_lagged(S, cls_of, A, (l_owner[fwd], l_ts[fwd], l_te[fwd], l_tp[fwd]), Sp, has_spread,
                                        dedup, max_group)
        F = (np.concatenate([l_owner[fwd], l_owner[fwd]]), np.concatenate([l_ts[fwd], l_te[fwd]]),
             np.concatenate([np.ones(int(fwd.sum())), -np.ones(int(fwd.sum()))]))
        wA = self.fixed_sign * self.notional * self.coupon
        wF = -self.fixed_sign * self.notional
        wS = wF * self.spread
        if dedup:
            flat = self._flatten_shared(S, cls_of, A, F, Sp if has_spread else None, wA, wF, wS, max_group)
        else:
            flat = self._flatten_private(S, cls_of, A, F, Sp if has_spread else None, wA, wF, wS)
        if tiles:
            plan = curve.path_b_plan()
            flat.with_tiles(plan.n_nodes, plan)
        return flat

    # -- shared units ------------------------------------------------------------------
    def _plan(self, t):
        plan = self.curve.path_b_plan()
        a, b, wa, wb = plan_queries(t, plan.node_time, self.curve._interp_type)
        weight = np.stack([wa, wb], axis=1)
        node = np.stack([a, b], axis=1).astype(np.int32)
        node[weight == 0.0] = 0
        return weight, node

    def _flatten_shared(self, S, cls_of, A, F, Sp, wA, wF, wS, max_group) -> FlatPortfolio:
        parts = [A, F] + ([Sp] if Sp is not None else [])
        ws = [wA, wF] + ([wS] if Sp is not None else [])
        offs, ts, amts, ids, wcols = [], [], [], [], []
        base = 0
        for (owner, t, a), w in zip(parts, ws):
            off, t, a = _merge_terms(owner, t, a, S)
            cnt = np.diff(off)
            # classes without live terms in this part get no unit; their trades point at unit 0 with weight 0
            has = cnt > 0
            uid = np.cumsum(has) - 1 + base
            offs.append(cnt[has])
            ts.append(t)
            amts.append(a)
            ids.append(np.where(has[cls_of], uid[cls_of], 0))
            wcols.append(np.where(has[cls_of], w, 0.0))
            base += int(has.sum())
        n_units = n_units_real = base
        if n_units == 0:       # every trade has matured: one zero unit keeps the layout valid
            offs, ts, amts, n_units = [np.array([1], dtype=I64)], [np.array([0.0])], [np.array([0.0])], 1
        counts = np.concatenate(offs)
        unit_offsets = np.zeros(n_units + 1, dtype=I64)
        np.cumsum(counts, out=unit_offsets[1:])
        weight, node = self._plan(np.concatenate(ts))
        return group_trades(n_units, unit_offsets, 2, np.concatenate(amts), weight.reshape(-1), node.reshape(-1),
                            np.stack(ids, axis=1).astype(np.int32), np.stack(wcols, axis=1), max_group,
                            run_key=cls_of if n_units_real else None)

    # -- private units -------------------------------------------------------------------
    def _flatten_private(self, S, cls_of, A, F, Sp, wA, wF, wS) -> FlatPortfolio:
        n = self.n_trades
        parts = [A, F] + ([Sp] if Sp is not None else [])
        ws = [wA, wF] + ([wS] if Sp is not None else [])
        # template per class: the union of the term times of its parts, one amount vector per part
        owner = np.concatenate([p[0] for p in parts])
        t = np.concatenate([p[1] for p in parts])
        part = np.concatenate([np.full(p[0].shape[0], k, dtype=I64) for k, p in enumerate(parts)])
        amt = np.concatenate([p[2] for p in parts])
        order = np.lexsort((part, t, owner))
        owner, t, part, amt = owner[order], t[order], part[order], amt[order]
        new = np.ones(owner.shape[0], dtype=bool)
        new[1:] = (owner[1:] != owner[:-1]) | (t[1:] != t[:-1])
        slot = np.cumsum(new) - 1                       # template row of every entry
        T = int(slot[-1]) + 1 if slot.shape[0] else 0
        vec = np.zeros((len(parts), T))
        np.add.at(vec, (part, slot), amt)
        t_owner, t_time = owner[new], t[new]
        t_off = np.zeros(S + 1, dtype=I64)
        np.cumsum(np.bincount(t_owner, minlength=S), out=t_off[1:])
        weight, node = self._plan(t_time)
        # trades in class order (neighbouring units bracket the same nodes); rows return through out_index
        order = np.argsort(cls_of, kind="stable")
        cls_s = cls_of[order]
        cnt = np.diff(t_off)[cls_s]
        unit_offsets = np.zeros(n + 1, dtype=I64)
        np.cumsum(cnt, out=unit_offsets[1:])
        n_terms = int(unit_offsets[-1])
        src = np.repeat(t_off[:-1][cls_s] - unit_offsets[:-1], cnt) + np.arange(n_terms, dtype=I64)
        amt_t = np.zeros(n_terms)
        for k, w in enumerate(ws):
            amt_t += np.repeat(w[order], cnt) * vec[k][src]
        return FlatPortfolio(n, n_terms, unit_offsets, 2, amt_t, np.ascontiguousarray(weight[src].reshape(-1)),
                             np.ascontiguousarray(node[src].reshape(-1)), n, 1, np.ones(n), n,
                             np.arange(n + 1, dtype=I64), np.arange(n, dtype=np.int32), order.astype(I64), np.ones(n))

    # -- payment lag: DF(s) DF(p) / DF(e) product terms (6 pairs per term) ---------------------
    def _flatten_lagged(self, S, cls_of, A, Fw, Sp, has_spread, dedup, max_group) -> FlatPortfolio:
        if not dedup:
            raise LibError("private units with a payment lag are built by the object-based Flattener")
        owner, ts, te, tp = Fw
        plan = self.curve.path_b_plan()

        def q(t):
            a, b, wa, wb = plan_queries(t, plan.node_time, self.curve._interp_type)
            return np.stack([wa, wb], axis=1), np.stack([a, b], axis=1).astype(np.int32)
        units = []          # (offsets per class, amt, weight[T,6], node[T,6]) per part
        # annuity
        offA, tA, aA = _merge_terms(A[0], A[1], A[2], S)
        w, nd = q(tA)
        z = np.zeros((tA.shape[0], 4))
        units.append((offA, aA, np.concatenate([w, z], axis=1), np.concatenate([nd, z.astype(np.int32)], axis=1)))
        # floating: per coupon  +DF(s) DF(p)/DF(e)  and  -DF(p)   (flatten.ois_components, lagged branch)
        order = np.lexsort((tp, owner))
        owner, ts, te, tp = owner[order], ts[order], te[order], tp[order]
        ws_, ns_ = q(ts)
        we_, ne_ = q(te)
        wp_, np_ = q(tp)
        prod_w = np.concatenate([ws_, -we_, wp_], axis=1)
        prod_n = np.concatenate([ns_, ne_, np_], axis=1)
        offP, tP, aP = _merge_terms(owner, tp, -np.ones(tp.shape[0]), S)
        wP, nP = q(tP)
        zP = np.zeros((tP.shape[0], 4))
        cntF = np.bincount(owner, minlength=S) + np.diff(offP)
        offF = np.zeros(S + 1, dtype=I64)
        np.cumsum(cntF, out=offF[1:])
        # unit layout: merged single-DF terms first, then the product terms (as _merge_single_df_terms orders them)
        nF = int(offF[-1])
        aF = np.empty(nF)
        wF6 = np.empty((nF, 6))
        nF6 = np.empty((nF, 6), dtype=np.int32)
        single_owner = np.repeat(np.arange(S, dtype=I64), np.diff(offP))
        single_pos = np.arange(tP.shape[0], dtype=I64) - offP[:-1][single_owner] + offF[:-1][single_owner]
        prod_pos = np.arange(owner.shape[0], dtype=I64) - np.searchsorted(owner, owner, side="left") \
            + offF[:-1][owner] + np.diff(offP)[owner]
        aF[single_pos], wF6[single_pos], nF6[single_pos] = aP, np.concatenate([wP, zP], axis=1), \
            np.concatenate([nP, zP.astype(np.int32)], axis=1)
        aF[prod_pos], wF6[prod_pos], nF6[prod_pos] = 1.0, prod_w, prod_n
        units.append((offF, aF, wF6, nF6))
        if has_spread:
            offS, tS, aS = _merge_terms(Sp[0], Sp[1], Sp[2], S)
            w, nd = q(tS)
            z = np.zeros((tS.shape[0], 4))
            units.append((offS, aS, np.concatenate([w, z], axis=1), np.concatenate([nd, z.astype(np.int32)], axis=1)))
        wA = self.fixed_sign * self.notional * self.coupon
        wFl = -self.fixed_sign * self.notional
        wts = [wA, wFl] + ([wFl * self.spread] if has_spread else [])
        counts, amts, wgt, nod, ids, wcols = [], [], [], [], [], []
        base = 0
        for (off, a, w6, n6), wt in zip(units, wts):
            cnt = np.diff(off)
            has = cnt > 0
            uid = np.cumsum(has) - 1 + base
            counts.append(cnt[has])
            amts.append(a)
            wgt.append(w6)
            nod.append(n6)
            ids.append(np.where(has[cls_of], uid[cls_of], 0))
            wcols.append(np.where(has[cls_of], wt, 0.0))
            base += int(has.sum())
        if base == 0:
            counts, amts, wgt, nod, base = [np.array([1], dtype=I64)], [np.zeros(1)], [np.zeros((1, 6))], \
                [np.zeros((1, 6), dtype=np.int32)], 1
        unit_offsets = np.zeros(base + 1, dtype=I64)
        np.cumsum(np.concatenate(counts), out=unit_offsets[1:])
        weight = np.concatenate(wgt)
        node = np.concatenate(nod)
        node[weight == 0.0] = 0
        return group_trades(base, unit_offsets, 6, np.concatenate(amts), np.ascontiguousarray(weight.reshape(-1)),
                            np.ascontiguousarray(node.reshape(-1)), np.stack(ids, axis=1).astype(np.int32),
                            np.stack(wcols, axis=1), max_group, run_key=cls_of)

    # ---- valuation -----------------------------------------------------------------------
    # ---- device flattening ------------------------------------------------------------------
    def device_conv(self):
        """cav_book_conv of this book, or None when only the host flattener handles its conventions."""
        from . import _native
        if self.payment_lag != 0:
            return None
        supported = set(_FIXED_DEN) | set(_THIRTY) | {DayCountTypes.ACT_ACT_ISDA, DayCountTypes.ZERO}
        if self.fixed_dc_type not in supported or self.float_dc_type not in supported:
            return None
        try:
            steps = [int(12 / annual_frequency(f)) for f in (self.fixed_freq_type, self.float_freq_type)]
        except (LibError, ZeroDivisionError):
            return None
        if any(st < 1 or st > 12 for st in steps):
            return None
        return _native.BookConv(self.curve._value_dt._n, steps[0], steps[1], self.fixed_dc_type.value, self.float_dc_type.value,
                                _cal_table(self.cal_type)[0].value, self.bd_type.value, self.dg_type.value, 0, 0)

    def upload(self, ctx, tiles: bool = True, dedup: bool = True, device_flatten: bool = True) -> str:
        """Make this book the portfolio of `ctx` (a _native.Context whose curve is this book's).  Shared-unit books with
        conventions the device flattener knows are flattened on the GPU from the per-trade arrays (cav_book_from_arrays: no
        flat arrays are built on the host or cross PCIe); everything else goes through flatten() + cav_portfolio_upload.
        Returns "device" or "host"."""
        from . import _native
        conv = self.device_conv() if (dedup and device_flatten and self.n_trades > 0) else None
        if conv is not None:
            try:
                unit = _native.TENOR_YEARS if self._tenor_unit == "Y" else _native.TENOR_MONTHS
                has_term = self._termination is not None or self._term_wire is not None
                narrow = self._eff_wire is not None and (not has_term or self._term_wire is not None)
                eff = self._eff_wire if narrow else self.effective
                term = None if not has_term else (self._term_wire if narrow else self.termination)
                hol = _cal_table(self.cal_type)[1]
                if hol is not None:                       # holiday calendar: the device walks its non-business-day bitmap
                    ctx.book_set_holidays(hol)
                ctx.book_from_arrays(conv, eff, term, None if has_term else self._tenor, unit,
                                     self._sign_wire if self._sign_wire is not None else self.fixed_sign, self.coupon, self.notional,
                                     self._spread, tiles=tiles)
                self._validated = True
                return "device"
            except LibError as ex:
                if ex.code != _native.E_UNSUPPORTED:
                    # the reference's own message for invalid trades (the native prefix stays in __cause__)
                    msg = str(ex).split(": ", 1)[-1]
                    raise LibError(msg, ex.code) from ex
        ctx.portfolio_upload(self.flatten(dedup=dedup, tiles=tiles))
        return "host"

    def _value(self, mask: int, device: int, dedup: bool, per_trade: bool, device_flatten: bool = True, out=None):
        """(totals [1057] on the host, {"pv", "delta", "gamma"} device rows in trade order)."""
        import torch
        from . import _native
        from .position import CurveSession
        sess = CurveSession.get(self.curve, device)
        n = self.n_trades
        if n:
            self.upload(sess.ctx, tiles=bool(mask & _native.REQ_GAMMA), dedup=dedup, device_flatten=device_flatten)
        elif mask & _native.REQ_ALLREDUCE:      # an empty shard: whatever book the context held before must not be valued again
            z64, zf = np.zeros(1, dtype=np.int64), np.zeros(0)
            sess.ctx.portfolio_upload(FlatPortfolio(0, 0, z64, 2, zf, zf, np.zeros(0, dtype=np.int32), 0, 1, zf, 0, z64,
                                                    np.zeros(0, dtype=np.int32), None, zf))
        rows = {}
        if per_trade:
            dev = torch.device("cuda", device)
            want = [(k, shape) for k, bit, shape in (("pv", _native.REQ_VALUE, (n,)), ("delta", _native.REQ_DELTA, (n, 32)),
                                                     ("gamma", _native.REQ_GAMMA, (n, 32, 32))) if mask & bit]
            for k, shape in want:
                if out is not None and k in out:        # caller-owned result rows (reused across valuations)
                    t = out[k]
                    if tuple(t.shape) != shape or t.dtype != torch.float64 or not t.is_contiguous() or t.device != dev:
                        raise LibError(f"out[{k!r}] must be a contiguous float64 {shape} tensor on {dev}")
                    rows[k] = t
                else:
                    rows[k] = torch.empty(shape, dtype=torch.float64, device=dev)
        ptr = lambda k: rows[k].data_ptr() if k in rows else None  # noqa: E731
        if n or (mask & _native.REQ_ALLREDUCE):        # an empty shard still takes part in the exchange of the totals
            agg = sess.ctx.portfolio_value_host(mask, ptr("pv"), ptr("delta"), ptr("gamma"))
        else:
            agg = np.zeros(_native.NOUT)
        return np.array(agg, dtype=np.float64), rows

    def compute(self, request_list, device: int = 0, dedup: bool = True, per_trade: bool = True, device_flatten: bool = True,
                all_ranks: bool = False, out=None):
        """One batched device valuation.  Returns (AnalyticsResult of the book totals, rows) where rows is a dict
        of torch CUDA tensors {"pv": [N], "delta": [N,32], "gamma": [N,32,32]} in trade order (per_trade=True; `out` may
        supply preallocated tensors under the same keys).  device_flatten=False forces the host flattener (same results;
        used by the parity tests).  all_ranks=True: every rank of the torch.distributed group (one process per GPU) calls
        this with ITS OWN book and gets the totals of all books, summed inside the totals kernel over NVLink."""
        from . import _native
        from .position import CurveSession, request_mask, _result_from_totals
        mask = request_mask(request_list)
        if all_ranks:
            from .parallel import init_device_allreduce
            if init_device_allreduce(CurveSession.get(self.curve, device).ctx):
                mask |= _native.REQ_ALLREDUCE
        agg, rows = self._value(mask, device, dedup, per_trade, device_flatten, out)
        # currency / index of the totals: those of the curve's calibration swaps (books are single-curve)
        return _result_from_totals(agg, mask, self.curve, self.curve._used_swaps[0]), rows

    # ---- scenario revaluation (BASELINE config 4) ------------------------------------------------
    def scenario_values(self, rates, device: int = 0, dedup: bool = True, pnl: bool = False, out=None):
        """Trade values under S shocked curves in one device pass: `rates` is [S, R] decimal par rates
        (`Model.scenario_rates(curve_name, shocks)`), the result a torch CUDA tensor [S, n_trades] in trade order
        (pnl=True: minus the values on this book's own curve).  Row s equals the per-trade VALUE of this book built on
        `model.scenario(curve_name, shocks[s])`."""
        from .scenarios import scenario_values_uploaded
        return scenario_values_uploaded(self.curve, self.n_trades, lambda ctx: self.upload(ctx, tiles=False, dedup=dedup),
                                        rates, device, pnl, out)

    def scenario_values_distributed(self, rates, device: int | None = None, dedup: bool = True, pnl: bool = False):
        """Every rank of the initialised torch.distributed group calls this with the SAME book and rates; rank r values
        its contiguous slice of the scenarios on its GPU.  No collective: returns (rows [S_r, n_trades], (lo, hi))."""
        import os
        import torch.distributed as dist
        from .scenarios import check_rates, scenario_bounds
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", rank))
        rates = check_rates(self.curve, rates)
        lo, hi = scenario_bounds(rates.shape[0], world)[rank]
        return self.scenario_values(rates[lo:hi], device, dedup, pnl), (lo, hi)

    # ---- multi-GPU: one process per GPU, trades shard by rank, totals all-reduced ----
    def shard(self, rank: int, world: int) -> "OISBook":
        """The contiguous slice of this book rank `rank` of `world` owns, balanced by coupon count."""
        from .parallel import shard_bounds
        per_year = annual_frequency(self.fixed_freq_type) + annual_frequency(self.float_freq_type)
        cost = np.maximum((self.termination - self.effective) / 365.25 * per_year, 1.0)
        lo, hi = shard_bounds(cost, world)[rank]
        sl = slice(lo, hi)
        return OISBook(self.curve, self.effective[sl], self.termination[sl], self.fixed_sign[sl], self.coupon[sl],
                       self.notional[sl], None if self._spread is None else self._spread[sl], self.fixed_freq_type,
                       self.fixed_dc_type, self.float_freq_type, self.float_dc_type, self.payment_lag, self.cal_type,
                       self.bd_type, self.dg_type)

    def shard_by_schedule(self, rank: int, world: int):
        """Strong-scaling shard: trades are dealt to the ranks by SCHEDULE (effective, termination), so that every schedule's
        shared units are valued on one rank only (a contiguous slice of a shuffled book touches almost every schedule on
        every rank and repeats the units stage `world` times).  Schedules go to ranks as contiguous key ranges with balanced
        coupon counts.  Returns (book of this rank's trades in their original order, their indices in the whole book)."""
        from .parallel import shard_bounds
        key = self.effective.astype(np.int64) * (1 << 21) + (self.termination - self.effective).astype(np.int64)
        uniq, inv = np.unique(key, return_inverse=True)
        per_year = annual_frequency(self.fixed_freq_type) + annual_frequency(self.float_freq_type)
        cost = np.maximum((self.termination - self.effective) / 365.25 * per_year, 1.0)
        cls_cost = np.bincount(inv, weights=cost, minlength=uniq.shape[0])
        lo, hi = shard_bounds(cls_cost, world)[rank]
        idx = np.nonzero((inv >= lo) & (inv < hi))[0]
        book = OISBook(self.curve, self.effective[idx], self.termination[idx], self.fixed_sign[idx], self.coupon[idx],
                       self.notional[idx], None if self._spread is None else self._spread[idx], self.fixed_freq_type,
                       self.fixed_dc_type, self.float_freq_type, self.float_dc_type, self.payment_lag, self.cal_type,
                       self.bd_type, self.dg_type)
        return book, idx

    def compute_distributed(self, request_list, device: int | None = None, dedup: bool = True, per_trade: bool = True,
                            shard: str = "range"):
        """Every rank of the initialised torch.distributed group calls this with the SAME book: each values its own
        shard on its GPU (no data-path collective) and the 1057 totals are summed over the ranks.  With an NCCL group (one
        process per GPU) the sum happens inside the totals kernel over NVLink peer memory (REQ_ALLREDUCE, csrc/cav_comm.cu):
        no collective call, one device->host read.  Other groups (gloo) all-reduce the host totals.
        Returns (AnalyticsResult of the WHOLE book, rows of this rank's shard, (lo, hi)); with shard="schedule" the trades are
        dealt by schedule (shard_by_schedule) and the third item is the index array of this rank's trades."""
        import os
        import torch
        import torch.distributed as dist
        from . import _native
        from .parallel import all_reduce_totals, init_device_allreduce, shard_bounds
        from .position import CurveSession, request_mask, _result_from_totals
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", rank))
        mask = request_mask(request_list)
        if shard not in ("range", "schedule"):
            raise LibError("shard must be 'range' or 'schedule'")
        mine, idx = self.shard_by_schedule(rank, world) if shard == "schedule" else (self.shard(rank, world), None)
        backend = dist.get_backend() if dist.is_initialized() else None
        if backend == "nccl" and world > 1:
            init_device_allreduce(CurveSession.get(self.curve, device).ctx, rank, world)
            agg, rows = mine._value(mask | _native.REQ_ALLREDUCE, device, dedup, per_trade)
            tot = agg
        else:
            agg, rows = mine._value(mask, device, dedup, per_trade)
            tot = all_reduce_totals(torch.from_numpy(agg)).numpy()
        res = _result_from_totals(tot, mask, self.curve, self.curve._used_swaps[0])
        if idx is not None:
            return res, rows, idx
        per_year = annual_frequency(self.fixed_freq_type) + annual_frequency(self.float_freq_type)
        cost = np.maximum((self.termination - self.effective) / 365.25 * per_year, 1.0)
        return (res, rows, shard_bounds(cost, world)[rank])
