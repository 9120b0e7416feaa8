"""Holiday calendars as day-serial bitmaps.

The reference decides `Calendar.is_holiday(dt)` by walking an if-chain per country for every date it is asked about
(cavour/utils/calendar.py:278-1099, Easter from a 299-entry day-of-year table :49-80).  Here each calendar is a short
declarative rule table (fixed dates, weekday-in-a-day-window rules, offsets from Easter Monday, year conditions), evaluated
ONCE with NumPy over every day of 1901..2199 - the span of the reference's Easter table - into

    holiday[i]        the reference's is_holiday for serial BASE + i (weekends count only where a rule lands on them)
    non_business[i]   weekend or holiday  (= not Calendar.is_business_day)
    next_bd[i]/prev_bd[i]   first business day at or after / last business day at or before BASE + i

so that `Calendar.adjust` / `add_business_days` are table look-ups on the host (arrays included, batch.adjust) and a
bitmap walk on the device: `non_business` packed to 32-bit words is what `cav_book_set_holidays` uploads for the device
flattener (include/adrates_b200.h).  Every bit is pinned against the unmodified reference (tests/golden/ref_calendars.npz,
generator tests/golden/gen/make_golden_calendars.py).  Easter Monday comes from the Gregorian computus; it reproduces the
reference's table for all 299 years (same test).
"""
import numpy as np

from .dates import CalendarTypes, _ordinal
from .error import LibError

YEAR_LO, YEAR_HI = 1901, 2199            # years the reference's Easter table covers (calendar.py:49-80)
BASE = _ordinal(1, 1, YEAR_LO)           # serial of bit 0
N_DAYS = _ordinal(1, 1, YEAR_HI + 1) - BASE

MON, TUE, WED, THU, FRI, SAT, SUN = range(7)


def F(m, d, **yc):
    """fixed date"""
    return ("md", m, d, d, None, yc)


def W(m, dlo, dhi, *wd, **yc):
    """month m, dlo <= day <= dhi, on one of the weekdays wd"""
    return ("md", m, dlo, dhi, wd, yc)


def E(off):
    """days from Easter Monday"""
    return ("easter", off)


GOOD_FRIDAY, EASTER_MONDAY, ASCENSION, WHIT_MONDAY = E(-3), E(0), E(38), E(49)

# One entry per rule of the reference's holiday_<country> methods (calendar.py:338-1099), merged where two of its rules
# name the same days.  Year conditions: eq / ne / gt (Italy's Republic Day since 2000, the 2021 Olympic moves in Japan,
# the 2022 Jubilee in the UK).
RULES = {
    CalendarTypes.AUSTRALIA: [
        F(1, 1), F(1, 26), W(1, 27, 28, MON), GOOD_FRIDAY, EASTER_MONDAY, F(4, 25), W(4, 26, 26, MON), W(6, 8, 14, MON),
        W(8, 1, 7, MON), W(10, 1, 7, MON), F(12, 25), F(12, 26), W(12, 27, 28, MON)],
    CalendarTypes.CANADA: [
        F(1, 1), W(1, 2, 3, MON), W(2, 15, 21, MON), GOOD_FRIDAY, W(5, 18, 24, MON), F(7, 1), W(7, 2, 3, MON), W(8, 1, 7, MON),
        W(9, 1, 7, MON), W(10, 8, 14, MON), F(11, 11), W(11, 12, 13, MON), F(12, 25), F(12, 26), W(12, 27, 27, MON),
        W(12, 28, 28, TUE)],
    CalendarTypes.FRANCE: [
        F(1, 1), EASTER_MONDAY, GOOD_FRIDAY, F(5, 1), F(5, 8), ASCENSION, WHIT_MONDAY, F(7, 14), F(8, 15), F(11, 1), F(11, 11),
        F(12, 25), F(12, 26)],
    CalendarTypes.GERMANY: [
        F(1, 1), EASTER_MONDAY, GOOD_FRIDAY, F(5, 1), ASCENSION, WHIT_MONDAY, F(10, 3), F(12, 24), F(12, 25), F(12, 26)],
    CalendarTypes.ITALY: [
        F(1, 1), F(1, 6), EASTER_MONDAY, GOOD_FRIDAY, F(4, 25), F(5, 1), F(6, 2, gt=1999), F(8, 15), F(11, 1), F(12, 8),
        F(12, 25), F(12, 26)],
    CalendarTypes.JAPAN: [
        F(1, 1), W(1, 2, 3, MON), W(1, 8, 14, MON), F(2, 11), W(2, 12, 12, MON), F(2, 23), W(2, 24, 24, MON), F(3, 20),
        W(3, 21, 21, MON), F(4, 29), W(4, 30, 30, MON), F(5, 3), F(5, 4), F(5, 5), W(5, 6, 6, MON),
        W(7, 15, 21, MON, ne=2021), F(7, 22, eq=2021), F(7, 23, eq=2021), F(8, 11, ne=2021), W(8, 12, 12, MON, ne=2021),
        W(8, 9, 9, MON, eq=2021), W(9, 15, 21, MON), F(9, 23), W(9, 24, 24, MON), W(10, 8, 14, MON, ne=2021), F(11, 3),
        W(11, 4, 4, MON), F(11, 23)],
    CalendarTypes.NEW_ZEALAND: [
        F(1, 1), W(1, 2, 3, MON), W(1, 19, 25, MON), F(2, 6), GOOD_FRIDAY, EASTER_MONDAY, F(4, 25), W(6, 1, 7, MON),
        W(10, 22, 28, MON), F(12, 25), F(12, 26), W(12, 27, 28, MON)],
    CalendarTypes.NORWAY: [
        F(1, 1), E(-4), GOOD_FRIDAY, EASTER_MONDAY, ASCENSION, WHIT_MONDAY, F(5, 1), F(5, 17), F(12, 25), F(12, 26)],
    CalendarTypes.SWEDEN: [
        F(1, 1), F(1, 6), GOOD_FRIDAY, EASTER_MONDAY, ASCENSION, F(5, 1), F(6, 6), W(6, 19, 25, FRI), F(12, 24), F(12, 25),
        F(12, 26), F(12, 31)],
    CalendarTypes.SWITZERLAND: [
        F(1, 1), F(1, 2), EASTER_MONDAY, GOOD_FRIDAY, ASCENSION, WHIT_MONDAY, F(5, 1), F(8, 1), F(12, 25), F(12, 26)],
    CalendarTypes.TARGET: [
        F(1, 1), F(5, 1), GOOD_FRIDAY, EASTER_MONDAY, F(12, 25), F(12, 26)],
    CalendarTypes.UNITED_KINGDOM: [
        F(1, 1), W(1, 2, 3, MON), EASTER_MONDAY, GOOD_FRIDAY, W(5, 1, 7, MON), W(5, 25, 31, MON), F(6, 2, eq=2022),
        F(6, 3, eq=2022), W(8, 25, 31, MON), F(12, 25), F(12, 26), W(12, 27, 28, MON, TUE)],
    CalendarTypes.UNITED_STATES: [
        F(1, 1), W(1, 2, 3, MON), W(1, 15, 21, MON), W(2, 15, 21, MON), W(5, 25, 31, MON), F(7, 4), W(7, 5, 5, MON),
        W(7, 3, 3, FRI), W(9, 1, 7, MON), W(10, 8, 14, MON), F(11, 11), W(11, 12, 12, MON), W(11, 10, 10, FRI),
        W(11, 22, 28, THU), W(12, 24, 24, FRI), F(12, 25), W(12, 26, 26, MON), W(12, 31, 31, FRI)],
}


def easter_monday_serial(y):
    """Serial of Easter Monday of year(s) y (anonymous Gregorian computus)."""
    y = np.asarray(y, dtype=np.int64)
    a, b, c = y % 19, y // 100, y % 100
    d, e = b // 4, b % 4
    f = (b + 8) // 25
    g = (b - f + 1) // 3
    h = (19 * a + b - d - g + 15) % 30
    i, k = c // 4, c % 4
    L = (32 + 2 * e + 2 * i - h - k) % 7
    mm = (a + 11 * h + 22 * L) // 451
    month = (h + L - 7 * mm + 114) // 31
    day = (h + L - 7 * mm + 114) % 31 + 1
    # serial of (day, month, y): March / April only, so the civil-from-days month shift is month - 3
    yoe = y % 400
    doy = (153 * (month - 3) + 2) // 5 + day - 1
    sunday = (y // 400) * 146097 + yoe * 365 + yoe // 4 - yoe // 100 + doy
    return sunday + 1


class _Table:
    __slots__ = ("holiday", "non_business", "next_bd", "prev_bd", "_words")

    def __init__(self, holiday):
        n = np.arange(BASE, BASE + N_DAYS, dtype=np.int64)
        weekend = ((n + 2) % 7) >= 5
        self.holiday = holiday
        self.non_business = holiday | weekend
        idx = np.arange(N_DAYS, dtype=np.int64)
        good = np.where(~self.non_business, idx, -1)
        prev = np.maximum.accumulate(good)                                     # -1 before the first business day
        nxt = np.minimum.accumulate(np.where(~self.non_business, idx, N_DAYS)[::-1])[::-1]
        self.prev_bd = prev
        self.next_bd = nxt
        self._words = None

    def words(self) -> np.ndarray:
        """non_business packed into uint32 words, bit (i & 31) of word (i >> 5) = serial BASE + i (device layout)"""
        if self._words is None:
            pad = (-N_DAYS) % 32
            bits = np.concatenate([self.non_business, np.zeros(pad, dtype=bool)]).astype(np.uint8)
            self._words = np.ascontiguousarray(np.packbits(bits, bitorder="little").view(np.uint32))
        return self._words


_ymdw = None
_tables = {}


def _calendar_columns():
    global _ymdw
    if _ymdw is None:
        n = np.arange(BASE, BASE + N_DAYS, dtype=np.int64)
        era = n // 146097
        doe = n - era * 146097
        yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
        doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
        mp = (5 * doy + 2) // 153
        d = doy - (153 * mp + 2) // 5 + 1
        m = np.where(mp < 10, mp + 3, mp - 9)
        y = yoe + era * 400 + (m <= 2)
        _ymdw = (y, m, d, (n + 2) % 7, n - easter_monday_serial(y))
    return _ymdw


def _evaluate(rules) -> np.ndarray:
    y, m, d, wd, from_em = _calendar_columns()
    hol = np.zeros(N_DAYS, dtype=bool)
    for r in rules:
        if r[0] == "easter":
            hol |= from_em == r[1]
            continue
        _, mm, dlo, dhi, wds, yc = r
        hit = (m == mm) & (d >= dlo) & (d <= dhi)
        if wds:
            hit &= np.isin(wd, wds)
        for op, val in yc.items():
            hit &= {"eq": y == val, "ne": y != val, "gt": y > val}[op]
        hol |= hit
    return hol


def table(cal_type, constituents=()) -> _Table:
    """Tables of one calendar; INTERSECTION = a day is a holiday if it is one in ANY constituent (calendar.py:284-286)."""
    if cal_type == CalendarTypes.INTERSECTION:
        key = (cal_type,) + tuple(sorted(c.value for c in constituents))
        if key not in _tables:
            hol = np.zeros(N_DAYS, dtype=bool)
            for c in constituents:
                hol |= table(c).holiday
            _tables[key] = _Table(hol)
        return _tables[key]
    if cal_type not in _tables:
        if cal_type == CalendarTypes.NONE:
            raise LibError("The NONE calendar has no holiday table")
        if cal_type == CalendarTypes.WEEKEND:
            n = np.arange(BASE, BASE + N_DAYS, dtype=np.int64)
            _tables[cal_type] = _Table(((n + 2) % 7) >= 5)
        else:
            _tables[cal_type] = _Table(_evaluate(RULES[cal_type]))
    return _tables[cal_type]


def check_range(n):
    n = np.asarray(n)
    if n.size and (n.min() < BASE or n.max() >= BASE + N_DAYS):
        raise LibError(f"Holiday calendars cover {YEAR_LO}-{YEAR_HI} (the span of the reference's Easter table)")
