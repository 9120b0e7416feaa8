"""Host-side date arithmetic for the cashflow flattener: Date, Calendar, DayCount, Schedule.

These mirror the reference's interface (same class/method names, argument meaning and
error behaviour) so that trades written against the reference build unchanged:

    cavour/utils/date.py        Date (d,m,y), add_days/add_weekdays/add_months/add_tenor   :234-880
    cavour/utils/calendar.py    Calendar.adjust, BusDayAdjustTypes, DateGenRuleTypes        :83-217
    cavour/utils/day_count.py   DayCount.year_frac, DayCountTypes                           :91-370
    cavour/utils/schedule.py    Schedule.generate (ISDA backward/forward)                   :163-270
    cavour/utils/frequency.py   FrequencyTypes, annual_frequency

The implementation is independent: dates are proleptic-Gregorian day ordinals (integer
arithmetic).  WEEKEND / NONE calendars are closed-form; the national / TARGET holiday calendars
are day-serial tables built from rule lists (adrates_b200/holidays.py).
"""
from __future__ import annotations

import math
from enum import Enum

from .argcheck import check_argument_types
from .error import LibError

_MONTH_DAYS = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)
_MONTH_NAMES = ("JAN", "FEB", "MAR", "APR", "MAY", "JUN", "JUL", "AUG", "SEP", "OCT", "NOV", "DEC")
_DAY_NAMES = ("MON", "TUE", "WED", "THU", "FRI", "SAT", "SUN")


class DateFormatTypes(Enum):
    """How `repr(Date)` / `str(Date)` print (reference date.py:46-66); the reference's default is UK_LONG."""
    BLOOMBERG = 1
    US_SHORT = 2
    US_MEDIUM = 3
    US_LONG = 4
    US_LONGEST = 5
    UK_SHORT = 6
    UK_MEDIUM = 7
    UK_LONG = 8
    UK_LONGEST = 9
    DATETIME = 10


_date_format = DateFormatTypes.UK_LONG


def set_date_format(format_type) -> None:
    global _date_format
    _date_format = format_type


def is_leap_year(y: int) -> bool:
    return (y % 4 == 0 and y % 100 != 0) or (y % 400 == 0)


def days_in_month(m: int, y: int) -> int:
    if m == 2 and is_leap_year(y):
        return 29
    return _MONTH_DAYS[m - 1]


def _ordinal(d: int, m: int, y: int) -> int:
    """Days since 0000-03-01 (civil-from-days algorithm); Monday-aligned via weekday()."""
    yy = y - (m <= 2)
    era = yy // 400
    yoe = yy - era * 400
    mp = (m + 9) % 12
    doy = (153 * mp + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe


def _from_ordinal(n: int):
    era = n // 146097
    doe = n - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + 3 if mp < 10 else mp - 9
    return d, m, y + (m <= 2)


class Date:
    """Calendar date; `dt2 - dt1` is the day difference (reference date.py:421-425)."""

    MON, TUE, WED, THU, FRI, SAT, SUN = range(7)
    __slots__ = ("_d", "_m", "_y", "_n", "_hh", "_mm", "_ss", "_t")

    def __init__(self, d: int, m: int, y: int, hh: int = 0, mm: int = 0, ss: int = 0):
        """Day, month, four-digit year from 1900 on, optionally a time of day (reference date.py:248-326).  `_n` is the
        integer day serial all date arithmetic runs on; `_t` adds the time of day for ordering and differences."""
        if 1900 <= d < 2100 and 0 < y <= 31:
            raise LibError("Date arguments must now be in the order Date(dd, mm, yyyy)")
        if y < 1900:
            raise LibError("Year cannot be before 1900")
        if not (1 <= m <= 12) or d < 1 or d > days_in_month(m, y):
            raise LibError(f"Date: invalid day/month/year {d}/{m}/{y}")
        if hh < 0 or hh > 23:
            raise LibError("Hours must be in range 0-23")
        if mm < 0 or mm > 59:
            raise LibError("Minutes must be in range 0-59")
        if ss < 0 or ss > 59:
            raise LibError("Seconds must be in range 0-59")
        self._d, self._m, self._y = int(d), int(m), int(y)
        self._hh, self._mm, self._ss = hh, mm, ss
        self._n = _ordinal(self._d, self._m, self._y)
        self._t = self._n if not (hh or mm or ss) else self._n + (hh / 24.0 + mm / 24.0 / 60.0 + ss / 24.0 / 60.0 / 60.0)

    @classmethod
    def _of(cls, n: int) -> "Date":
        return cls(*_from_ordinal(n))

    @classmethod
    def from_string(cls, date_string: str, formatString: str) -> "Date":
        """Date.from_string("15-06-2023", "%d-%m-%Y") (reference date.py:353-360)"""
        import datetime
        t = datetime.datetime.strptime(date_string, formatString)
        return cls(t.day, t.month, t.year)

    @classmethod
    def from_date(cls, date) -> "Date":
        """From a datetime.date or a numpy datetime64 (reference date.py:362-380)"""
        import datetime
        import numpy as np
        if isinstance(date, np.datetime64):
            date = date.astype("datetime64[D]").astype(datetime.date)
        if isinstance(date, datetime.date):
            return cls(date.day, date.month, date.year)
        raise LibError("from_date needs a datetime.date or a numpy datetime64")

    def datetime(self):
        import datetime
        return datetime.date(self._y, self._m, self._d)

    def d(self): return self._d
    def m(self): return self._m
    def y(self): return self._y
    def serial(self) -> int: return self._n

    def excel_dt(self) -> float:
        """Days since the turn of 1900 in the spreadsheet convention the reference counts in (1 January 1900 is day 1 and the
        year 1900 is given a 29 February), plus the time of day as a fraction (reference date.py:343-345, 382-392)."""
        base = 693899 if self._n >= _ordinal(1, 3, 1900) else 693900
        return float(self._n - base) + (self._t - self._n)

    def weekday(self) -> int:
        # 0000-03-01 is a Wednesday in the proleptic Gregorian calendar
        return (self._n + 2) % 7

    def is_weekend(self) -> bool:
        return self.weekday() >= Date.SAT

    def eom(self) -> "Date":
        return Date(days_in_month(self._m, self._y), self._m, self._y)

    def is_eom(self) -> bool:
        return self._d == days_in_month(self._m, self._y)

    def __gt__(self, o): return self._t > o._t
    def __lt__(self, o): return self._t < o._t
    def __ge__(self, o): return self._t >= o._t
    def __le__(self, o): return self._t <= o._t
    def __eq__(self, o): return isinstance(o, Date) and self._t == o._t
    def __ne__(self, o): return not self.__eq__(o)
    def __hash__(self): return hash(self._t)
    def __sub__(self, o): return self._t - o._t

    @staticmethod
    def _next_quarter_end_month(m: int, d: int, y: int, roll_day: int):
        """(month, year) of the next March / June / September / December date: the quarter month of m, or the following one
        when the date is already on or past that month's roll day."""
        q = 3 * ((m + 2) // 3)
        if m == q and d >= roll_day:
            q += 3
        return (q, y) if q <= 12 else (3, y + 1)

    def next_cds_date(self, mm: int = 0) -> "Date":
        """The 20th of the next March / June / September / December after this date moved by mm months (date.py:698-732)."""
        base = self.add_months(mm)
        m, y = Date._next_quarter_end_month(base._m, base._d, base._y, 20)
        return Date(20, m, y)

    def third_wednesday_of_month(self, m: int, y: int) -> int:
        """Day of the month of the third Wednesday - between the 15th and the 21st (date.py:736-755)."""
        for d in range(15, 22):
            if Date(d, m, y).weekday() == Date.WED:
                return d
        raise LibError("Third Wednesday not found")

    def next_imm_date(self) -> "Date":
        """Next third Wednesday of March / June / September / December after this date (date.py:759-795)."""
        q = 3 * ((self._m + 2) // 3)
        m, y = Date._next_quarter_end_month(self._m, self._d, self._y, self.third_wednesday_of_month(q, self._y))
        return Date(self.third_wednesday_of_month(m, y), m, y)

    def add_hours(self, hours) -> "Date":
        """The date `hours` later, minutes and seconds kept (reference date.py:487-503)."""
        if hours < 0:
            raise LibError("Number of hours must be positive")
        total = self._hh + hours
        day = self.add_days(int(total / 24))
        return Date(day._d, day._m, day._y, total % 24, self._mm, self._ss)

    def add_days(self, num_days: int = 1) -> "Date":
        return Date._of(self._n + int(num_days))

    def add_weekdays(self, num_days: int) -> "Date":
        """Step over Saturdays/Sundays only (reference date.py:529-593)."""
        if not isinstance(num_days, int):
            raise LibError("Num days must be an integer")
        step = 1 if num_days > 0 else -1
        left, n = abs(num_days), self._n
        while left > 0:
            n += step
            if (n + 2) % 7 < Date.SAT:
                left -= 1
        return Date._of(n)

    def add_months(self, mm):
        """Same day-of-month mm months on, clipped to the month length (date.py:597-653)."""
        if isinstance(mm, (list, tuple)):
            return [self.add_months(x) for x in mm]
        if int(mm) != mm:
            raise LibError("Must only pass integers or float integers.")
        k = self._y * 12 + (self._m - 1) + int(mm)
        y, m = divmod(k, 12)
        m += 1
        return Date(min(self._d, days_in_month(m, y)), m, y)

    def add_years(self, yy):
        if isinstance(yy, (list, tuple)):
            return [self.add_years(x) for x in yy]
        mmi = int(yy * 12.0)
        ddi = int((yy * 12.0 - mmi) * (365.242 / 12.0))
        return self.add_months(mmi).add_days(ddi)

    def add_tenor(self, tenor):
        """'ON','TN','nD','nW','nM','nY' (case-insensitive); unadjusted (date.py:796-879).

        Months/years are applied one period at a time like the reference, so a 29-Feb
        start rolls to the 28th and stays there for year tenors, while month tenors
        restore the original day-of-month when the final month allows it."""
        if isinstance(tenor, list):
            for ten in tenor:
                if not isinstance(ten, str):
                    raise LibError("Tenor must be a string e.g. '5Y'")
            return [self.add_tenor(t) for t in tenor]
        if not isinstance(tenor, str):
            raise LibError("Tenor must be a string e.g. '5Y'")
        ts = tenor.upper()
        if ts in ("ON", "TN"):
            return self.add_days(1)
        unit = ts[-1]
        if unit not in "DWMY":
            raise LibError("Unknown tenor type in " + tenor)
        try:
            n = int(ts[:-1])
        except ValueError:
            raise LibError("Unknown tenor type in " + tenor)
        if unit == "D":
            return self.add_days(n)
        if unit == "W":
            return self.add_days(7 * n)
        sgn = 1 if n >= 0 else -1
        dt = self
        if unit == "M":
            for _ in range(abs(n)):
                dt = dt.add_months(sgn)
            return Date(min(self._d, days_in_month(dt._m, dt._y)), dt._m, dt._y)
        for _ in range(abs(n)):
            dt = dt.add_months(12 * sgn)
        return dt

    def __repr__(self):
        """The date in the format chosen by set_date_format (reference date.py:908-1008)."""
        day, mon, name, year = f"{self._d:02d}", f"{self._m:02d}", _MONTH_NAMES[self._m - 1], str(self._y)
        f = _date_format
        if f == DateFormatTypes.UK_LONGEST:
            return f"{_DAY_NAMES[self.weekday()]} {day} {name} {year}"
        if f == DateFormatTypes.UK_LONG:
            return f"{day}-{name}-{year}"
        if f == DateFormatTypes.UK_MEDIUM:
            return f"{day}/{mon}/{year}"
        if f == DateFormatTypes.UK_SHORT:
            return f"{day}/{mon}/{year[2:]}"
        if f == DateFormatTypes.US_LONGEST:
            return f"{_DAY_NAMES[self.weekday()]} {name} {day} {year}"
        if f == DateFormatTypes.US_LONG:
            return f"{name}-{day}-{year}"
        if f == DateFormatTypes.US_MEDIUM:
            return f"{mon}-{day}-{year}"
        if f == DateFormatTypes.US_SHORT:
            return f"{mon}-{day}-{year[2:]}"
        if f == DateFormatTypes.BLOOMBERG:
            return f"{mon}/{day}/{year[2:]}"
        if f == DateFormatTypes.DATETIME:
            return f"{day}/{mon}/{year} {self._hh:02d}:{self._mm:02d}:{self._ss:02d}"
        raise LibError("Unknown date format")

    def str(self):
        return f"{self._d:02d}{_MONTH_NAMES[self._m - 1]}{self._y}"


def datediff(d1: Date, d2: Date) -> int:
    return d2 - d1


def from_datetime(dt) -> Date:
    """Date of anything with .day / .month / .year (reference date.py:1051-1056)."""
    return Date(dt.day, dt.month, dt.year)


def daily_working_day_schedule(start_dt: Date, end_dt: Date) -> list:
    """start_dt, then every following weekday until end_dt is reached or passed (reference date.py:1024-1037)."""
    out, dt = [start_dt], start_dt
    while dt < end_dt:
        dt = dt.add_weekdays(1)
        out.append(dt)
    return out


def date_range(start_dt: Date, end_dt: Date, tenor: str = "1D") -> list:
    """Dates from start_dt to end_dt, both included, `tenor` apart (reference date.py:1075-1093); the end date closes the list
    even when the stride steps over it."""
    if start_dt > end_dt:
        return []
    out, dt = [], start_dt
    while dt < end_dt:
        out.append(dt)
        dt = dt.add_tenor(tenor)
    out.append(end_dt)
    return out


# --------------------------------------------------------------------------------------
class FrequencyTypes(Enum):
    ZERO = -1
    SIMPLE = 0
    ANNUAL = 1
    SEMI_ANNUAL = 2
    TRI_ANNUAL = 3
    QUARTERLY = 4
    MONTHLY = 12
    CONTINUOUS = 99


def annual_frequency(freq_type: FrequencyTypes) -> float:
    if not isinstance(freq_type, FrequencyTypes):
        raise LibError("Unknown frequency type")
    table = {FrequencyTypes.CONTINUOUS: -1, FrequencyTypes.ZERO: 1.0, FrequencyTypes.ANNUAL: 1.0,
             FrequencyTypes.SEMI_ANNUAL: 2.0, FrequencyTypes.TRI_ANNUAL: 3.0, FrequencyTypes.QUARTERLY: 4.0,
             FrequencyTypes.MONTHLY: 12.0}
    return table.get(freq_type)          # SIMPLE has no annual frequency: None, as the reference's function falls through


class BusDayAdjustTypes(Enum):
    NONE = 1
    FOLLOWING = 2
    MODIFIED_FOLLOWING = 3
    PRECEDING = 4
    MODIFIED_PRECEDING = 5


class CalendarTypes(Enum):
    NONE = 1
    WEEKEND = 2
    AUSTRALIA = 3
    CANADA = 4
    FRANCE = 5
    GERMANY = 6
    ITALY = 7
    JAPAN = 8
    NEW_ZEALAND = 9
    NORWAY = 10
    SWEDEN = 11
    SWITZERLAND = 12
    TARGET = 13
    UNITED_STATES = 14
    UNITED_KINGDOM = 15
    INTERSECTION = 16


class DateGenRuleTypes(Enum):
    FORWARD = 1
    BACKWARD = 2


class Calendar:
    """Business-day calendar (reference calendar.py:117-335).  WEEKEND and NONE are closed-form weekday arithmetic (the hot
    path's default, ois.py:114); the thirteen national / TARGET calendars and INTERSECTION are look-ups in the day-serial
    tables of `adrates_b200.holidays` (rule tables evaluated once over 1901-2199, pinned bit for bit against the
    reference's is_holiday)."""

    def __init__(self, cal_type: CalendarTypes, constituent_calendars=None):
        if cal_type not in CalendarTypes:
            raise LibError("Need to pass FinCalendarType and not " + str(cal_type))
        self._cal_type = cal_type
        self._constituent_calendars = constituent_calendars or []

    def _table(self):
        """holidays._Table of this calendar, or None for NONE / WEEKEND (no table needed)"""
        if self._cal_type in (CalendarTypes.NONE, CalendarTypes.WEEKEND):
            return None
        from . import holidays
        if self._cal_type == CalendarTypes.INTERSECTION:
            flat = []
            for c in self._constituent_calendars:
                flat.extend(c._leaf_types())
            return holidays.table(CalendarTypes.INTERSECTION, tuple(flat))
        return holidays.table(self._cal_type)

    def _leaf_types(self):
        if self._cal_type == CalendarTypes.INTERSECTION:
            out = []
            for c in self._constituent_calendars:
                out.extend(c._leaf_types())
            return out
        return [] if self._cal_type == CalendarTypes.NONE else [self._cal_type]

    @staticmethod
    def _slot(dt: Date) -> int:
        from . import holidays
        i = dt._n - holidays.BASE
        if i < 0 or i >= holidays.N_DAYS:
            raise LibError(f"Holiday calendars cover {holidays.YEAR_LO}-{holidays.YEAR_HI} (the span of the reference's Easter table)")
        return i

    def is_holiday(self, dt: Date) -> bool:
        if self._cal_type == CalendarTypes.NONE:
            return False
        if self._cal_type == CalendarTypes.WEEKEND:
            return dt.is_weekend()
        return bool(self._table().holiday[self._slot(dt)])

    def is_business_day(self, dt: Date) -> bool:
        if self._cal_type == CalendarTypes.INTERSECTION and not self._constituent_calendars:
            return True                      # all() over no calendars
        if dt.is_weekend():                  # every calendar, NONE included, answers False at weekends (calendar.py:268-270)
            return False
        return not self.is_holiday(dt)

    def adjust(self, dt: Date, bd_type: BusDayAdjustTypes) -> Date:
        """Roll a non-business day (reference calendar.py:139-217)."""
        if type(bd_type) != BusDayAdjustTypes:
            raise LibError("Invalid type passed. Need Finbd_type")
        if self._cal_type == CalendarTypes.NONE or bd_type == BusDayAdjustTypes.NONE:
            return dt
        if bd_type not in (BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.MODIFIED_FOLLOWING,
                           BusDayAdjustTypes.PRECEDING, BusDayAdjustTypes.MODIFIED_PRECEDING):
            raise LibError("Unknown adjustment convention" + str(bd_type))
        fwd = bd_type in (BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.MODIFIED_FOLLOWING)
        step = 1 if fwd else -1
        out = dt
        while not self.is_business_day(out):
            out = out.add_days(step)
        modified = bd_type in (BusDayAdjustTypes.MODIFIED_FOLLOWING, BusDayAdjustTypes.MODIFIED_PRECEDING)
        if modified and out._m != dt._m:
            out = dt
            while not self.is_business_day(out):
                out = out.add_days(-step)
        return out

    def add_business_days(self, start_dt: Date, num_days: int) -> Date:
        if not isinstance(num_days, int):
            raise LibError("Num days must be an integer")
        step = 1 if num_days >= 0 else -1
        left, dt = abs(num_days), start_dt
        while left > 0:
            dt = dt.add_days(step)
            if self.is_business_day(dt):
                left -= 1
        return dt

    def get_holiday_list(self, year: int):
        """Business-day holidays of a year as strings (reference calendar.py:1107-1121)."""
        out, dt, end = [], Date(1, 1, year), Date(1, 1, year + 1)
        while dt < end:
            if not self.is_business_day(dt) and not dt.is_weekend():
                out.append(f"{dt._d:02d}-{_MONTH_NAMES[dt._m - 1]}-{dt._y}")      # the reference's default date format
            dt = dt.add_days(1)
        return out

    def easter_monday(self, year: int) -> Date:
        """Reference calendar.py:1124-1137 (its Easter table; here the Gregorian computus, equal on 1901-2199)."""
        if year > 2100:
            raise LibError("Unable to determine Easter monday in year " + str(year))
        from . import holidays
        return Date._of(int(holidays.easter_monday_serial(year)))

    def __str__(self):
        return self._cal_type.name

    def __repr__(self):
        return self._cal_type.name


def create_calendar_intersection(*calendars) -> Calendar:
    """A date is a business day only if it is one in ALL calendars (reference calendar.py:1153-1176)."""
    if len(calendars) < 2:
        raise LibError("Need at least 2 calendars to create intersection")
    for cal in calendars:
        if not isinstance(cal, Calendar):
            raise LibError("All arguments must be Calendar objects")
    return Calendar(CalendarTypes.INTERSECTION, list(calendars))


# --------------------------------------------------------------------------------------
class DayCountTypes(Enum):
    ZERO = 0
    THIRTY_360_BOND = 1
    THIRTY_E_360 = 2
    THIRTY_E_360_ISDA = 3
    THIRTY_E_PLUS_360 = 4
    ACT_ACT_ISDA = 5
    ACT_ACT_ICMA = 6
    ACT_365F = 7
    ACT_360 = 8
    ACT_365L = 9
    SIMPLE = 10


def _last_day_of_feb(dt: Date) -> bool:
    return dt._m == 2 and dt._d == days_in_month(2, dt._y)


def is_last_day_of_feb(dt: Date):
    """The reference's public helper (day_count.py:64-74): True on the last day of February, False outside February - and
    None on the other February days (its `if` falls through there)."""
    if dt._m != 2:
        return False
    return True if _last_day_of_feb(dt) else None


class DayCount:
    """year_frac(dt1, dt2) -> (fraction, numerator, denominator) (day_count.py:122-330)."""

    def __init__(self, dccType: DayCountTypes):
        if dccType not in DayCountTypes:
            raise LibError("Need to pass FinDayCountType")
        self._type = dccType

    def year_frac(self, dt1: Date, dt2: Date, dt3: Date = None,
                  freq_type: FrequencyTypes = FrequencyTypes.ANNUAL, isTerminationDate: bool = False):
        t = self._type
        d1, m1, y1 = dt1._d, dt1._m, dt1._y
        d2, m2, y2 = dt2._d, dt2._m, dt2._y
        T = DayCountTypes
        if t in (T.THIRTY_360_BOND, T.THIRTY_E_360, T.THIRTY_E_360_ISDA, T.THIRTY_E_PLUS_360):
            if d1 == 31:
                d1 = 30
            if t == T.THIRTY_360_BOND:
                if d2 == 31 and d1 == 30:
                    d2 = 30
            elif t == T.THIRTY_E_360:
                if d2 == 31:
                    d2 = 30
            elif t == T.THIRTY_E_360_ISDA:
                if _last_day_of_feb(dt1):
                    d1 = 30
                if d2 == 31:
                    d2 = 30
                if _last_day_of_feb(dt2) and not isTerminationDate:
                    d2 = 30
            else:  # 30E+/360: a 31st end date rolls to the 1st of the next month
                if d2 == 31:
                    m2, d2 = m2 + 1, 1
            num = 360 * (y2 - y1) + 30 * (m2 - m1) + (d2 - d1)
            return num / 360, num, 360
        if t in (T.ACT_ACT_ISDA, T.ZERO):
            den1 = 366 if is_leap_year(y1) else 365
            den2 = 366 if is_leap_year(y2) else 365
            if y1 == y2:
                num = dt2 - dt1
                return num / den1, num, den1
            n1 = Date(1, 1, y1 + 1) - dt1
            n2 = dt2 - Date(1, 1, y2)
            return n1 / den1 + n2 / den2 + (y2 - y1 - 1.0), n1 + n2, den1 + den2
        if t == T.ACT_ACT_ICMA:
            freq = annual_frequency(freq_type)
            if dt3 is None or freq is None:
                raise LibError("ACT_ACT_ICMA requires three dates and a freq")
            num = dt2 - dt1
            den = freq * (dt3 - dt1)
            return num / den, num, den
        if t == T.ACT_365F:
            num = dt2 - dt1
            return num / 365, num, 365
        if t == T.ACT_360:
            num = dt2 - dt1
            return num / 360, num, 360
        if t == T.ACT_365L:
            frequency = annual_frequency(freq_type)
            y3 = y2 if dt3 is None else dt3._y
            num, den = dt2 - dt1, 365
            if is_leap_year(y1):
                feb29 = Date(29, 2, y1)
            elif is_leap_year(y3):
                feb29 = Date(29, 2, y3)
            else:
                feb29 = Date(1, 1, 1900)
            if frequency == 1:
                if dt3 is not None and feb29 > dt1 and feb29 <= dt3:
                    den = 366
            elif is_leap_year(y3):
                den = 366
            return num / den, num, den
        if t == T.SIMPLE:
            num = dt2 - dt1
            return num / 365.0, num, 365.0
        raise LibError(str(t) + " is not one of DayCountTypes")

    def days_in_year(self):
        t, T = self._type, DayCountTypes
        if t in (T.THIRTY_360_BOND, T.THIRTY_E_360, T.THIRTY_E_360_ISDA, T.THIRTY_E_PLUS_360, T.ACT_360):
            return 360
        if t is T.ACT_365F:
            return 365
        if t is T.SIMPLE:
            return 365.0
        raise LibError(f"No fixed days-in-year defined for convention {t}")

    def __repr__(self):
        return str(self._type)


# --------------------------------------------------------------------------------------
class Schedule:
    """ISDA coupon schedule; `_adjusted_dts[0]` is the previous coupon date
    (reference schedule.py:163-270, including its quirks: the FORWARD rule's first
    generated date repeats the effective date and is then dropped by the duplicate
    filter; the termination date is adjusted by default)."""

    def __init__(self, effective_dt: Date, termination_dt: Date,
                 freq_type: FrequencyTypes = FrequencyTypes.ANNUAL,
                 cal_type: CalendarTypes = CalendarTypes.WEEKEND,
                 bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING,
                 dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD,
                 adjust_termination_dt: bool = True, end_of_month: bool = False,
                 first_dt=None, next_to_last_dt=None):
        check_argument_types(self.__init__, locals())
        if effective_dt >= termination_dt:
            raise LibError("Effective date must be before termination date.")
        self._effective_dt = effective_dt
        self._termination_dt = termination_dt
        self._freq_type = freq_type
        self._cal_type = cal_type
        self._bd_type = bd_type
        self._dg_type = dg_type
        self._adjust_termination_dt = adjust_termination_dt
        self._end_of_month = end_of_month is True
        self._adjusted_dts = None
        self.generate()

    def schedule_dts(self):
        return self._adjusted_dts

    def generate(self):
        cal = Calendar(self._cal_type)
        step = int(12 / annual_frequency(self._freq_type))
        eff, term = self._effective_dt, self._termination_dt
        if self._dg_type == DateGenRuleTypes.BACKWARD:
            rolls = []
            nxt, k = term, 0
            while nxt > eff:
                rolls.append(nxt)
                k += 1
                nxt = term.add_months(-step * k)
                if self._end_of_month:
                    nxt = nxt.eom()
            # rolls: termination first, then backwards; nxt is the previous coupon date
            out = [nxt] + [cal.adjust(d, self._bd_type) for d in reversed(rolls[1:])] + [term]
        elif self._dg_type == DateGenRuleTypes.FORWARD:
            rolls = [eff]
            nxt, k = eff, 1
            while nxt < term:
                rolls.append(nxt)
                nxt = eff.add_months(step * k)
                k += 1
            out = [cal.adjust(d, self._bd_type) for d in rolls[1:]] + [term]
        else:
            raise LibError("Unknown date generation rule")
        if out[0] < eff:
            out[0] = eff
        if self._adjust_termination_dt:
            self._termination_dt = cal.adjust(term, self._bd_type)
            out[-1] = self._termination_dt
        if len(out) < 2:
            raise LibError("Schedule has two dates only.")
        # the reference drops the head whenever two consecutive dates coincide
        snapshot = list(out)
        prev = snapshot[0]
        for dt in snapshot[1:]:
            if dt == prev:
                out.pop(0)
            if dt < prev:
                raise LibError("Dates are not monotonic")
            prev = dt
        self._adjusted_dts = out
        return out


def to_tenor(x):
    """Year fraction(s) -> tenor label(s), with the reference's rounding rules
    (helpers.py:201-242): <1/12 -> ceil weeks, <1 -> nearest month, else years+months."""
    def one(val: float) -> str:
        if val < 1 / 12:
            return f"{math.ceil(val * 365 / 7)}W"
        if val < 1:
            return f"{max(int(round(val * 12)), 1)}M"
        years = int(math.floor(val))
        rem = int(round((val - years) * 12))
        if rem == 12:
            years, rem = years + 1, 0
        return f"{years}Y" if rem == 0 else f"{years}Y{rem}M"
    if isinstance(x, list):
        return [one(v) for v in x]
    return one(x)


def times_from_dates(dt, value_dt: Date, day_count_type: DayCountTypes = None):
    """Date(s) -> year fraction(s) from value_dt (helpers.py:154-197)."""
    if not isinstance(value_dt, Date):
        raise LibError("Valuation date is not a Date")
    # day counts that are a plain day difference over a constant: one vectorised division on the serials
    # (same floating-point operation as DayCount.year_frac: num / den)
    den = {None: 365.0, DayCountTypes.ACT_365F: 365, DayCountTypes.ACT_360: 360, DayCountTypes.SIMPLE: 365.0}.get(day_count_type)
    dc = None if (day_count_type is None or den is not None) else DayCount(day_count_type)

    def one(d):
        return (d._n - value_dt._n) / den if dc is None else dc.year_frac(value_dt, d)[0]
    if isinstance(dt, Date):
        return one(dt)
    if isinstance(dt, list) and len(dt) and isinstance(dt[0], Date):
        import numpy as np
        if dc is None:
            return (np.array([d._n for d in dt], dtype=np.int64) - value_dt._n) / den
        return np.array([one(d) for d in dt])
    raise LibError("Discount factor must take dates.")
