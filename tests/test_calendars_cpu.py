"""Holiday calendars (adrates_b200.holidays rule tables -> day-serial tables) against the unmodified reference:
every day of 1901-2199 of every calendar (Calendar.is_holiday, reference calendar.py:278-1099), the Easter table
(calendar.py:49-80) against the computus, Calendar.adjust for the five roll conventions, add_business_days, an
INTERSECTION calendar, and 3 675 schedules rolled on holiday calendars (schedule.py:163-270) - through the object layer
(dates.py), the array layer (batch.py) and the device flattener's rules compiled for the host (cav_book_core.h).
Goldens: tests/golden/ref_calendars.npz, generator tests/golden/gen/make_golden_calendars.py."""
import os

import numpy as np
import pytest

from adrates_b200 import batch as B
from adrates_b200 import holidays as H
from adrates_b200.curves import OISCurve
from adrates_b200.dates import (BusDayAdjustTypes, Calendar, CalendarTypes, Date, DateGenRuleTypes, DayCountTypes,
                                FrequencyTypes, Schedule, create_calendar_intersection)
from adrates_b200.error import LibError
from adrates_b200.global_types import InterpTypes
from tests import native_book as nb
from tests.test_book_core_cpu import _random_book, book_conv9
from tests.util_trades import make_calibration_swaps

BDS = [BusDayAdjustTypes.NONE, BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.MODIFIED_FOLLOWING,
       BusDayAdjustTypes.PRECEDING, BusDayAdjustTypes.MODIFIED_PRECEDING]
FREQ = [FrequencyTypes.ANNUAL, FrequencyTypes.SEMI_ANNUAL, FrequencyTypes.QUARTERLY]
SBD = [BusDayAdjustTypes.MODIFIED_FOLLOWING, BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.PRECEDING,
       BusDayAdjustTypes.MODIFIED_PRECEDING]
SDG = [DateGenRuleTypes.BACKWARD, DateGenRuleTypes.FORWARD]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_calendars.npz"))


def _ser(packed):
    p = np.asarray(packed, dtype=np.int64)
    return B.ordinal(p % 100, (p // 100) % 100, p // 10000)


def _cal(name):
    if name == "US_UK":
        return create_calendar_intersection(Calendar(CalendarTypes.UNITED_STATES), Calendar(CalendarTypes.UNITED_KINGDOM))
    return Calendar(CalendarTypes[name])


def test_easter_table_is_the_computus(gold):
    years = np.arange(H.YEAR_LO, H.YEAR_HI + 1)
    doy = H.easter_monday_serial(years) - B.ordinal(np.ones_like(years), np.ones_like(years), years) + 1
    assert np.array_equal(doy, gold["easter"])
    assert Calendar(CalendarTypes.TARGET).easter_monday(2024) == Date(1, 4, 2024)


def test_every_day_of_every_calendar_matches_the_reference(gold):
    names = [str(n) for n in gold["names"]]
    assert len(names) == 14
    for nm in names:
        ref = np.unpackbits(gold["hol_" + nm])[:H.N_DAYS].astype(bool)
        tab = H.table(CalendarTypes[nm])
        assert np.array_equal(tab.holiday, ref), nm
        # the derived tables: non-business days, next / previous business day, packed words
        n = np.arange(H.BASE, H.BASE + H.N_DAYS)
        closed = ref | (B.weekday(n) >= 5)
        assert np.array_equal(tab.non_business, closed)
        idx = np.nonzero(~closed)[0]
        probe = np.arange(idx[0], idx[-1] + 1)
        assert np.array_equal(tab.next_bd[probe], idx[np.searchsorted(idx, probe, side="left")])
        assert np.array_equal(tab.prev_bd[probe], idx[np.searchsorted(idx, probe, side="right") - 1])
        w = tab.words()
        assert w.dtype == np.uint32 and w.shape[0] == (H.N_DAYS + 31) // 32
        assert np.array_equal(((w[probe >> 5] >> (probe & 31).astype(np.uint32)) & 1).astype(bool), closed[probe])
    # object layer spot checks (holiday on a weekend date counts as a holiday, a plain weekend day does not)
    uk = Calendar(CalendarTypes.UNITED_KINGDOM)
    assert uk.is_holiday(Date(25, 12, 2021)) and not uk.is_business_day(Date(27, 12, 2021)) and uk.is_holiday(Date(3, 6, 2022))
    assert not uk.is_holiday(Date(4, 5, 2024)) and not uk.is_business_day(Date(4, 5, 2024))
    assert uk.get_holiday_list(2024) == ["01-JAN-2024", "29-MAR-2024", "01-APR-2024", "06-MAY-2024", "27-MAY-2024",
                                         "26-AUG-2024", "25-DEC-2024", "26-DEC-2024"]
    with pytest.raises(LibError):
        uk.is_holiday(Date(1, 1, 2200))


def test_adjust_and_add_business_days_match_the_reference(gold):
    n = _ser(gold["adj_in"])
    for nm in [str(x) for x in gold["names"]] + ["US_UK"]:
        cal = _cal(nm)
        ref = _ser(gold["adj_" + nm])
        arg = cal if nm == "US_UK" else CalendarTypes[nm]
        if nm not in ("US_UK", "WEEKEND"):
            nb.set_holidays(H.table(CalendarTypes[nm]).words(), H.BASE, H.N_DAYS)
        for j, bd in enumerate(BDS):
            assert np.array_equal(B.adjust(n, bd, arg), ref[:, j]), (nm, bd)                      # array layer
            for k in range(0, n.shape[0], 5):                                                    # object layer
                assert cal.adjust(Date._of(int(n[k])), bd)._n == ref[k, j], (nm, bd, k)
            if nm != "US_UK":                                                                    # device rules on the host
                assert np.array_equal(nb.adjust(n, bd.value, CalendarTypes[nm].value), ref[:, j]), (nm, bd)
        ra = _ser(gold["abd_" + nm])
        for k in range(0, ra.shape[0], 2):
            d = Date._of(int(n[k]))
            assert cal.add_business_days(d, 7)._n == ra[k, 0] and cal.add_business_days(d, -4)._n == ra[k, 1], (nm, k)
    nb.set_holidays(None, 0, 0)


def test_schedules_on_holiday_calendars_match_the_reference(gold):
    names = [str(x) for x in gold["names"]]
    rows, off, dates = gold["sch_rows"], gold["sch_off"], _ser(gold["sch_dates"])
    eff, term = _ser(rows[:, 1]), _ser(rows[:, 2])
    keys = sorted(set(map(tuple, rows[:, [0, 3, 4, 5]])))
    assert len(rows) == 3675
    for key in keys:
        cal_type = CalendarTypes[names[key[0]]]
        sel = np.nonzero((rows[:, [0, 3, 4, 5]] == key).all(1))[0]
        sch = B.roll_schedules(eff[sel], term[sel], FREQ[key[1]], cal_type, SBD[key[2]], SDG[key[3]])
        nb.set_holidays(H.table(cal_type).words(), H.BASE, H.N_DAYS)
        step = {0: 12, 1: 6, 2: 3}[key[1]]
        for q, i in enumerate(sel):
            want = dates[off[i]:off[i + 1]]
            assert np.array_equal(sch.dates[sch.offsets[q]:sch.offsets[q + 1]], want), (key, i)
            got = nb.schedule(eff[i], term[i], step, cal_type.value, SBD[key[2]].value, SDG[key[3]].value)
            assert np.array_equal(got, want), (key, i)
        for i in sel[::7]:
            s = Schedule(Date._of(int(eff[i])), Date._of(int(term[i])), FREQ[key[1]], cal_type, SBD[key[2]], SDG[key[3]])
            assert [d._n for d in s._adjusted_dts] == list(dates[off[i]:off[i + 1]])
    nb.set_holidays(None, 0, 0)


@pytest.mark.parametrize("cal", ["UNITED_KINGDOM", "TARGET", "JAPAN"])
def test_class_units_on_a_holiday_calendar_equal_batch_flatten(ref_curves, cal):
    """Unit arrays of a random book rolled on a holiday calendar: the device flattener's rules (host build) against
    batch.OISBook.flatten, bit for bit; and the holiday calendar does change the book (it is not silently WEEKEND)."""
    cv = ref_curves["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    rng = np.random.default_rng(41)
    conv = dict(fixed_freq_type=FrequencyTypes.SEMI_ANNUAL, fixed_dc_type=DayCountTypes.ACT_365F, float_freq_type=FrequencyTypes.QUARTERLY,
                float_dc_type=DayCountTypes.ACT_360, bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
    arrays = _random_book(curve, 400, rng, spread=True)
    book = B.OISBook.from_arrays(curve, **arrays, cal_type=CalendarTypes[cal], **conv)
    flat = book.flatten(dedup=True, tiles=False)
    plain = B.OISBook.from_arrays(curve, **arrays, **conv).flatten(dedup=True, tiles=False)
    assert flat.n_terms != plain.n_terms or not np.array_equal(flat.amt, plain.amt)
    eff, term, cls_of = book.schedule_classes()
    with_spread = np.bincount(cls_of[book.spread != 0.0], minlength=eff.shape[0]) > 0
    nb.set_holidays(H.table(CalendarTypes[cal]).words(), H.BASE, H.N_DAYS)
    err, off, amt, weight, node, *_ = nb.flatten_classes(book_conv9(book), eff, term, with_spread, curve.path_b_plan().node_time, True)
    nb.set_holidays(None, 0, 0)
    assert err == 0
    assert np.array_equal(off, flat.unit_offsets) and np.array_equal(amt, flat.amt)
    assert np.array_equal(weight, flat.weight) and np.array_equal(node, flat.node)
    assert book.device_conv().cal_type == CalendarTypes[cal].value


def test_reference_intersection_known_answers():
    """The known answers of the reference's own tests/test_calendar_intersection.py (:60-262), as one table: joint US / UK (and
    US / UK / TARGET) business days, holiday flags of the single calendars, the four roll conventions around 4 July and the
    August bank holiday 2024, business-day stepping over a US holiday, constructor errors, the WEEKEND baseline."""
    us, uk, tgt = (Calendar(CalendarTypes[n]) for n in ("UNITED_STATES", "UNITED_KINGDOM", "TARGET"))
    joint = create_calendar_intersection(us, uk)
    assert joint._cal_type == CalendarTypes.INTERSECTION and joint._constituent_calendars == [us, uk]
    assert len(Calendar(CalendarTypes.INTERSECTION, [us, uk])._constituent_calendars) == 2
    with pytest.raises(LibError, match="at least 2 calendars"):
        create_calendar_intersection(us)
    with pytest.raises(LibError, match="must be Calendar objects"):
        create_calendar_intersection(us, "not a calendar")
    # (date, US holiday, UK holiday, joint business day)
    for dmy, h_us, h_uk, bd in [((5, 6, 2024), False, False, True), ((4, 7, 2024), True, False, False),
                                ((26, 8, 2024), False, True, False), ((25, 12, 2024), True, True, False),
                                ((1, 1, 2024), True, True, False), ((2, 1, 2024), False, False, True)]:
        d = Date(*dmy)
        assert (us.is_holiday(d), uk.is_holiday(d)) == (h_us, h_uk), dmy
        assert joint.is_holiday(d) == (h_us or h_uk) and joint.is_business_day(d) == bd, dmy
    assert not joint.is_business_day(Date(1, 6, 2024)) and not joint.is_business_day(Date(2, 6, 2024))     # weekend
    july4, aug26 = Date(4, 7, 2024), Date(26, 8, 2024)
    assert joint.adjust(july4, BusDayAdjustTypes.FOLLOWING) == Date(5, 7, 2024)
    assert joint.adjust(july4, BusDayAdjustTypes.PRECEDING) == Date(3, 7, 2024)
    assert joint.adjust(aug26, BusDayAdjustTypes.MODIFIED_FOLLOWING) == Date(27, 8, 2024)
    assert joint.adjust(july4, BusDayAdjustTypes.NONE) == july4
    triple = create_calendar_intersection(us, uk, tgt)
    assert len(triple._constituent_calendars) == 3
    assert triple.is_holiday(Date(1, 5, 2024)) and not triple.is_business_day(Date(1, 5, 2024))           # TARGET Labour Day
    assert joint.add_business_days(Date(3, 7, 2024), 3) == Date(9, 7, 2024)
    assert joint.add_business_days(Date(5, 7, 2024), -1) == Date(3, 7, 2024)
    assert str(joint) == "INTERSECTION"
    wk = Calendar(CalendarTypes.WEEKEND)
    assert not wk.is_business_day(Date(1, 6, 2024)) and not wk.is_business_day(Date(2, 6, 2024)) and wk.is_business_day(Date(3, 6, 2024))
    assert not wk.is_holiday(Date(25, 12, 2024))
    # the array layer takes the intersection as a Calendar object
    n = np.array([july4._n, aug26._n, Date(5, 6, 2024)._n])
    assert np.array_equal(B.adjust(n, BusDayAdjustTypes.FOLLOWING, joint), [Date(5, 7, 2024)._n, Date(27, 8, 2024)._n, n[2]])
