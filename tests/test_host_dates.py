"""Host date/schedule/day-count logic against vectors produced by the reference's own
Schedule / DayCount / Date / to_tenor (tests/golden/gen/make_golden.py)."""
from adrates_b200.dates import (Date, Calendar, CalendarTypes, BusDayAdjustTypes, DateGenRuleTypes, DayCount,
                                DayCountTypes, FrequencyTypes, Schedule, to_tenor)
from adrates_b200.error import LibError


def dmy(dt):
    return [dt.d(), dt.m(), dt.y()]


def test_schedules_match_reference(ref_schedules):
    n = 0
    for row in ref_schedules["schedules"]:
        eff = Date(*row["eff"])
        term = eff.add_tenor(row["tenor"])
        assert dmy(term) == row["term"], row
        try:
            s = Schedule(eff, term, FrequencyTypes[row["freq"]], CalendarTypes.WEEKEND, BusDayAdjustTypes[row["bd"]],
                         DateGenRuleTypes[row["dg"]])
            got = [dmy(d) for d in s._adjusted_dts]
        except LibError as ex:
            got = "ERR:LibError"
        assert got == row["dates"], row
        n += 1
    assert n == len(ref_schedules["schedules"]) and n > 1500


def test_daycounts_and_date_math_match_reference(ref_schedules):
    cal = Calendar(CalendarTypes.WEEKEND)
    for row in ref_schedules["daycounts"]:
        d1, d2 = Date(*row["d1"]), Date(*row["d2"])
        assert d2 - d1 == row["serial_diff"]
        assert d1.weekday() == row["wd1"]
        for dc in ["ACT_365F", "ACT_360", "THIRTY_E_360", "THIRTY_360_BOND", "THIRTY_E_360_ISDA", "ACT_ACT_ISDA",
                   "THIRTY_E_PLUS_360", "SIMPLE"]:
            assert DayCount(DayCountTypes[dc]).year_frac(d1, d2)[0] == row[dc], (dc, row)
        assert dmy(d1.add_weekdays(5)) == row["add_wd_5"]
        assert dmy(d1.add_weekdays(-3)) == row["add_wd_m3"]
        assert dmy(d1.add_months(-7)) == row["add_months_m7"]
        assert dmy(cal.adjust(d1, BusDayAdjustTypes.MODIFIED_FOLLOWING)) == row["adj_mf"]
        assert dmy(cal.adjust(d1, BusDayAdjustTypes.MODIFIED_PRECEDING)) == row["adj_mp"]


def test_to_tenor_matches_reference(ref_schedules):
    assert to_tenor(ref_schedules["to_tenor_in"]) == ref_schedules["to_tenor_out"]


def test_date_repr_formats_are_the_references():
    """repr / str of Date in every DateFormatTypes member (reference date.py:908-1008; the default is UK_LONG: error messages
    and Cashflows.to_dict() carry dates in it).  Known answers printed by the unmodified reference's Date in this container."""
    from adrates_b200 import DateFormatTypes, set_date_format
    dates = [(30, 4, 2024), (1, 1, 2000), (29, 2, 2028), (5, 11, 1999), (31, 12, 2199)]
    ref = {"BLOOMBERG": ["04/30/24", "01/01/00", "02/29/28", "11/05/99", "12/31/99"],
           "US_SHORT": ["04-30-24", "01-01-00", "02-29-28", "11-05-99", "12-31-99"],
           "US_MEDIUM": ["04-30-2024", "01-01-2000", "02-29-2028", "11-05-1999", "12-31-2199"],
           "US_LONG": ["APR-30-2024", "JAN-01-2000", "FEB-29-2028", "NOV-05-1999", "DEC-31-2199"],
           "US_LONGEST": ["TUE APR 30 2024", "SAT JAN 01 2000", "TUE FEB 29 2028", "FRI NOV 05 1999", "TUE DEC 31 2199"],
           "UK_SHORT": ["30/04/24", "01/01/00", "29/02/28", "05/11/99", "31/12/99"],
           "UK_MEDIUM": ["30/04/2024", "01/01/2000", "29/02/2028", "05/11/1999", "31/12/2199"],
           "UK_LONG": ["30-APR-2024", "01-JAN-2000", "29-FEB-2028", "05-NOV-1999", "31-DEC-2199"],
           "UK_LONGEST": ["TUE 30 APR 2024", "SAT 01 JAN 2000", "TUE 29 FEB 2028", "FRI 05 NOV 1999", "TUE 31 DEC 2199"],
           "DATETIME": ["30/04/2024 00:00:00", "01/01/2000 00:00:00", "29/02/2028 00:00:00", "05/11/1999 00:00:00",
                        "31/12/2199 00:00:00"]}
    assert repr(Date(30, 4, 2024)) == "30-APR-2024"                      # the default
    try:
        for f in DateFormatTypes:
            set_date_format(f)
            assert [repr(Date(*d)) for d in dates] == ref[f.name], f
            assert [str(Date(*d)) for d in dates] == ref[f.name], f
    finally:
        set_date_format(DateFormatTypes.UK_LONG)
    assert [Date(*d).str() for d in dates] == ["30APR2024", "01JAN2000", "29FEB2028", "05NOV1999", "31DEC2199"]
