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
