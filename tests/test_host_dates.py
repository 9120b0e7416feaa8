"""Host date/schedule/day-count logic against vectors produced by the reference's own
Schedule / DayCount / Date / to_tenor (tests/golden/gen/make_golden.py)."""
import pytest

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


def test_quarterly_roll_dates_and_collateral_names_are_the_references():
    """next_cds_date / next_imm_date / third_wednesday_of_month / daily_working_day_schedule / from_datetime (reference
    date.py:698-795, 1024-1056) and the collateral helpers (global_types.py:157-300): known answers printed by the unmodified
    reference in this container (the date rules were also compared on 4 000 random dates there)."""
    import datetime
    from adrates_b200.dates import daily_working_day_schedule, from_datetime
    from adrates_b200.global_types import (CollateralType, CurrencyTypes, collateral_to_currency, get_discount_curve_name,
                                           is_bond_collateral, is_currency_collateral)
    cases = [(19, 3, 2024), (20, 3, 2024), (21, 12, 2024), (15, 6, 2025), (18, 6, 2025), (17, 9, 2025), (16, 9, 2025), (31, 12, 2030),
             (1, 1, 2031), (30, 4, 2024)]
    ref = [((20, 3, 2024), (20, 3, 2024)), ((20, 6, 2024), (19, 6, 2024)), ((20, 3, 2025), (19, 3, 2025)), ((20, 6, 2025), (18, 6, 2025)),
           ((20, 6, 2025), (17, 9, 2025)), ((20, 9, 2025), (17, 12, 2025)), ((20, 9, 2025), (17, 9, 2025)), ((20, 3, 2031), (19, 3, 2031)),
           ((20, 3, 2031), (19, 3, 2031)), ((20, 6, 2024), (19, 6, 2024))]
    dmy = lambda x: (x.d(), x.m(), x.y())  # noqa: E731
    assert [(dmy(Date(*c).next_cds_date()), dmy(Date(*c).next_imm_date())) for c in cases] == ref
    assert [dmy(Date(30, 4, 2024).next_cds_date(mm)) for mm in (1, 2, 8, -5)] == [(20, 6, 2024), (20, 9, 2024), (20, 3, 2025), (20, 12, 2023)]
    assert Date(1, 1, 2024).third_wednesday_of_month(5, 2024) == 15 and Date(1, 1, 2024).third_wednesday_of_month(8, 2024) == 21
    days = daily_working_day_schedule(Date(27, 12, 2023), Date(8, 1, 2024))
    assert len(days) == 9 and all(not d.is_weekend() for d in days) and days[-1] == Date(8, 1, 2024)
    assert from_datetime(datetime.datetime(2024, 4, 30, 12, 0)) == Date(30, 4, 2024)
    names = [[get_discount_curve_name(CurrencyTypes[c], CollateralType[k]) for k in ("USD", "GBP", "EUR", "GBP_GILTS")]
             for c in ("GBP", "USD", "JPY")]
    assert names == [["GBP_USD_XCCY", "GBP_OIS_SONIA", "GBP_EUR_XCCY", "GBP_GBP_GILTS_XCCY"],
                     ["USD_OIS_SOFR", "USD_GBP_XCCY", "USD_EUR_XCCY", "USD_GBP_GILTS_XCCY"],
                     ["JPY_USD_XCCY", "JPY_GBP_XCCY", "JPY_EUR_XCCY", "JPY_GBP_GILTS_XCCY"]]
    assert [k.name for k in CollateralType if is_currency_collateral(k)] == ["USD", "GBP", "EUR", "JPY", "CHF", "AUD", "CAD"]
    assert [k.name for k in CollateralType if is_bond_collateral(k)] == ["USD_TIPS", "EUR_OATS", "EUR_BUNDS", "GBP_GILTS", "JGB"]
    assert collateral_to_currency(CollateralType.JGB) == CurrencyTypes.JPY
    for bad in (lambda: collateral_to_currency(CollateralType.UNCOLLATERALIZED),
                lambda: get_discount_curve_name(CurrencyTypes.GBP, CollateralType.UNCOLLATERALIZED),
                lambda: get_discount_curve_name(CurrencyTypes.NZD, CollateralType.USD) and get_discount_curve_name(CurrencyTypes.SEK, CollateralType.UNCOLLATERALIZED)):
        with pytest.raises(ValueError):
            bad()
