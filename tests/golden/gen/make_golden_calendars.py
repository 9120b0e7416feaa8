#!/usr/bin/env python
"""Holiday-calendar goldens from the UNMODIFIED reference (/root/reference).  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_calendars.py

Writes tests/golden/ref_calendars.npz:
  easter[299]                day of year of Easter Monday, the reference's table (calendar.py:49-80), years 1901..2199
  hol_<NAME>                 np.packbits of Calendar(NAME).is_holiday(d) for every day 1-Jan-1901 .. 31-Dec-2199
  adj_in[n] (d, m, y), adj_<NAME>[n][5]   Calendar(NAME).adjust for the five BusDayAdjustTypes, as (d, m, y) packed d + 100 m + 10000 y
  abd_<NAME>[n][2]           add_business_days(+7) / (-4)
  sch_*                      schedules rolled on holiday calendars (flat dates + offsets)
Nothing here is imported by the product; the tests compare adrates_b200.holidays / dates / batch and the device flattener
with these arrays.
"""
import os
import time

import numpy as np

from cavour.utils import calendar as refcal
from cavour.utils.calendar import BusDayAdjustTypes, Calendar, CalendarTypes, DateGenRuleTypes, create_calendar_intersection
from cavour.utils.date import Date
from cavour.utils.frequency import FrequencyTypes
from cavour.utils.schedule import Schedule

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
NAMES = ["AUSTRALIA", "CANADA", "FRANCE", "GERMANY", "ITALY", "JAPAN", "NEW_ZEALAND", "NORWAY", "SWEDEN", "SWITZERLAND",
         "TARGET", "UNITED_STATES", "UNITED_KINGDOM", "WEEKEND"]


def pk(d):
    return d._d + 100 * d._m + 10000 * d._y


def main():
    t0 = time.time()
    out = {"easter": np.asarray(refcal.easterMondayDay, dtype=np.int32)}
    # every day of 1901..2199 as reference Date objects, built once
    days = []
    dt = Date(1, 1, 1901)
    end = Date(1, 1, 2200)
    while dt < end:
        days.append(dt)
        dt = dt.add_days(1)
    print(len(days), "days", time.time() - t0)
    for nm in NAMES:
        cal = Calendar(CalendarTypes[nm])
        out["hol_" + nm] = np.packbits(np.fromiter((bool(cal.is_holiday(d)) for d in days), dtype=bool, count=len(days)))
        print(nm, int(np.unpackbits(out["hol_" + nm]).sum()), time.time() - t0)

    rng = np.random.default_rng(77)
    probes = [Date(1, 1, 1990).add_days(int(k)) for k in rng.integers(0, 365 * 80, size=1500)]
    # month ends / starts and the days around Easter and Christmas are where the modified rules and the long weekends bite
    for y in (2021, 2022, 2024, 2027, 2038):
        for (d, m) in ((31, 12), (30, 12), (1, 1), (2, 1), (25, 12), (26, 12), (27, 12), (28, 12), (31, 3), (30, 4), (1, 5),
                       (31, 5), (31, 8), (30, 6)):
            probes.append(Date(d, m, y))
        em = Calendar(CalendarTypes.TARGET).easter_monday(y)
        for k in range(-5, 3):
            probes.append(em.add_days(k))
    out["adj_in"] = np.asarray([pk(d) for d in probes], dtype=np.int64)
    bds = [BusDayAdjustTypes.NONE, BusDayAdjustTypes.FOLLOWING, BusDayAdjustTypes.MODIFIED_FOLLOWING,
           BusDayAdjustTypes.PRECEDING, BusDayAdjustTypes.MODIFIED_PRECEDING]
    cals = {nm: Calendar(CalendarTypes[nm]) for nm in NAMES}
    cals["US_UK"] = create_calendar_intersection(cals["UNITED_STATES"], cals["UNITED_KINGDOM"])
    for nm, cal in cals.items():
        out["adj_" + nm] = np.asarray([[pk(cal.adjust(d, b)) for b in bds] for d in probes], dtype=np.int64)
        out["abd_" + nm] = np.asarray([[pk(cal.add_business_days(d, 7)), pk(cal.add_business_days(d, -4))]
                                       for d in probes[:400]], dtype=np.int64)
    print("adjust", time.time() - t0)

    # schedules on holiday calendars
    rows, flat, off = [], [], [0]
    effs = [(30, 4, 2024), (17, 12, 2024), (29, 2, 2024), (31, 8, 2023), (24, 12, 2025), (1, 1, 2026), (28, 3, 2024)]
    tens = ["3M", "1Y", "18M", "2Y", "5Y", "13Y", "30Y"]
    codes = {"freq": ["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"], "bd": ["MODIFIED_FOLLOWING", "FOLLOWING", "PRECEDING",
                                                                      "MODIFIED_PRECEDING"], "dg": ["BACKWARD", "FORWARD"]}
    for nm in ("UNITED_KINGDOM", "UNITED_STATES", "TARGET", "JAPAN", "SWEDEN"):
        for e in effs:
            for tn in tens:
                for fi, fq in enumerate(codes["freq"]):
                    for bi, bd in enumerate(codes["bd"]):
                        for di, dg in enumerate(codes["dg"]):
                            if dg == "FORWARD" and bd != "MODIFIED_FOLLOWING":
                                continue
                            ed = Date(*e)
                            td = ed.add_tenor(tn)
                            s = Schedule(ed, td, FrequencyTypes[fq], CalendarTypes[nm], BusDayAdjustTypes[bd],
                                         DateGenRuleTypes[dg])
                            dts = [pk(d) for d in s._adjusted_dts]
                            rows.append([NAMES.index(nm), pk(ed), pk(td), fi, bi, di])
                            flat.extend(dts)
                            off.append(len(flat))
    out["sch_rows"] = np.asarray(rows, dtype=np.int64)
    out["sch_dates"] = np.asarray(flat, dtype=np.int64)
    out["sch_off"] = np.asarray(off, dtype=np.int64)
    out["names"] = np.asarray(NAMES)
    np.savez_compressed(os.path.join(OUT, "ref_calendars.npz"), **out)
    print("done", len(rows), "schedules", time.time() - t0)


if __name__ == "__main__":
    main()
