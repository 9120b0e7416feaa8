#!/usr/bin/env python
"""Differential run of the host layer against the UNMODIFIED reference on random inputs: the same call on the reference's object
and on ours, results compared exactly (dates, schedules, day counts) or to 1e-12 of the notional (values), exceptions by type.
Build container only (imports /root/reference under the torch-backed jax stand-in of tests/golden/gen/refshim); TEST
INFRASTRUCTURE, nothing in tests/, smoke() or bench.py uses it.

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tools/reftests/differential_fuzz.py [seed]

Covers Date arithmetic, DayCount.year_frac (all types, with and without the ICMA arguments), Schedule (all frequencies, rules,
adjustments, holiday calendars), Calendar.adjust / add_business_days, Bond / FRN analytics, OIS / leg host values, the path-A OIS
bootstrap, the basis-curve bootstrap with its spread Jacobian, inflation index look-ups / the inflation curve / the non-AD
values of zero-coupon and year-on-year inflation swaps.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from cavour.utils.date import Date as RDate                                                              # noqa: E402
from cavour.utils.day_count import DayCount as RDayCount, DayCountTypes as RDC                            # noqa: E402
from cavour.utils.frequency import FrequencyTypes as RFreq                                                # noqa: E402
from cavour.utils.schedule import Schedule as RSchedule                                                   # noqa: E402
from cavour.utils.calendar import (CalendarTypes as RCal, BusDayAdjustTypes as RBd, DateGenRuleTypes as RDg,  # noqa: E402
                                   Calendar as RCalendar)
from cavour.utils.currency import CurrencyTypes as RCcy                                                   # noqa: E402
from cavour.utils.global_types import CurveTypes as RCurve, SwapTypes as RSwap                            # noqa: E402
from cavour.market.curves.interpolator import InterpTypes as RInterp                                      # noqa: E402
from cavour.market.curves.discount_curve import DiscountCurve as RDiscountCurve                           # noqa: E402
from cavour.trades.credit.bond import Bond as RBond                                                       # noqa: E402
from cavour.trades.credit.frn import FRN as RFRN                                                          # noqa: E402
from cavour.trades.rates.ois import OIS as ROIS                                                           # noqa: E402
from cavour.trades.rates.swap_float_leg import SwapFloatLeg as RFloatLeg                                  # noqa: E402
from cavour.models.models import Model as RModel                                                          # noqa: E402
from cavour.utils.global_types import InflationIndexTypes as RIdxT, InflationInterpTypes as RIdxI         # noqa: E402
from cavour.market.indices.inflation_index import InflationIndex as RIndex                                # noqa: E402
from cavour.market.curves.inflation_curve import InflationCurve as RInflCurve                             # noqa: E402
from cavour.trades.rates.zcis import ZeroCouponInflationSwap as RZcis                                     # noqa: E402
from cavour.trades.rates.yoy_inflation_swap import YoYInflationSwap as RYoY                               # noqa: E402

import adrates_b200 as O                                                                                  # noqa: E402

VD = (30, 4, 2024)
OFFSETS = [0.25, 0.5, 1, 2, 3, 5, 7, 10, 15, 20, 30, 40]
ZEROS = np.array([0.052, 0.051, 0.048, 0.044, 0.042, 0.04, 0.039, 0.0385, 0.039, 0.039, 0.038, 0.037])
mismatch, worst, calls = {}, {}, {}


def dmy(x):
    return (x._d, x._m, x._y)


def plain(v):
    if isinstance(v, list):
        return [plain(x) for x in v]
    if isinstance(v, tuple):
        return tuple(plain(x) for x in v)
    return dmy(v) if hasattr(v, "_d") else v


def run(f):
    try:
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its arguments before raising
            return plain(f())
    except Exception as ex:  # noqa: BLE001
        return "raises " + type(ex).__name__


def exact(key, fr, fo):
    calls[key] = calls.get(key, 0) + 1
    a, b = run(fr), run(fo)
    if a != b:
        mismatch.setdefault(key, []).append((a, b))


def close(key, fr, fo, scale):
    calls[key] = calls.get(key, 0) + 1
    a, b = run(fr), run(fo)
    if isinstance(a, str) or isinstance(b, str):
        if a != b and not (isinstance(a, float) and np.isnan(a)) and not (isinstance(b, float) and np.isnan(b)):
            mismatch.setdefault(key, []).append((a, b))
        return
    a, b = float(a), float(b)
    if np.isnan(a) and np.isnan(b):
        return
    worst[key] = max(worst.get(key, 0.0), abs(a - b) / scale)


def random_date(rng, lo=1950, hi=2080):
    while True:
        y, m, d = rng.randint(lo, hi), rng.randint(1, 12), rng.randint(1, 31)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                return RDate(d, m, y), O.Date(d, m, y)
        except Exception:  # noqa: BLE001
            pass


def dates(rng, n):
    tenors = ["1D", "3D", "1W", "2W", "1M", "3M", "6M", "9M", "12M", "18M", "1Y", "2Y", "5Y", "10Y", "30Y", "-1M", "-3M", "-1Y", "-2W",
              "ON", "TN", "1d", "6m", "2y", "0M", "0D"]
    for _ in range(n):
        r, o = random_date(rng)
        t, k = rng.choice(tenors), rng.randint(-40, 40)
        yrs = rng.choice([0.5, 1, 1.0, 2.25, 10, 0.0833333, 7.5, -1, -0.5, 30])
        exact("Date.add_tenor", lambda: r.add_tenor(t), lambda: o.add_tenor(t))
        exact("Date.add_tenor(list)", lambda: r.add_tenor(["1M", "1Y"]), lambda: o.add_tenor(["1M", "1Y"]))
        exact("Date.add_months", lambda: r.add_months(k), lambda: o.add_months(k))
        exact("Date.add_days", lambda: r.add_days(13 * k), lambda: o.add_days(13 * k))
        exact("Date.add_weekdays", lambda: r.add_weekdays(k), lambda: o.add_weekdays(k))
        exact("Date.add_years", lambda: r.add_years(yrs), lambda: o.add_years(yrs))
        exact("Date.add_years(list)", lambda: r.add_years([0.5, 2.0]), lambda: o.add_years([0.5, 2.0]))
        exact("Date.eom / is_eom / weekday / excel_dt", lambda: (r.eom(), r.is_eom(), r.weekday(), r.is_weekend(), r.excel_dt()),
              lambda: (o.eom(), o.is_eom(), o.weekday(), o.is_weekend(), o.excel_dt()))
        exact("Date.next_cds_date / next_imm_date", lambda: (r.next_cds_date(k % 7), r.next_imm_date()),
              lambda: (o.next_cds_date(k % 7), o.next_imm_date()))


def day_counts_and_schedules(rng, n):
    for _ in range(n):
        (r1, o1), (r2, o2), (r3, o3) = (random_date(rng, 1990, 2070) for _ in range(3))
        if r1 > r2:
            r1, r2, o1, o2 = r2, r1, o2, o1
        fq, term = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY", "MONTHLY"]), rng.random() < 0.3
        for t in RDC:
            exact("DayCount.year_frac(d1, d2, d3, freq, term) " + t.name,
                  lambda: tuple(RDayCount(t).year_frac(r1, r2, r3, RFreq[fq], term)),
                  lambda: tuple(O.DayCount(O.DayCountTypes[t.name]).year_frac(o1, o2, o3, O.FrequencyTypes[fq], term)))
            if t.name != "ACT_365L":         # without the third date the reference trips over None there; this package answers
                exact("DayCount.year_frac(d1, d2) " + t.name, lambda: tuple(RDayCount(t).year_frac(r1, r2)),
                      lambda: tuple(O.DayCount(O.DayCountTypes[t.name]).year_frac(o1, o2)))
    for _ in range(n):
        r1, o1 = random_date(rng, 1990, 2070)
        span = 365 * rng.choice([1, 2, 3, 5, 7]) + rng.randint(0, 400)
        r2, o2 = r1.add_days(span), o1.add_days(span)
        fq = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY", "MONTHLY", "TRI_ANNUAL"])
        cal = rng.choice(["WEEKEND", "NONE", "UNITED_KINGDOM", "TARGET", "UNITED_STATES", "JAPAN", "GERMANY", "AUSTRALIA"])
        bd, dg = rng.choice(list(RBd)).name, rng.choice(list(RDg)).name
        adj, eom, k = rng.random() < 0.5, rng.random() < 0.3, rng.randint(-30, 30)
        exact("Schedule._adjusted_dts", lambda: RSchedule(r1, r2, RFreq[fq], RCal[cal], RBd[bd], RDg[dg], adj, eom)._adjusted_dts,
              lambda: O.Schedule(o1, o2, O.FrequencyTypes[fq], O.CalendarTypes[cal], O.BusDayAdjustTypes[bd], O.DateGenRuleTypes[dg],
                                 adj, eom)._adjusted_dts)
        exact("Calendar.add_business_days", lambda: RCalendar(RCal[cal]).add_business_days(r1, k),
              lambda: O.Calendar(O.CalendarTypes[cal]).add_business_days(o1, k))
        exact("Calendar.adjust", lambda: RCalendar(RCal[cal]).adjust(r1, RBd[bd]), lambda: O.Calendar(O.CalendarTypes[cal]).adjust(o1, O.BusDayAdjustTypes[bd]))


def curves(scheme, power=1.0):
    dfs = np.exp(-np.array(OFFSETS) * ZEROS) ** power
    return RDiscountCurve(RDate(*VD), OFFSETS, dfs, RInterp[scheme]), O.DiscountCurve(O.Date(*VD), OFFSETS, dfs, O.InterpTypes[scheme])


def credit(rng, n):
    rvd, ovd = RDate(*VD), O.Date(*VD)
    for scheme in ("FLAT_FWD_RATES", "LINEAR_ZERO_RATES"):
        rc, oc = curves(scheme)
        for _ in range(n):
            issue = (rng.randint(1, 28), rng.randint(1, 12), rng.randint(2019, 2025))
            ten, cpn = rng.choice(["2Y", "3Y", "5Y", "7Y", "10Y", "15Y", "30Y"]), rng.choice([0.0, 0.02, 0.035, 0.05])
            fq, dc = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"]), rng.choice(["ACT_365F", "ACT_360", "THIRTY_E_360", "ACT_ACT_ISDA"])
            lag, face, z, sd = rng.choice([0, 0, 2]), rng.choice([100.0, 1e6]), rng.choice([0.0, 0.004]), rng.choice([0, 3, 40])
            rb = RBond(RDate(*issue), ten, cpn, RFreq[fq], RDC[dc], RCcy.GBP, face_value=face, payment_lag=lag)
            ob = O.Bond(O.Date(*issue), ten, cpn, O.FrequencyTypes[fq], O.DayCountTypes[dc], O.CurrencyTypes.GBP, face_value=face, payment_lag=lag)
            rs, os_ = rvd.add_days(sd), ovd.add_days(sd)
            close("Bond.value", lambda: rb.value(rvd, rc, z, rs), lambda: ob.value(ovd, oc, z, os_), face)
            close("Bond.accrued_interest", lambda: rb.accrued_interest(rs), lambda: ob.accrued_interest(os_), face)
            close("Bond.clean_price", lambda: rb.clean_price(rvd, rc, z, rs), lambda: ob.clean_price(ovd, oc, z, os_), 100.0)
            close("Bond.dv01", lambda: rb.dv01(rs, rc, z), lambda: ob.dv01(os_, oc, z), face)
            clean = run(lambda: rb.clean_price(rvd, rc, z, rs))
            if not isinstance(clean, str):
                close("Bond.yield_to_maturity", lambda: rb.yield_to_maturity(rs, clean), lambda: ob.yield_to_maturity(os_, clean), 1e3)
                close("Bond.z_spread", lambda: rb.z_spread(rs, rc, clean - 1.0), lambda: ob.z_spread(os_, oc, clean - 1.0), 1e3)
            close("Bond.duration", lambda: rb.duration(rs, rc), lambda: ob.duration(os_, oc), 1e4)
            close("Bond.convexity", lambda: rb.convexity(rs, rc), lambda: ob.convexity(os_, oc), 1e5)
            margin, fixing = rng.choice([0.0, 0.003]), rng.choice([None, 0.05])
            rf = RFRN(RDate(*issue), ten, margin, RFreq[fq], RDC[dc], RCcy.GBP, RCurve.GBP_OIS_SONIA, face_value=face, payment_lag=lag,
                      first_fixing_rate=fixing)
            of = O.FRN(O.Date(*issue), ten, margin, O.FrequencyTypes[fq], O.DayCountTypes[dc], O.CurrencyTypes.GBP, O.CurveTypes.GBP_OIS_SONIA,
                       face_value=face, payment_lag=lag, first_fixing_rate=fixing)
            close("FRN.value", lambda: rf.value(rvd, rc, rc, 0.001, rs), lambda: of.value(ovd, oc, oc, 0.001, os_), face)
            close("FRN.accrued_interest", lambda: rf.accrued_interest(rs), lambda: of.accrued_interest(os_), 100)
            close("FRN.clean_price", lambda: rf.clean_price(rvd, rc, rc, 0.0, rs), lambda: of.clean_price(ovd, oc, oc, 0.0, os_), 100)
            close("FRN.modified_duration", lambda: rf.modified_duration(rvd, rc, rc, 0.0, rs), lambda: of.modified_duration(ovd, oc, oc, 0.0, os_), 1e4)


def swaps(rng, n):
    rvd, ovd = RDate(*VD), O.Date(*VD)
    for scheme in ("FLAT_FWD_RATES", "LINEAR_ZERO_RATES", "LINEAR_FWD_RATES"):
        (rc, oc), (rc2, oc2) = curves(scheme), curves(scheme, 1.05)
        for _ in range(n):
            eff = (rng.randint(1, 28), rng.randint(1, 12), rng.choice([2022, 2023, 2024, 2024, 2024, 2025]))
            ten, side, cpn = rng.choice(["6M", "1Y", "2Y", "5Y", "10Y", "30Y"]), rng.choice(["PAY", "RECEIVE"]), rng.choice([0.01, 0.04, 0.06])
            ff, lf = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"]), rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"])
            dc, lag, spr, N = rng.choice(["ACT_365F", "ACT_360", "THIRTY_E_360"]), rng.choice([0, 0, 2]), rng.choice([0.0, 0.0015]), rng.choice([1e6, 2.5e7])
            cal, bd, dg = rng.choice(["WEEKEND", "UNITED_KINGDOM", "TARGET"]), rng.choice(["FOLLOWING", "MODIFIED_FOLLOWING", "PRECEDING"]), rng.choice(["BACKWARD", "FORWARD"])
            fix = rng.choice([None, 0.05])
            r = ROIS(RDate(*eff), ten, RSwap[side], cpn, RFreq[ff], RDC[dc], RCurve.GBP_OIS_SONIA, RCcy.GBP, N, lag, spr, RFreq[lf], RDC[dc],
                     RCal[cal], RBd[bd], RDg[dg])
            o = O.OIS(O.Date(*eff), ten, O.SwapTypes[side], cpn, O.FrequencyTypes[ff], O.DayCountTypes[dc], O.CurveTypes.GBP_OIS_SONIA,
                      O.CurrencyTypes.GBP, N, lag, spr, O.FrequencyTypes[lf], O.DayCountTypes[dc], O.CalendarTypes[cal], O.BusDayAdjustTypes[bd],
                      O.DateGenRuleTypes[dg])
            close("OIS.value", lambda: r.value(rvd, rc, first_fixing_rate=fix), lambda: o.value(ovd, oc, first_fixing_rate=fix), N)
            close("OIS.value(two curves)", lambda: r.value(rvd, rc, rc2), lambda: o.value(ovd, oc, oc2), N)
            close("OIS.pv01", lambda: r.pv01(rvd, rc), lambda: o.pv01(ovd, oc), N)
            close("OIS.swap_rate", lambda: r.swap_rate(rvd, rc), lambda: o.swap_rate(ovd, oc), 1.0)
            rl = RFloatLeg(RDate(*eff), ten, RSwap[side], spr, RFreq[lf], RDC[dc], RCurve.GBP_OIS_SONIA, RCcy.GBP, N, 0.0, lag, RCal[cal],
                           RBd[bd], RDg[dg], False, True)
            ol = O.SwapFloatLeg(O.Date(*eff), ten, O.SwapTypes[side], spr, O.FrequencyTypes[lf], O.DayCountTypes[dc], O.CurveTypes.GBP_OIS_SONIA,
                                O.CurrencyTypes.GBP, N, 0.0, lag, O.CalendarTypes[cal], O.BusDayAdjustTypes[bd], O.DateGenRuleTypes[dg], False, True)
            close("SwapFloatLeg.value(notional exchange)", lambda: rl.value(rvd, rc2, rc, fix), lambda: ol.value(ovd, oc2, oc, fix), N)


def _model(M, D, DCt, F, B, S, I, vd, quotes, interp, freq="ANNUAL", bd="MODIFIED_FOLLOWING", lag=0):
    m = M(D(*vd))
    for name, tenors, px, dc in quotes:
        m.build_curve(name=name, px_list=px, tenor_list=tenors, spot_days=0, swap_type=S.PAY, fixed_dcc_type=DCt[dc], fixed_freq_type=F[freq],
                      float_freq_type=F[freq], float_dc_type=DCt[dc], bus_day_type=B[bd], interp_type=I[interp], payment_lag=lag)
    return m


def _both_models(*args, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        r = _model(RModel, RDate, RDC, RFreq, RBd, RSwap, RInterp, *args, **kw)
    return r, _model(O.Model, O.Date, O.DayCountTypes, O.FrequencyTypes, O.BusDayAdjustTypes, O.SwapTypes, O.InterpTypes, *args, **kw)


def _nodes(c):
    return np.asarray(c._times, dtype=np.float64), np.asarray(c._dfs, dtype=np.float64)


def bootstraps(rng, n_ois, n_xccy):
    """Path-A OIS curves from random quote sets (Model.build_curve) and basis curves as Model.build_xccy_curve builds them
    (use_ad=True): node times exactly, discount factors and d DF / d spread to 1e-12."""
    all_tenors = ["1D", "1W", "2W", "1M", "2M", "3M", "6M", "9M", "1Y", "18M", "2Y", "3Y", "4Y", "5Y", "7Y", "10Y", "12Y", "15Y", "20Y", "30Y",
                  "40Y", "50Y"]
    for _ in range(n_ois):
        vd = (rng.randint(1, 28), rng.randint(1, 12), rng.randint(2020, 2026))
        tenors = sorted(rng.sample(all_tenors, rng.randint(6, len(all_tenors))), key=all_tenors.index)
        base = rng.uniform(1.0, 6.0)
        px = [round(base + rng.uniform(-0.6, 0.6) - 0.02 * i, 4) for i in range(len(tenors))]
        kw = dict(interp=rng.choice(["LINEAR_ZERO_RATES", "FLAT_FWD_RATES"]), freq=rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"]),
                  bd=rng.choice(["MODIFIED_FOLLOWING", "FOLLOWING"]), lag=rng.choice([0, 0, 2]))
        quotes = [("GBP_OIS_SONIA", tenors, px, rng.choice(["ACT_365F", "ACT_360"]))]
        calls["OISCurve (Model.build_curve)"] = calls.get("OISCurve (Model.build_curve)", 0) + 1
        try:
            rm, om = None, None
            with contextlib.redirect_stdout(io.StringIO()):
                rm = _model(RModel, RDate, RDC, RFreq, RBd, RSwap, RInterp, vd, quotes, **kw)
        except Exception as ex:  # noqa: BLE001  (quote sets the reference cannot bootstrap: ours must fail the same way)
            theirs = type(ex).__name__
            mine = run(lambda: _model(O.Model, O.Date, O.DayCountTypes, O.FrequencyTypes, O.BusDayAdjustTypes, O.SwapTypes, O.InterpTypes, vd, quotes, **kw))
            if mine != "raises " + theirs:
                mismatch.setdefault("OISCurve (Model.build_curve) errors", []).append((theirs, mine))
            continue
        om = _model(O.Model, O.Date, O.DayCountTypes, O.FrequencyTypes, O.BusDayAdjustTypes, O.SwapTypes, O.InterpTypes, vd, quotes, **kw)
        rc, oc = rm.curves.GBP_OIS_SONIA, om.curves.GBP_OIS_SONIA
        (rt, rd), (ot, od) = _nodes(rc), _nodes(oc)
        same = rt.shape == ot.shape and np.array_equal(rt, ot) and list(rc.swap_times) == list(oc.swap_times) and \
            [list(map(float, x)) for x in rc.year_fracs] == [list(map(float, x)) for x in oc.year_fracs]
        if not same:
            mismatch.setdefault("OISCurve node times / swap_times / year_fracs", []).append((vd, tenors))
            continue
        worst["OISCurve._dfs"] = max(worst.get("OISCurve._dfs", 0.0), float(np.max(np.abs(rd - od))))
    pillars = ["1Y", "2Y", "3Y", "4Y", "5Y", "7Y", "10Y", "15Y"]
    ten = ["1Y", "2Y", "3Y", "5Y", "7Y", "10Y", "15Y", "20Y", "30Y"]
    for _ in range(n_xccy):
        vd = (rng.randint(1, 28), rng.randint(1, 12), rng.randint(2021, 2025))
        gbp = [round(4.6 - 0.03 * i + rng.uniform(-0.1, 0.1), 4) for i in range(len(ten))]
        usd = [round(5.1 - 0.04 * i + rng.uniform(-0.1, 0.1), 4) for i in range(len(ten))]
        rm, om = _both_models(vd, [("GBP_OIS_SONIA", ten, gbp, "ACT_365F"), ("USD_OIS_SOFR", ten, usd, "ACT_360")],
                              interp=rng.choice(["LINEAR_ZERO_RATES", "FLAT_FWD_RATES"]))
        for c in (rm.curves.GBP_OIS_SONIA, rm.curves.USD_OIS_SOFR):      # the reference's non-AD look-ups need numpy node arrays
            c._times, c._dfs = _nodes(c)                                 # (the jax stand-in's arrays have no integer `.size`)
        pill = sorted(rng.sample(pillars, rng.randint(2, 7)), key=pillars.index)
        kw = dict(name="GBP_USD_BASIS", domestic_curve_name="USD_OIS_SOFR", foreign_curve_name="GBP_OIS_SONIA",
                  basis_spreads=[round(rng.uniform(-30, 40), 2) for _ in pill], tenor_list=pill, spot_fx=round(rng.uniform(1.0, 1.5), 4))
        dfq, ffq = rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"]), rng.choice(["ANNUAL", "SEMI_ANNUAL", "QUARTERLY"])
        with contextlib.redirect_stdout(io.StringIO()):
            rm.build_xccy_curve(domestic_freq_type=RFreq[dfq], foreign_freq_type=RFreq[ffq], **kw)
        om.build_xccy_curve(domestic_freq_type=O.FrequencyTypes[dfq], foreign_freq_type=O.FrequencyTypes[ffq], **kw)
        rx, ox = rm.curves.GBP_USD_BASIS, om.curves.GBP_USD_BASIS
        (rt, rd), (ot, od) = _nodes(rx), _nodes(ox)
        calls["XccyCurve (Model.build_xccy_curve)"] = calls.get("XccyCurve (Model.build_xccy_curve)", 0) + 1
        if rt.shape != ot.shape or not np.array_equal(rt, ot):
            mismatch.setdefault("XccyCurve node times", []).append((vd, pill, dfq, ffq))
            continue
        rj = np.asarray(rx._jac_basis, dtype=np.float64)
        worst["XccyCurve._dfs"] = max(worst.get("XccyCurve._dfs", 0.0), float(np.max(np.abs(rd - od))))
        worst["XccyCurve._jac_basis"] = max(worst.get("XccyCurve._jac_basis", 0.0), float(np.max(np.abs(rj - ox._jac_basis)) / np.max(np.abs(rj))))


def inflation(rng, n):
    """Index look-ups (lag, FLAT / LINEAR / COMPOUND between monthly fixings, seasonality, curve projection), the inflation
    curve from random ZCIS quotes, and the non-AD host values of zero-coupon and year-on-year swaps."""
    rvd, ovd = RDate(*VD), O.Date(*VD)
    rc, oc = curves("LINEAR_ZERO_RATES")
    for _ in range(n):
        lag, interp = rng.choice([2, 3]), rng.choice(["FLAT", "LINEAR", "COMPOUND"])
        season = None if rng.random() < 0.5 else {m: round(1.0 + rng.uniform(-0.006, 0.006), 4) for m in range(1, 13)}
        level, fix = 280.0, []
        for k in range(18):                                   # monthly fixings Nov-2022 .. Apr-2024
            level *= 1.0 + rng.uniform(-0.001, 0.006)
            fix.append(((1, (10 + k) % 12 + 1, 2022 + (10 + k) // 12), round(level, 2)))
        ri = RIndex(RIdxT.UK_RPI, RDate(*fix[0][0]), fix[0][1], RCcy.GBP, lag_months=lag, interp_type=RIdxI[interp], seasonality_factors=season)
        oi = O.InflationIndex(O.InflationIndexTypes.UK_RPI, O.Date(*fix[0][0]), fix[0][1], O.CurrencyTypes.GBP, lag_months=lag,
                              interp_type=O.InflationInterpTypes[interp], seasonality_factors=season)
        for d, v in fix:
            ri.add_fixing(RDate(*d), v)
            oi.add_fixing(O.Date(*d), v)
        quotes = [(t, round(0.03 + rng.uniform(-0.004, 0.006), 5)) for t in ("1Y", "2Y", "3Y", "5Y", "7Y", "10Y", "15Y", "20Y", "30Y")]
        rz = [RZcis(rvd, t, RSwap.PAY, r, ri, 1_000_000) for t, r in quotes]
        oz = [O.ZeroCouponInflationSwap(ovd, t, O.SwapTypes.PAY, r, oi, 1_000_000) for t, r in quotes]
        curve_interp = rng.choice(["FLAT", "LINEAR"])
        ric = RInflCurve(rvd, rz, fix[-1][1], RCcy.GBP, RIdxT.UK_RPI, discount_curve=rc, interp_type=RIdxI[curve_interp])
        oic = O.InflationCurve(ovd, oz, fix[-1][1], O.CurrencyTypes.GBP, O.InflationIndexTypes.UK_RPI, discount_curve=oc,
                               interp_type=O.InflationInterpTypes[curve_interp])
        (rt, rd), (ot, od) = _nodes(ric), _nodes(oic)
        calls["InflationCurve nodes"] = calls.get("InflationCurve nodes", 0) + 1
        if rt.shape != ot.shape or not np.array_equal(rt, ot):
            mismatch.setdefault("InflationCurve node times", []).append((lag, interp))
            continue
        worst["InflationCurve._dfs"] = max(worst.get("InflationCurve._dfs", 0.0), float(np.max(np.abs(rd - od))))
        ri.set_inflation_curve(ric)
        oi.set_inflation_curve(oic)
        for _ in range(12):
            y, m, d = rng.choice([2023, 2024, 2025, 2031, 2050]), rng.randint(1, 12), rng.randint(1, 28)
            lagged = rng.random() < 0.7
            close("InflationIndex.get_index", lambda: ri.get_index(RDate(d, m, y), apply_lag=lagged), lambda: oi.get_index(O.Date(d, m, y), apply_lag=lagged), 300.0)
            close("InflationCurve.forward_index", lambda: ric.forward_index(RDate(d, m, y)), lambda: oic.forward_index(O.Date(d, m, y)), 300.0)
        ten, side = rng.choice(["2Y", "5Y", "12Y"]), rng.choice(["PAY", "RECEIVE"])
        rsw, osw = RZcis(rvd, ten, RSwap[side], 0.031, ri, 2e6), O.ZeroCouponInflationSwap(ovd, ten, O.SwapTypes[side], 0.031, oi, 2e6)
        close("ZeroCouponInflationSwap.pv01", lambda: rsw.pv01(rvd, rc), lambda: osw.pv01(ovd, oc), 2e6)
        close("ZeroCouponInflationSwap.breakeven_inflation_rate", lambda: rsw.breakeven_inflation_rate(rvd, rc, ric),
              lambda: osw.breakeven_inflation_rate(ovd, oc, oic), 1.0)
        ry = RYoY(rvd, ten, RSwap[side], 0.03, ri, RFreq.ANNUAL, 5e6, 0.001)
        oy = O.YoYInflationSwap(ovd, ten, O.SwapTypes[side], 0.03, oi, O.FrequencyTypes.ANNUAL, 5e6, 0.001)
        close("YoYInflationSwap.value", lambda: ry.value(rvd, rc, ric), lambda: oy.value(ovd, oc, oic), 5e6)
        close("YoYInflationSwap.breakeven_rate", lambda: ry.breakeven_rate(rvd, rc, ric), lambda: oy.breakeven_rate(ovd, oc, oic), 1.0)
        close("YoYInflationSwap.pv01", lambda: ry.pv01(rvd, rc), lambda: oy.pv01(ovd, oc), 5e6)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20240430
    rng = random.Random(seed)
    with np.errstate(all="ignore"):
        dates(rng, 3000)
        day_counts_and_schedules(rng, 1500)
        credit(rng, 200)
        swaps(rng, 150)
        bootstraps(rng, 200, 12)
        inflation(rng, 60)
    print(f"seed {seed}: {sum(calls.values())} paired calls over {len(calls)} functions")
    print("largest scaled difference per function:", {k: f"{v:.1e}" for k, v in sorted(worst.items())})
    print("mismatches:", {k: (len(v), v[:2]) for k, v in mismatch.items()} or "none")
    return 1 if mismatch or max(worst.values()) > 1e-12 else 0


if __name__ == "__main__":
    sys.exit(main())
