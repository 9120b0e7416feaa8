"""The public surface of the reference modules on or next to the path - class names, enum members, methods, functions and
the NAMES / order / optionality of their parameters (tests/golden/ref_api_names.json, read with `ast` from the unmodified
reference by tests/golden/gen/make_api_names.py) - against this package: a user of the reference who calls by keyword or by
position finds the same names here.  What is deliberately not mirrored is listed in EXCLUDED with the reason."""
import enum
import importlib
import inspect

import pytest

from tests.conftest import load_golden

# reference module -> modules of this package searched for its names, in order
HOME = {
    "cavour/utils/date.py": ["dates"], "cavour/utils/calendar.py": ["dates"], "cavour/utils/day_count.py": ["dates"],
    "cavour/utils/schedule.py": ["dates"], "cavour/utils/frequency.py": ["dates"], "cavour/utils/global_types.py": ["global_types"],
    "cavour/utils/currency.py": ["global_types"], "cavour/utils/error.py": ["error"],
    "cavour/market/curves/interpolator.py": ["interpolator", "global_types"], "cavour/market/curves/discount_curve.py": ["curves"],
    "cavour/market/curves/inflation_curve.py": ["inflation"], "cavour/market/indices/inflation_index.py": ["inflation"],
    "cavour/market/position/position.py": ["position"], "cavour/market/portfolio/portfolio.py": ["position"],
    "cavour/requests/results.py": ["results", "cashflows"], "cavour/models/models.py": ["models"],
    "cavour/trades/rates/ois.py": ["trades"], "cavour/trades/rates/ois_curve.py": ["curves"],
    "cavour/trades/rates/swap_fixed_leg.py": ["trades"], "cavour/trades/rates/swap_float_leg.py": ["trades"],
    "cavour/trades/rates/xccy_basis_swap.py": ["trades"], "cavour/trades/rates/xccy_curve.py": ["xccy_curve"],
    "cavour/trades/rates/xccy_fix_float_swap.py": ["trades"], "cavour/trades/rates/xccy_fix_fix_swap.py": ["trades"],
    "cavour/trades/rates/zcis.py": ["inflation"], "cavour/trades/rates/swap_inflation_leg.py": ["inflation"],
    "cavour/trades/rates/yoy_inflation_swap.py": ["inflation"], "cavour/trades/rates/swap_yoy_inflation_leg.py": ["inflation"],
    "cavour/trades/credit/bond.py": ["credit"], "cavour/trades/credit/frn.py": ["credit"],
}

# names of the reference that are not mirrored, with the reason (checked: every entry must exist in the reference table)
EXCLUDED = {
    # internals of the reference's date counter / its own test hook
    "cavour/utils/date.py": {"parse_dt", "calculate_list", "date_index", "date_from_index", "weekday", "vectorisation_helper",
                             "test_type", "Date._refresh", "Date._print"},
    # the per-country rule chains: evaluated once into day-serial tables (adrates_b200/holidays.py)
    "cavour/utils/calendar.py": {"Calendar.holiday_" + c for c in (
        "weekend", "australia", "united_kingdom", "france", "sweden", "germany", "switzerland", "japan", "new_zealand", "norway",
        "united_states", "canada", "italy", "target", "none")},
    "cavour/utils/schedule.py": {"Schedule._print"},
    # bootstraps: one planner + device kernels replace the host builders (curves.py, xccy_curve.py, inflation.py)
    "cavour/trades/rates/ois_curve.py": {"OISCurve._prepare_curve_builder_inputs", "OISCurve._build_curve_ad", "OISCurve._build_curve"},
    "cavour/trades/rates/xccy_curve.py": {"XccyCurve._prepare_curve_builder_inputs", "XccyCurve._build_curve", "XccyCurve._build_curve_ad",
                                          "XccyCurve._prepare_ad_inputs", "XccyCurve._run_jax_bootstrap", "XccyCurve._run_jax_bootstrap_impl"},
    "cavour/market/curves/inflation_curve.py": {"InflationCurve._prepare_curve_builder_inputs", "InflationCurve._build_curve",
                                                "InflationCurve._build_curve_ad"},
    "cavour/market/curves/discount_curve.py": {"DiscountCurve._df_ad", "DiscountCurve._linear_forward_interp", "DiscountCurve._print"},
    # Bloomberg market data
    "cavour/models/models.py": {"Model.prebuilt_curve", "Model.prebuilt_fx", "Model.prebuilt_xccy_curve"},
    # Plotly figures
    "cavour/requests/results.py": {"Gamma.plot", "CrossGamma.plot"},
}
PRINTERS = ("_print", "print_payments", "print_valuation")     # console reports: offered for legs and swaps, not for every class


def _find(ref_module, name):
    for m in HOME[ref_module]:
        mod = importlib.import_module("adrates_b200." + m)
        if hasattr(mod, name):
            return getattr(mod, name)
    return None


def _params(obj):
    try:
        sig = inspect.signature(obj)
    except (TypeError, ValueError):
        return None
    return [p for p in sig.parameters.values()]


def _check_signature(where, ref_params, ours, problems):
    ours = _params(ours)
    if ours is None:
        return
    ref = [p for p in ref_params if p["name"] not in ("self", "cls") and not p["name"].startswith("*")]
    mine = [p for p in ours if p.name not in ("self", "cls") and p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)]
    names = [p.name for p in mine]
    for i, rp in enumerate(ref):
        if i >= len(names) or names[i] != rp["name"]:
            problems.append(f"{where}: parameter {i} is '{names[i] if i < len(names) else None}', reference '{rp['name']}'")
            return
        if rp["default"] and mine[i].default is inspect.Parameter.empty:
            problems.append(f"{where}: parameter '{rp['name']}' is optional in the reference")
    for p in mine[len(ref):]:
        if p.default is inspect.Parameter.empty:
            problems.append(f"{where}: extra required parameter '{p.name}'")


def test_public_surface_matches_reference():
    table = load_golden("ref_api_names.json")
    problems, seen_excluded = [], set()
    for ref_module, spec in sorted(table.items()):
        skip = EXCLUDED.get(ref_module, set())
        for fname, f in spec["functions"].items():
            if fname.startswith("_") and fname not in ("_uinterpolate", "_vinterpolate"):
                continue
            if fname in skip:
                seen_excluded.add((ref_module, fname))
                continue
            ours = _find(ref_module, fname)
            if ours is None:
                problems.append(f"{ref_module}: function {fname} missing")
            else:
                _check_signature(f"{ref_module}:{fname}", f["params"], ours, problems)
        for cname, c in spec["classes"].items():
            if cname in skip:
                seen_excluded.add((ref_module, cname))
                continue
            cls = _find(ref_module, cname)
            if cls is None:
                problems.append(f"{ref_module}: class {cname} missing")
                continue
            if inspect.isclass(cls) and issubclass(cls, enum.Enum):
                missing = [m for m in c["assigned"] if m not in cls.__members__]
                if missing:
                    problems.append(f"{ref_module}: enum {cname} lacks {missing}")
                continue
            for mname, m in c["methods"].items():
                key = f"{cname}.{mname}"
                if key in skip:
                    seen_excluded.add((ref_module, key))
                    continue
                if mname.startswith("__") and mname != "__init__":
                    continue
                if mname in PRINTERS and not hasattr(cls, mname):
                    continue
                if not hasattr(cls, mname):
                    problems.append(f"{ref_module}: {key} missing")
                    continue
                attr = inspect.getattr_static(cls, mname)
                if m["property"]:
                    if not isinstance(attr, property):
                        problems.append(f"{ref_module}: {key} is a property in the reference")
                    continue
                if isinstance(attr, property):
                    problems.append(f"{ref_module}: {key} is a method in the reference, a property here")
                    continue
                _check_signature(f"{ref_module}:{key}", m["params"], getattr(cls, mname), problems)
    assert not problems, "\n".join(problems)
    stale = {(m, n) for m, names in EXCLUDED.items() for n in names} - seen_excluded
    assert not stale, f"EXCLUDED lists names the reference does not have: {sorted(stale)}"


def _samples():
    """One instance of every class of the table that can be built without a device."""
    import numpy as np
    from adrates_b200 import (Bond, BusDayAdjustTypes, Calendar, CalendarTypes, CurrencyTypes, CurveTypes, Date, DayCount, DayCountTypes,
                              DiscountCurve, FRN, FrequencyTypes, InflationCurve, InflationIndex, InflationIndexTypes, InterpTypes,
                              LibError, Model, OIS, Portfolio, Schedule, SwapFixedLeg, SwapFloatLeg, SwapInflationLeg,
                              SwapTypes, SwapYoYInflationLeg, XccyBasisSwap, XccyCurve, XccyFixFix, XccyFixFloat, YoYInflationSwap,
                              ZeroCouponInflationSwap)
    from adrates_b200.interpolator import Interpolator
    from adrates_b200.models import CurveAccessor
    from adrates_b200.results import AnalyticsResult, Ladder, Risk
    from adrates_b200.cashflows import Cashflows
    vd = Date(30, 4, 2024)
    gbp, usd = CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_OIS_SOFR
    ois = OIS(vd, "5Y", SwapTypes.PAY, 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, gbp, CurrencyTypes.GBP)
    xkw = dict(domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.ANNUAL, domestic_dc_type=DayCountTypes.ACT_365F,
               foreign_dc_type=DayCountTypes.ACT_360, domestic_floating_index=gbp, foreign_floating_index=usd,
               domestic_currency=CurrencyTypes.GBP, foreign_currency=CurrencyTypes.USD)
    basis = XccyBasisSwap(vd, "2Y", 790_000.0, 1_000_000.0, 0.0, 0.002, **xkw)
    model = Model(vd)
    for name, dc, px in (("GBP_OIS_SONIA", DayCountTypes.ACT_365F, [4.5, 4.4, 4.3]), ("USD_OIS_SOFR", DayCountTypes.ACT_360, [5.2, 5.1, 5.0])):
        model.build_curve(name=name, px_list=px, tenor_list=["1Y", "2Y", "3Y"], spot_days=0, swap_type=SwapTypes.PAY, fixed_dcc_type=dc,
                          fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                          bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.FLAT_FWD_RATES)
    index = InflationIndex(InflationIndexTypes.UK_RPI, Date(1, 1, 2024), 290.0, CurrencyTypes.GBP)
    zcis = ZeroCouponInflationSwap(vd, "5Y", SwapTypes.PAY, 0.03, index)
    yoy = YoYInflationSwap(vd, "5Y", SwapTypes.PAY, 0.03, index, FrequencyTypes.ANNUAL)
    fit = Interpolator(InterpTypes.FLAT_FWD_RATES)
    return {
        "Date": vd, "Calendar": Calendar(CalendarTypes.WEEKEND), "DayCount": DayCount(DayCountTypes.ACT_360),
        "Schedule": Schedule(vd, vd.add_tenor("2Y")), "LibError": LibError("x"), "Interpolator": fit,
        "DiscountCurve": DiscountCurve(vd, [1.0, 2.0], np.array([0.95, 0.9])), "OISCurve": model.curves.GBP_OIS_SONIA,
        "CurveAccessor": CurveAccessor({}), "OIS": ois, "SwapFixedLeg": ois._fixed_leg, "SwapFloatLeg": ois._float_leg,
        "XccyBasisSwap": basis, "XccyFixFloat": XccyFixFloat(vd, "2Y", 790_000.0, 1_000_000.0, SwapTypes.PAY, 0.04, 0.001, **xkw),
        "XccyFixFix": XccyFixFix(vd, "2Y", 790_000.0, 1_000_000.0, SwapTypes.PAY, 0.04, 0.05, **xkw),
        "XccyCurve": XccyCurve(vd, [basis], model.curves.GBP_OIS_SONIA, model.curves.USD_OIS_SOFR, 0.79),
        "Bond": Bond(vd, "5Y", 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.GBP),
        "FRN": FRN(vd, "3Y", 0.002, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP, gbp),
        "InflationIndex": index, "ZeroCouponInflationSwap": zcis, "SwapInflationLeg": zcis._inflation_leg,
        "InflationCurve": InflationCurve(vd, [zcis, ZeroCouponInflationSwap(vd, "10Y", SwapTypes.PAY, 0.032, index)], 293.8,
                                         CurrencyTypes.GBP, InflationIndexTypes.UK_RPI),
        "YoYInflationSwap": yoy, "SwapYoYInflationLeg": yoy._inflation_leg, "Portfolio": Portfolio([]),
        "Position": ois.position(model), "AnalyticsResult": AnalyticsResult(), "Risk": Risk([]), "Ladder": Ladder({}, "c"),
        "Cashflows": Cashflows([], CurrencyTypes.GBP),
    }


# constructor attributes of the reference that this package does not keep, with the reason
ATTRS_NOT_KEPT = {
    "Date": {"_excel_dt"},                                        # excel_dt() computes it from the day serial
    "Position": {"_engine"},                                      # the engine is made per compute() call (its curve cache is the device session)
    "XccyCurve": {"_use_ad", "_interpolator"},                    # one bootstrap (planner + exact tangents); node look-ups are functions of (_times, _dfs)
    "SwapFixedLeg": {"_payment_dts_ad"}, "SwapFloatLeg": {"_payment_dts_ad", "_payment_dts_float"},   # float-year copies for jax tracing
    "Schedule": {"_adjust_termination_dt", "_first_dt", "_next_to_last_dt"},   # stub-date options the reference stores and never uses
    "FRN": {"_rates", "_coupon_payments", "_payment_dfs", "_payment_pvs"},     # empty until value() fills them (credit_analytics.py)
    "LibError": set(),
}


def test_constructor_attributes_match_reference():
    """Every attribute the reference's constructors (and the schedule generators they call) leave on an object exists on ours:
    code that reads `swap._fixed_leg._payment_dts`, `curve._times`, `leg._year_fracs` ... keeps working."""
    table = load_golden("ref_api_names.json")
    samples = _samples()
    missing, checked = [], 0
    for spec in table.values():
        for cname, c in spec["classes"].items():
            if not c["init_attrs"]:
                continue
            assert cname in samples, f"no sample instance for {cname}"
            skip = ATTRS_NOT_KEPT.get(cname, set())
            assert skip <= set(c["init_attrs"]), (cname, skip - set(c["init_attrs"]))
            for a in c["init_attrs"]:
                if a in skip:
                    continue
                checked += 1
                if not hasattr(samples[cname], a):
                    missing.append(f"{cname}.{a}")
    assert not missing, missing
    assert checked > 300


def test_argument_type_checks_match_reference():
    """The constructors (and the two value() methods) that call `check_argument_types` in the reference do so here, over the
    same annotations, with the same LibError("Argument Type Error") - helpers.py:618-636."""
    from adrates_b200 import (CurrencyTypes, CurveTypes, Date, DayCountTypes, DiscountCurve, FrequencyTypes, LibError, OIS, Schedule,
                              SwapTypes, Bond, FRN)
    table = load_golden("ref_api_names.json")
    n = 0
    for ref_module, spec in table.items():
        for cname, c in spec["classes"].items():
            for mname, m in c["methods"].items():
                if not m.get("checks_types"):
                    continue
                fn = getattr(_find(ref_module, cname), mname)
                assert "check_argument_types(" in inspect.getsource(fn), f"{cname}.{mname} does not check its argument types"
                ours = fn.__annotations__
                for p in m["params"]:
                    if "ann" in p:
                        got = ours.get(p["name"])
                        got = got if isinstance(got, str) else getattr(got, "__name__", str(got))
                        assert (got or "").replace(" ", "") == p["ann"].replace(" ", ""), (cname, mname, p["name"], got, p["ann"])
                n += 1
    assert n == 20
    vd = Date(30, 4, 2024)
    ok = dict(effective_dt=vd, term_dt_or_tenor="5Y", fixed_leg_type=SwapTypes.PAY, fixed_coupon=0.04, fixed_freq_type=FrequencyTypes.ANNUAL,
              fixed_dc_type=DayCountTypes.ACT_365F, floating_index=CurveTypes.GBP_OIS_SONIA, currency=CurrencyTypes.GBP)
    OIS(**dict(ok, fixed_coupon=4, notional=1_000_000, term_dt_or_tenor=vd.add_tenor("5Y")))      # ints pass for floats, a Date for a tenor
    for bad in (dict(effective_dt="2024-04-30"), dict(term_dt_or_tenor=5), dict(fixed_leg_type="PAY"), dict(fixed_coupon="4%"),
                dict(fixed_freq_type=1), dict(currency="GBP"), dict(payment_lag=1.5), dict(notional=None)):
        with pytest.raises(LibError, match="Argument Type Error"):
            OIS(**dict(ok, **bad))
    with pytest.raises(LibError, match="Argument Type Error"):
        DiscountCurve(vd, [1.0, 2.0], [0.95, 0.9])                     # values must be an array, as in the reference
    with pytest.raises(LibError, match="Argument Type Error"):
        Schedule(vd, "5Y")
    with pytest.raises(LibError, match="Argument Type Error"):
        Bond(vd, "5Y", 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.GBP, amortization_schedule=(80.0, 60.0))
    FRN(vd, "3Y", 0.002, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA, cap_rate=None)
    with pytest.raises(LibError, match="Argument Type Error"):
        FRN(vd, "3Y", 0.002, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA, cap_rate="5%")
