"""DiscountCurve rate views (zero_rate, cc_rate, swap_rate, fwd, _fwd, fwd_rate, bump, _zero_to_df, survival_prob; reference
discount_curve.py:96-296, 438-600) against known answers of the unmodified reference (tests/golden/ref_discount_curve.json,
tests/golden/gen/make_golden_discount_curve.py) on plain curves in four interpolation schemes and on the bootstrapped README
SONIA curve.  Host arithmetic on `df()` in the reference and here."""
import numpy as np
import pytest

from adrates_b200 import Date, DayCountTypes, DiscountCurve, FrequencyTypes, InterpTypes, LibError
from tests.conftest import load_golden
from tests.util_trades import build_model

TOL = 2e-13


def _near(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape and np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) <= tol, (a, b)


def _check(c, vd, g, ref):
    dts = [vd.add_tenor(t) for t in g["tenors"]]
    _near(c.zero_rate(dts), ref["zero_cont_act360"])
    z = c.zero_rate(dts[3])
    assert np.ndim(z) == 0
    _near(z, ref["zero_single"])
    _near(c.cc_rate(dts), ref["cc_rate"])
    _near(c.survival_prob(dts[2]), ref["survival"])
    _near(c.fwd(dts), ref["fwd"], 1e-10)              # a one-day log difference: 365 x rounding of the DFs
    _near(c.fwd(dts[1]), ref["fwd_single"], 1e-10)
    _near(c._fwd(np.array([0.0, 0.3, 1.0, 4.2, 11.0])), ref["_fwd"], 1e-9)     # 1e-6 central difference
    _near(c.fwd_rate(dts, "3M"), ref["fwd_rate_3m"])
    _near(c.fwd_rate(dts[0], dts[4], DayCountTypes.ACT_365F), ref["fwd_rate_single"])
    _near(c.fwd_rate(dts[:3], dts[3:]), ref["fwd_rate_lists"])
    _near(c.swap_rate(vd, [dts[2], dts[3], dts[4]]), ref["swap_rate"])
    one = c.swap_rate(vd.add_tenor("6M"), dts[3], FrequencyTypes.SEMI_ANNUAL, DayCountTypes.ACT_360)
    assert isinstance(one, np.ndarray) and one.size == 1              # always an array, as the reference returns it
    _near(one, ref["swap_rate_single"])
    for key, want in ref["zero"].items():
        fq, dc = key.split("/")
        _near(c.zero_rate(dts, FrequencyTypes[fq], DayCountTypes[dc]), want)
    b = c.bump(0.0025)
    assert type(b) is DiscountCurve
    _near(b._times, ref["bump"]["times"])
    _near(b._dfs, ref["bump"]["dfs"])
    _near(b.df(dts), ref["bump"]["df"])
    for fq, want in ref["zero_to_df"].items():
        _near(c._zero_to_df(vd, np.array([0.01, 0.03, 0.05]), np.array([0.0, 1.5, 7.0]), FrequencyTypes[fq], DayCountTypes.ACT_360), want)
    _near(c._zero_to_df(vd, 0.04, 2.5, FrequencyTypes.ANNUAL, DayCountTypes.ACT_360), ref["zero_to_df_scalar"])
    assert c.value_dt() == vd


def test_plain_curve_views_match_reference():
    g = load_golden("ref_discount_curve.json")
    vd = Date(*g["value_dt"])
    for scheme, ref in g["plain"].items():
        _check(DiscountCurve(vd, g["offsets"], np.array(g["values"]), InterpTypes[scheme]), vd, g, ref)


def test_ois_curve_views_match_reference(ref_curves):
    g = load_golden("ref_discount_curve.json")
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    _check(model.curves[cv["name"]], Date(*g["value_dt"]), g, g["ois_gbp_readme_lzr"])


def test_view_errors_are_the_references():
    vd = Date(30, 4, 2024)
    c = DiscountCurve(vd, [1.0, 2.0], np.array([0.95, 0.9]))
    with pytest.raises(LibError, match="Invalid Frequency type"):
        c.zero_rate(vd.add_tenor("1Y"), 2)
    with pytest.raises(LibError, match="Invalid Day Count type"):
        c.zero_rate(vd.add_tenor("1Y"), FrequencyTypes.ANNUAL, "ACT_360")
    with pytest.raises(LibError, match="starts before the curve valuation date"):
        c.swap_rate(vd.add_days(-1), vd.add_tenor("1Y"))
    with pytest.raises(LibError, match="simple yield freq"):
        c.swap_rate(vd, vd.add_tenor("1Y"), FrequencyTypes.SIMPLE)
    with pytest.raises(LibError, match="continuous freq"):
        c.swap_rate(vd, vd.add_tenor("1Y"), FrequencyTypes.CONTINUOUS)
    with pytest.raises(LibError, match="before the swap start date"):
        c.swap_rate(vd.add_tenor("1Y"), vd.add_tenor("1Y"))
    with pytest.raises(LibError, match="must be same types"):
        c.fwd_rate((vd,), "3M")
    with pytest.raises(LibError, match="Unknown Frequency type"):
        c._zero_to_df(vd, 0.02, 1.0, FrequencyTypes.TRI_ANNUAL)
    with pytest.raises(LibError, match="do not have same length"):
        c._df_to_zero([0.9, 0.8], [vd.add_tenor("1Y")], FrequencyTypes.ANNUAL, DayCountTypes.ACT_360)
