"""Host interpolation (adrates_b200/interpolator.py, DiscountCurve.df) against known answers of the unmodified reference
(tests/golden/ref_interp.json from tests/golden/gen/make_golden_interp.py): the module-level `interpolate` on the three node
schemes, `Interpolator.fit / interpolate` on the five spline schemes, `DiscountCurve.df` over dates in all eight."""
import numpy as np
import pytest

from adrates_b200 import Date, DiscountCurve, InterpTypes, LibError
from adrates_b200.interpolator import Interpolator, interpolate
from tests.conftest import load_golden

NODE = ("FLAT_FWD_RATES", "LINEAR_FWD_RATES", "LINEAR_ZERO_RATES")


def _close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.all(np.abs(a[fin] - b[fin]) <= tol * np.maximum(1.0, np.abs(b[fin]))), np.max(np.abs(a[fin] - b[fin]))


def test_interpolate_function_matches_reference():
    g = load_golden("ref_interp.json")
    for key, rec in g["function"].items():
        name, scheme = key.split("/")
        t, d = (np.array(g["node_sets"][name][k]) for k in ("times", "dfs"))
        m = InterpTypes[scheme].value
        got = [interpolate(float(q), t, d, m) for q in rec["q"]]
        assert all(isinstance(x, float) for x in got)
        _close(got, rec["scalar"], 1e-15)
        _close(interpolate(np.array(rec["q"]), t, d, m), rec["array"], 1e-15)
    with pytest.raises(LibError, match="must all be >= 0"):
        interpolate(-0.1, t, d, m)
    with pytest.raises(LibError, match="must all be >= 0"):
        interpolate(np.array([0.5, -0.1]), t, d, m)
    with pytest.raises(LibError, match="Invalid interpolation scheme"):
        interpolate(0.7, t, d, InterpTypes.PCHIP_ZERO_RATES.value)
    with pytest.raises(LibError):
        interpolate(1, t, d, m)                      # an int is not a recognised input type in the reference either


def test_interpolator_class_matches_reference():
    g = load_golden("ref_interp.json")
    assert len(g["class_scalar"]) == 20              # 4 node sets x 5 spline schemes
    for key, rec in g["class_scalar"].items():
        name, scheme = key.split("/")
        t, d = (np.array(g["node_sets"][name][k]) for k in ("times", "dfs"))
        f = Interpolator(InterpTypes[scheme])
        f.fit(t, d)
        with np.errstate(over="ignore"):         # far beyond the grid a cubic in the zero rate overflows exp: inf, as the reference
            res = [f.interpolate(float(q)) for q in rec["q"]]
            arr = f.interpolate(np.array(rec["q"]))
        assert [isinstance(r, np.ndarray) for r in res] == rec["is_array"]       # t < 1e-12 answers the float 1.0
        _close([np.asarray(r).reshape(-1)[0] for r in res], rec["v"], 1e-13)
        _close(arr, g["class_array"][key]["v"], 1e-13)
    # node schemes through the class: the same arithmetic as the function, a float for a float
    for scheme in NODE:
        t, d = (g["node_sets"]["from_zero"][k] for k in ("times", "dfs"))
        f = Interpolator(InterpTypes[scheme])
        f.fit(list(t), list(d))
        rec = g["function"]["from_zero/" + scheme]
        big = [q >= 1e-12 for q in rec["q"]]          # the class answers a float below g_small with 1.0 before any look-up
        _close([f.interpolate(float(q)) for q, b in zip(rec["q"], big) if b], [v for v, b in zip(rec["scalar"], big) if b], 1e-15)
        assert f.interpolate(1e-13) == 1.0 and f.interpolate(0.0) == 1.0
        _close(f.interpolate(np.array(rec["q"])), rec["array"], 1e-15)
        _close(f.simple_interpolate(np.array(rec["q"]), np.array(t), np.array(d), InterpTypes[scheme].value), rec["array"], 1e-15)
    one = Interpolator(InterpTypes.PCHIP_ZERO_RATES)
    one.fit(np.array([1.0]), np.array([0.95]))       # a single node: nothing fitted, no error
    assert one._interp_fn is None and one._times is not None
    with pytest.raises(LibError, match="Dfs have not been set"):
        Interpolator(InterpTypes.FLAT_FWD_RATES).interpolate(1.0)
    with pytest.raises(LibError, match="not a recognized type"):
        f.interpolate(1)


def test_discount_curve_df_matches_reference_in_every_scheme():
    g = load_golden("ref_interp.json")
    inp = g["curve_df_inputs"]
    vd = Date(*inp["value_dt"])
    dates = [Date(*x) for x in inp["dates"]]
    for scheme, rec in g["curve_df"].items():
        c = DiscountCurve(vd, inp["offsets"], np.array(inp["values"]), InterpTypes[scheme])
        assert np.array_equal(c._times, rec["times"]) and np.array_equal(c._dfs, rec["dfs"])
        tol = 1e-15 if scheme in NODE else 1e-13
        singles = [c.df(x) for x in dates]
        assert [isinstance(s, np.ndarray) for s in singles] == rec["single_is_array"], scheme
        _close([np.asarray(s).reshape(-1)[0] for s in singles], rec["single"], tol)
        _close(c.df(dates), rec["list"], tol)
