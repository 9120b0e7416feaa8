"""ZCIS valuation on the GPU (cav_cashflow_pv / cav_curve_df through the host mirror of the reference API)
against ZeroCouponInflationSwap.value of the unmodified reference (tests/golden/ref_zcis.json)."""
import numpy as np
import pytest

from adrates_b200 import Date, DayCountTypes, LibError, _native
from adrates_b200.dates import times_from_dates
from adrates_b200.inflation import value_zcis_book, cashflow_pv
from tests.conftest import load_golden
from tests.util_zcis import make_index, make_inflation_curve, make_discount_curve, make_zcis

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def g():
    return load_golden("ref_zcis.json")


def test_curve_df_matches_reference_df(g):
    ctx = _native.lib(0)
    q = [Date(*d) for d in g["df_queries"]["dates"]]
    for name, ref in g["df_queries"]["dfs"].items():
        c = make_discount_curve(g, name)
        t = np.array([times_from_dates(d, c._value_dt, DayCountTypes.ACT_365F) for d in q])
        got = ctx.curve_df(c._interp_type.value, c._times, c._dfs, t)
        assert np.max(np.abs(got - np.array(ref)) / np.array(ref)) < 1e-13, name
        host = np.array([c.df(d, DayCountTypes.ACT_365F) for d in q])      # host mirror agrees with the device
        assert np.max(np.abs(got - host) / host) < 1e-14
    with pytest.raises(LibError, match="Interpolate times must all be >= 0"):
        ctx.curve_df(c._interp_type.value, c._times, c._dfs, np.array([-0.5]))


def test_zcis_values_match_reference(g):
    vd = Date(*g["value_dt"])
    idx = {n: make_index(g, n) for n in g["index_specs"]}
    ic = {n: make_inflation_curve(g, n, idx[n]) for n in g["index_specs"]}
    dc = {n: make_discount_curve(g, n) for n in g["discount_curves"]}
    for t in g["trades"]:
        z = make_zcis(t, idx[t["index"]])
        pv = z.value(vd, dc[t["discount"]], ic[t["index"]])
        assert abs(pv - t["value"]) <= TOL * max(abs(t["value"]), t["notional"]), t["id"]
        leg = z._inflation_leg.value(vd, dc[t["discount"]], ic[t["index"]])
        assert abs(leg - t["inflation_pv"]) <= TOL * max(abs(t["inflation_pv"]), t["notional"]), t["id"]


def test_zcis_book_in_one_call(g):
    """All trades of one (index, discount curve) pair in a single device call; total = fixed-order sum."""
    vd = Date(*g["value_dt"])
    for iname in g["index_specs"]:
        idx = make_index(g, iname)
        ic = make_inflation_curve(g, iname, idx)
        for dname in g["discount_curves"]:
            ts = [t for t in g["trades"] if t["index"] == iname and t["discount"] == dname]
            book = [make_zcis(t, idx) for t in ts]
            pv = value_zcis_book(book, vd, make_discount_curve(g, dname), ic)
            ref = np.array([t["value"] for t in ts])
            scale = np.maximum(np.abs(ref), np.array([t["notional"] for t in ts]))
            assert np.max(np.abs(pv - ref) / scale) < TOL, (iname, dname)


def test_cashflow_pv_edge_cases(g):
    vd = Date(*g["value_dt"])
    dc = make_discount_curve(g, "flat_ff")
    # empty book, a trade without cashflows, a cashflow on the value date (worth 0 like the reference), ragged lengths
    assert cashflow_pv(dc, vd, []).shape == (0,)
    pv = cashflow_pv(dc, vd, [[], [(vd, 5.0)], [(vd.add_days(365), 100.0)], [(vd.add_days(30 * k), 1.0) for k in range(1, 80)]])
    assert pv[0] == 0.0 and pv[1] == 0.0
    assert abs(pv[2] - 100.0 * dc.df(vd.add_days(365), DayCountTypes.ACT_365F)) < 1e-12
    ref3 = sum(dc.df(vd.add_days(30 * k), DayCountTypes.ACT_365F) for k in range(1, 80))
    assert abs(pv[3] - ref3) < 1e-12 * ref3
    # valuation after the curve date: DF(t)/DF(t_value)
    later = vd.add_days(100)
    pv = cashflow_pv(dc, later, [[(vd.add_days(500), 1.0)]])
    want = dc.df(vd.add_days(500), DayCountTypes.ACT_365F) / dc.df(later, DayCountTypes.ACT_365F)
    assert abs(pv[0] - want) < 1e-14


def test_cashflow_pv_on_device_resident_cashflows_equals_the_host_buffer_call():
    """cav_cashflow_pv_dev: same kernels on caller-owned device arrays, bit-identical PVs and total; a negative time -> NaN."""
    import torch
    from adrates_b200 import _native
    from adrates_b200.market_data import readme_gbp_curve
    curve = readme_gbp_curve()
    rng = np.random.default_rng(2)
    n = 5000
    cnt = rng.integers(1, 5, n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    t = rng.uniform(0.0, 45.0, int(off[-1]))
    amt = rng.uniform(-1e6, 1e6, int(off[-1]))
    ctx = _native.Context(0)
    args = (curve._interp_type.value, curve._times, curve._dfs, 0.25)
    pv_h, tot_h = ctx.cashflow_pv(*args, off, t, amt)
    dev = torch.device("cuda", 0)
    o_d, t_d, a_d = (torch.from_numpy(a).to(dev) for a in (off, t, amt))
    pv_d = torch.empty(n, dtype=torch.float64, device=dev)
    tot_d = torch.empty(1, dtype=torch.float64, device=dev)
    ctx.cashflow_pv_dev(*args, n, o_d.data_ptr(), t_d.data_ptr(), a_d.data_ptr(), pv_d.data_ptr(), tot_d.data_ptr())
    ctx.sync()
    assert np.array_equal(pv_d.cpu().numpy(), pv_h) and float(tot_d.item()) == tot_h
    t_d[int(off[17])] = -0.5
    ctx.cashflow_pv_dev(*args, n, o_d.data_ptr(), t_d.data_ptr(), a_d.data_ptr(), pv_d.data_ptr(), None)
    ctx.sync()
    got = pv_d.cpu().numpy()
    assert np.isnan(got[17]) and np.array_equal(np.delete(got, 17), np.delete(pv_h, 17))
