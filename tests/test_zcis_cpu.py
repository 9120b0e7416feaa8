"""Host side of the ZCIS path (index lag / fixings / interpolation / seasonality, inflation curve, cashflow
amounts) against goldens produced by the unmodified reference (tests/golden/gen/make_golden_zcis.py)."""
import numpy as np
import pytest

from adrates_b200 import Date, SwapTypes, CurrencyTypes, LibError
from adrates_b200.curves import DiscountCurve
from adrates_b200.global_types import InterpTypes
from adrates_b200.inflation import InflationIndex, InflationIndexTypes, ZeroCouponInflationSwap
from tests.conftest import load_golden
from tests.util_zcis import make_index, make_inflation_curve, make_zcis


@pytest.fixture(scope="module")
def g():
    return load_golden("ref_zcis.json")


def test_inflation_curve_nodes_and_projection(g):
    vd = Date(*g["value_dt"])
    for name, ref in g["inflation_curves"].items():
        ic = make_inflation_curve(g, name, make_index(g, name))
        assert ic._interp_type.name == ref["interp"]
        assert np.allclose(ic._times, ref["times"], rtol=0, atol=1e-15)
        assert np.allclose(ic._dfs, ref["dfs"], rtol=1e-15, atol=0)
        for y, v in ref["forward_index"].items():
            assert abs(ic.forward_index(vd.add_years(float(y))) - v) <= 1e-12 * v, (name, y)


def test_cashflow_amounts_match_reference(g):
    """Everything the host computes for a trade: CPI lookups, returns, payment date, breakeven."""
    idx = {n: make_index(g, n) for n in g["index_specs"]}
    ic = {n: make_inflation_curve(g, n, idx[n]) for n in g["index_specs"]}
    for t in g["trades"]:
        z = make_zcis(t, idx[t["index"]])
        (pay_dt, fixed), (_, infl) = z.cashflows(ic[t["index"]])
        assert [pay_dt.d(), pay_dt.m(), pay_dt.y()] == t["payment_dt"], t["id"]
        assert abs(z._fixed_return - t["fixed_return"]) <= 1e-14, t["id"]
        assert abs(z._inflation_leg._base_index - t["base_index"]) <= 1e-12 * t["base_index"], t["id"]
        assert abs(z._inflation_leg._final_index - t["final_index"]) <= 1e-12 * t["final_index"], t["id"]
        # undiscounted signed amounts: reference PVs divided by its payment DF
        assert abs(fixed - t["fixed_pv"] / t["payment_df"]) <= 1e-10 * t["notional"], t["id"]
        assert abs(infl - t["inflation_pv"] / t["payment_df"]) <= 1e-10 * t["notional"], t["id"]
        assert abs(z.breakeven_inflation_rate(Date(*g["value_dt"]), None, ic[t["index"]]) - t["breakeven"]) <= 1e-13


def test_index_errors_like_reference(g):
    idx = make_index(g, "rpi_linear")
    ic = make_inflation_curve(g, "rpi_linear", idx)
    idx.set_inflation_curve(ic)
    with pytest.raises(LibError, match="Cannot project CPI before value date"):
        ic.forward_index(Date(1, 1, 2024))
    bare = InflationIndex(InflationIndexTypes.UK_RPI, Date(1, 3, 2024), 293.0, CurrencyTypes.GBP)
    with pytest.raises(LibError, match="No fixing available"):
        bare.get_index(Date(1, 1, 2030))
    with pytest.raises(LibError, match="Base index must be positive"):
        InflationIndex(InflationIndexTypes.UK_RPI, Date(1, 3, 2024), 0.0, CurrencyTypes.GBP)
    with pytest.raises(LibError, match="Seasonality factors must include all months"):
        InflationIndex(InflationIndexTypes.UK_RPI, Date(1, 3, 2024), 293.0, CurrencyTypes.GBP, seasonality_factors={1: 1.0})
    with pytest.raises(LibError, match="Start date after maturity date"):
        ZeroCouponInflationSwap(Date(1, 3, 2025), Date(1, 3, 2024), SwapTypes.PAY, 0.03, bare)


def test_discount_curve_constructor_matches_reference_nodes(g):
    """DiscountCurve(value_dt, year offsets, dfs): node times of the reference's own curve object."""
    vd = Date(*g["value_dt"])
    c = DiscountCurve(vd, g["flat_curve"][0], np.array(g["flat_curve"][1]), InterpTypes.FLAT_FWD_RATES)
    ref = g["discount_curves"]["flat_ff"]
    assert np.allclose(c._times, ref["times"], rtol=0, atol=1e-15) and np.allclose(c._dfs, ref["dfs"], rtol=0, atol=0)
    # host df() (path A) against the reference's df() on the same dates
    q = [Date(*d) for d in g["df_queries"]["dates"]]
    from adrates_b200 import DayCountTypes
    got = [c.df(d, DayCountTypes.ACT_365F) for d in q]
    assert np.allclose(got, g["df_queries"]["dfs"]["flat_ff"], rtol=1e-14, atol=0)
