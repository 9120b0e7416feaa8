"""Year-on-year inflation swaps (Engine._compute_yoy_iis, engine.py:986-1350): schedules, the two flat layouts and
the inflation-curve tables against 14 swaps valued by the unmodified reference engine (tests/golden/ref_yoy.json).
No GPU: the flat arrays are evaluated with the dense numpy restatement of the kernel formulas."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200 import LibError, RequestTypes
from adrates_b200.yoy_engine import discount_flat, inflation_flat, inflation_tables, yoy_arrays
from tests.conftest import load_golden
from tests.flat_eval import eval_flat
from tests.util_trades import rel_err
from tests.util_yoy import make_model, make_swap, scales

TOL = 1e-10


@pytest.fixture(scope="module")
def g():
    return load_golden("ref_yoy.json")


def test_inflation_curve_and_tables(g):
    for name, ref in g["inflation_curves"].items():
        _, _, ic = make_model(g, name)
        assert ic._interp_type.name == ref["interp"]
        assert np.allclose(ic._times, ref["times"], rtol=0, atol=1e-15) and np.allclose(ic._dfs, ref["dfs"], rtol=1e-15)
        F, J, C = inflation_tables(ic)
        assert np.allclose(F, ref["dfs"], rtol=1e-15)
        # closed-form derivatives of (1 + b)^T against central differences
        b = np.array([z._fixed_rate for z in ic._used_swaps])
        T = np.array(ic.swap_times)
        h = 1e-5
        fd1 = ((1 + b + h) ** T - (1 + b - h) ** T) / (2 * h)
        fd2 = ((1 + b + h) ** T - 2 * (1 + b) ** T + (1 + b - h) ** T) / (h * h)
        k = np.arange(len(b))
        assert np.allclose(J[k + 1, k], fd1, rtol=1e-6) and np.allclose(C[k + 1, k, k], fd2, rtol=1e-4, atol=1e-4)
        assert np.count_nonzero(J) == len(b) and np.count_nonzero(C) <= len(b) and not J[0].any()


def test_schedules_match_reference(g):
    for c in g["cases"]:
        _, idx, _ = make_model(g, c["index"]) if c is g["cases"][0] else (None, make_model(g, c["index"])[1], None)
        leg = make_swap(c, idx)._inflation_leg
        assert [[d.d(), d.m(), d.y()] for d in leg._payment_dts] == c["payment_dts"], c["id"]
        assert [[d.d(), d.m(), d.y()] for d in leg._yoy_start_dts] == c["yoy_start_dts"], c["id"]
        assert leg._year_fracs == c["year_fracs"], c["id"]


def test_flat_layouts_reproduce_reference_engine(g):
    """VALUE + both ladders + both gamma matrices of every golden swap, single and as one summed book."""
    for name in g["inflation_curves"]:
        model, idx, ic = make_model(g, name)
        disc = model.curves.GBP_OIS_SONIA
        plan = orc.plan_path_b(disc.swap_times, disc.year_fracs)
        d, J, C = orc.bootstrap_tables(disc.swap_rates, plan)
        F, Ji, Ci = inflation_tables(ic)
        cases = [c for c in g["cases"] if c["index"] == name]
        legs = [yoy_arrays(make_swap(c, idx), model.value_dt) for c in cases]
        fd = discount_flat(legs, disc, np.asarray(ic._times), F, ic._interp_type)
        fi = inflation_flat(legs, plan["times"], d, disc._interp_type, np.asarray(ic._times), ic._interp_type)
        assert fd.n_trades == fi.n_trades == len(cases) and fi.n_pairs == 6
        pv, dl, gm = eval_flat(fd, d, J, C)
        _, dli, gmi = eval_flat(fi, F, Ji, Ci)
        Ri = len(ic.swap_times)
        for i, c in enumerate(cases):
            s_pv, s_d, s_g = scales(c)
            e = (rel_err(pv[i], c["value"], s_pv), rel_err(dl[i], c["disc_delta"], s_d),
                 rel_err(gm[i], c["disc_gamma"], s_g), rel_err(dli[i, :Ri], c["infl_delta"], s_d),
                 rel_err(gmi[i, :Ri, :Ri], c["infl_gamma"], s_g))
            assert max(e) < TOL, (c["id"], e)
            assert np.abs(np.array(c["infl_delta"])).max() > 0 and np.abs(np.array(c["disc_delta"])).max() > 0


def test_errors_like_reference(g):
    from adrates_b200.yoy_engine import compute_yoy
    model, idx, _ = make_model(g, "rpi_linear")
    sw = make_swap(g["cases"][0], idx)
    del model._curves_dict["GBP_RPI_INFLATION"]
    with pytest.raises(LibError, match="Inflation curve GBP_RPI_INFLATION not found in model"):
        compute_yoy([sw], model, [RequestTypes.VALUE])
    del model._curves_dict["GBP_OIS_SONIA"]
    with pytest.raises(LibError, match="Discount curve GBP_OIS_SONIA not found in model"):
        compute_yoy([sw], model, [RequestTypes.VALUE])
