"""Host planner + flattener against the reference engine's outputs (no GPU): the flat
arrays are evaluated with a dense numpy restatement of the kernel formulas."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200.curves import OISCurve, plan_path_b, plan_queries
from adrates_b200.flatten import Flattener
from adrates_b200.global_types import InterpTypes
from tests.flat_eval import eval_flat
from tests.util_trades import make_calibration_swaps, make_trade, rel_err, trade_scales

TOL = 1e-10


def _curve(cv):
    vd, swaps = make_calibration_swaps(cv)
    return OISCurve(vd, swaps, InterpTypes[cv["interp"]])


def test_plan_matches_oracle_plan(ref_curves):
    for key, cv in ref_curves.items():
        p = plan_path_b(cv["swap_times"], cv["year_fracs"])
        o = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
        assert np.array_equal(p.node_time, o["times"]) and np.array_equal(p.node_acc, o["acc"])
        assert np.array_equal(p.node_prev, o["prev"]) and np.array_equal(p.node_swap, o["swap"])


@pytest.mark.parametrize("method", ["LINEAR_ZERO_RATES", "FLAT_FWD_RATES"])
def test_query_planner_matches_oracle_interpolation(ref_curves, method):
    cv = ref_curves["gbp_readme_lzr"]
    x = np.array(cv["pathB_times"])
    d = np.array(cv["pathB_dfs"])
    rng = np.random.default_rng(3)
    t = np.concatenate([x[::7], x[5:40] + 3e-11, x[5:40] - 7e-11, x[5:40] + 2e-10, rng.uniform(0, 60, 400),
                        [0.0, 1e-13, 50.03, 50.0328767124, 75.0]])
    a, b, wa, wb = plan_queries(t, x, InterpTypes[method])
    got = np.exp(wa * np.log(d[a]) + wb * np.log(d[b]))
    ref = np.array([orc.simple_interpolate(tt, x, d, InterpTypes[method].value, dual=False) for tt in t])
    assert np.max(np.abs(got - ref) / ref) < 1e-14


@pytest.mark.parametrize("dedup", [True, False])
def test_flat_portfolio_reproduces_reference_engine(ref_curves, ref_trades, dedup):
    for key, cv in ref_curves.items():
        specs = [s for s in ref_trades if s["curve"] == key]
        if not specs:
            continue
        curve = _curve(cv)
        fl = Flattener(curve)
        for s in specs:
            fl.add_trade(make_trade(s, cv))
        flat = fl.finalize(dedup=dedup)
        plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
        d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
        pv, dl, gm = eval_flat(flat, d, J, C)
        for i, s in enumerate(specs):
            s_pv, s_d, s_g = trade_scales(s)
            e = (rel_err(pv[i], s["value"], s_pv), rel_err(dl[i], s["delta"], s_d), rel_err(gm[i], s["gamma"], s_g))
            assert max(e) < TOL, (s["id"], dedup, e)


def test_flat_layout_invariants(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    fl = Flattener(curve)
    specs = [s for s in ref_trades if s["curve"] == "gbp_readme_lzr"]
    for s in specs * 3:          # repeated trades share units when dedup is on
        fl.add_trade(make_trade(s, cv))
    flat = fl.finalize(dedup=True, max_group=2)
    assert flat.n_trades == 3 * len(specs)
    assert flat.unit_offsets[0] == 0 and flat.unit_offsets[-1] == flat.n_terms
    assert flat.group_offsets[0] == 0 and flat.group_offsets[-1] == flat.n_trades
    assert np.all(np.diff(flat.group_offsets) <= 2) and np.all(np.diff(flat.group_offsets) >= 1)
    assert sorted(flat.out_index.tolist()) == list(range(flat.n_trades))
    assert flat.node.min() >= 0 and flat.node.max() < curve.path_b_plan().n_nodes
    direct = fl.finalize(dedup=False)
    assert direct.n_units == direct.n_trades == direct.n_groups and direct.out_index is None
