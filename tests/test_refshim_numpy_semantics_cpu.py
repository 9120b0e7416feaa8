"""The golden generators run the unmodified reference under tests/golden/gen/refshim/jax, a torch stand-in for the jax / jnp entry
points the reference calls.  The only library-DEFINED semantics the parity of off-grid cashflows rests on are `jnp.searchsorted
(side='right')` on a grid with duplicate nodes and `jnp.interp` (interpolator_ad.py:230-246).  JAX documents both as following
NumPy; no JAX is installable here, so this test anchors the shim's versions on NumPy itself: searchsorted bit for bit on grids
WITH duplicates and ties, interp to 1 ulp on strictly increasing grids (where NumPy defines it), end clamping, and the
zero-width-cell rule of jax's `_interp` (value of the left node) on duplicate nodes, where NumPy leaves the result open.
Also: the oracle's own bracket rule (what the CUDA planner reproduces) equals the shim's on the engine grid with its duplicates."""
import importlib.util
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def jnp():
    spec = importlib.util.spec_from_file_location("_refshim_jnp", os.path.join(HERE, "golden", "gen", "refshim", "jax", "numpy.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _grid_with_duplicates(rng, n):
    x = np.sort(rng.uniform(0.0, 50.0, n))
    dup = rng.choice(n - 1, n // 5, replace=False)
    x[dup + 1] = x[dup]                      # exact duplicates, as the engine grid holds at every swap maturity
    return np.sort(x)


def test_searchsorted_right_equals_numpy_on_duplicate_nodes(jnp):
    rng = np.random.default_rng(3)
    for n in (2, 7, 66, 264):
        x = _grid_with_duplicates(rng, n)
        q = np.concatenate([rng.uniform(-1.0, 51.0, 500), x, x + 1e-12, x - 1e-12, [x[0], x[-1]]])
        for side in ("left", "right"):
            got = jnp.searchsorted(torch.from_numpy(x), torch.from_numpy(q), side=side).numpy()
            assert np.array_equal(got, np.searchsorted(x, q, side=side)), (n, side)


def test_interp_equals_numpy_where_numpy_defines_it(jnp):
    rng = np.random.default_rng(4)
    for n in (2, 5, 66):
        xp = np.cumsum(rng.uniform(0.01, 2.0, n))
        fp = rng.normal(size=n)
        q = np.concatenate([rng.uniform(xp[0] - 1.0, xp[-1] + 1.0, 400), xp, [xp[0], xp[-1]]])
        got = jnp.interp(torch.from_numpy(q), torch.from_numpy(xp), torch.from_numpy(fp)).numpy()
        ref = np.interp(q, xp, fp)
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 4 * np.finfo(float).eps
        assert np.array_equal(got[q < xp[0]], np.full((q < xp[0]).sum(), fp[0]))          # clamped, not extrapolated
        assert np.array_equal(got[q > xp[-1]], np.full((q > xp[-1]).sum(), fp[-1]))      # (ON the last node: the cell formula, 1 ulp)


def test_interp_on_duplicate_nodes_takes_the_left_value_of_a_zero_width_cell(jnp):
    xp = np.array([0.0, 1.0, 1.0, 2.0, 2.0, 2.0, 5.0])
    fp = np.array([10.0, 11.0, 12.0, 13.0, 14.0, 15.0, 16.0])
    # at a duplicate node side='right' lands beyond the run: the cell is (last duplicate, next node), delta = 0 -> value of the
    # LAST duplicate; strictly inside a cell NumPy's and jax's formulas agree
    q = np.array([1.0, 2.0, 1.5, 3.5, 0.5])
    got = jnp.interp(torch.from_numpy(q), torch.from_numpy(xp), torch.from_numpy(fp)).numpy()
    assert np.array_equal(got, [12.0, 15.0, 12.5, 15.5, 10.5])
    # a query that lands ON the run from the right clamps to the last cell when the run is at the end of the grid
    xe = np.array([0.0, 1.0, 3.0, 3.0])
    fe = np.array([1.0, 2.0, 5.0, 7.0])
    got = jnp.interp(torch.from_numpy(np.array([3.0, 2.0])), torch.from_numpy(xe), torch.from_numpy(fe)).numpy()
    assert np.array_equal(got, [5.0, 3.5])       # zero-width last cell: value of its left node, fp[i-1]


def test_oracle_bracket_rule_equals_the_shim_on_the_engine_grid(jnp, ref_curves):
    """`simple_interpolate` for LINEAR_ZERO_RATES is jnp.interp of the zero rates at t + 1e-12 (after the 1e-10 snap): the
    oracle's bracket planner - the rule the CUDA planner is tested against bit for bit - must give the shim's answer on the
    engine's own grid (264 nodes, a duplicate at every swap maturity)."""
    from oracle import cavour_oracle as orc
    cv = ref_curves["gbp_readme_lzr"]
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d = orc.bootstrap_dfs(cv["swap_rates"], plan)
    x = np.asarray(plan["times"])
    assert np.any(np.diff(x) == 0.0)             # the duplicates are there
    rng = np.random.default_rng(9)
    t = np.concatenate([rng.uniform(0.003, 55.0, 300), x[x > 0][::7]])
    got = np.array([orc.simple_interpolate(float(u), x, d, orc.LINEAR_ZERO_RATES, dual=False) for u in t])
    zero = np.where(x > 0, -np.log(d) / np.maximum(x, 1e-15), 0.0)
    snapped = t.copy()
    for k, u in enumerate(t):                     # interpolator_ad.py:210-226: exact hits (1e-10) return the FIRST nearest node
        j = int(np.argmin(np.abs(x - u)))
        if abs(x[j] - u) < 1e-10:
            snapped[k] = np.nan
            assert got[k] == d[j]
    live = ~np.isnan(snapped)
    z = jnp.interp(torch.from_numpy(t[live] + 1e-12), torch.from_numpy(x), torch.from_numpy(zero)).numpy()
    ref = np.exp(-z * t[live])
    assert np.max(np.abs(got[live] - ref)) <= 1e-14
