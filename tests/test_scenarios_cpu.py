"""Host side of the batched scenario revaluation: Model.scenario_rates must hold exactly the par rates of the curves
Model.scenario builds (models.py:507-557), and scenarios shard over ranks without gaps."""
import numpy as np
import pytest

from adrates_b200.scenarios import check_rates, scenario_bounds
from adrates_b200.error import LibError
from tests.util_trades import build_model


def test_scenario_rates_equal_the_rates_of_rebuilt_models(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    name = cv["name"]
    shocks = [0.01, -0.25, 1.0, {"10Y": 0.05, "2Y": -0.03}, {"1W": 0.2, "50Y": -0.1, "7Y": 0.0}, {}]
    got = model.scenario_rates(name, shocks)
    assert got.shape == (len(shocks), len(cv["px"]))
    for s, shock in enumerate(shocks):
        ref = model.scenario(name, shock).curves[name].swap_rates
        assert np.array_equal(got[s], np.asarray(ref)), shock                 # bitwise: same arithmetic, same order
    par = np.array([0.01, -0.25])
    assert np.array_equal(model.scenario_rates(name, par), model.scenario_rates(name, list(par)))
    rng = np.random.default_rng(3)
    per_pillar = rng.normal(0, 0.1, (4, len(cv["px"])))
    got = model.scenario_rates(name, per_pillar)
    for s in range(4):
        ref = model.scenario(name, dict(zip(cv["tenors"], per_pillar[s]))).curves[name].swap_rates
        assert np.array_equal(got[s], np.asarray(ref))
    with pytest.raises(ValueError):
        model.scenario_rates("USD_OIS_SOFR", [0.01])
    with pytest.raises(ValueError):
        model.scenario_rates(name, np.zeros((2, 5)))
    curve = model.curves[name]
    with pytest.raises(LibError):
        check_rates(curve, np.zeros((3, 5)))
    with pytest.raises(LibError):
        check_rates(curve, np.full((1, len(cv["px"])), np.nan))


def test_scenario_bounds_cover_without_gaps():
    for n, world in ((10_000, 8), (7, 8), (0, 4), (13, 1), (100, 3)):
        b = scenario_bounds(n, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
