"""Year-on-year inflation swaps through Position.compute / Portfolio.compute on the GPU against the unmodified
reference engine's results (tests/golden/ref_yoy.json)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from adrates_b200 import RequestTypes  # noqa: E402
from adrates_b200.position import Portfolio, Position  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.util_trades import rel_err  # noqa: E402
from tests.util_yoy import make_model, make_swap, scales  # noqa: E402

ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
TOL = 1e-10


def test_yoy_swaps_match_reference_engine():
    g = load_golden("ref_yoy.json")
    for name in g["inflation_curves"]:
        model, idx, ic = make_model(g, name)
        cases = [c for c in g["cases"] if c["index"] == name]
        positions = []
        for c in cases:
            pos = Position(make_swap(c, idx), model)
            positions.append(pos)
            res = pos.compute(ALL)
            s_pv, s_d, s_g = scales(c)
            e = (rel_err(res.value.amount, c["value"], s_pv),
                 rel_err(res.risk.GBP_OIS_SONIA.risk_ladder, c["disc_delta"], s_d),
                 rel_err(res.risk.GBP_RPI_INFLATION.risk_ladder, c["infl_delta"], s_d),
                 rel_err(res.gamma.GBP_OIS_SONIA.risk_ladder, c["disc_gamma"], s_g),
                 rel_err(res.gamma.GBP_RPI_INFLATION.risk_ladder, c["infl_gamma"], s_g))
            assert max(e) < TOL, (c["id"], e)
            assert list(res.risk.GBP_RPI_INFLATION.tenors) == c["infl_tenors"]
            assert res.value.currency.name == "GBP"
            only_v = pos.compute([RequestTypes.VALUE])
            assert only_v.risk is None and only_v.gamma is None
            assert abs(only_v.value.amount - c["value"]) <= TOL * s_pv
        # Portfolio.compute = sums over the positions (portfolio.py:48-65)
        tot = Portfolio(positions).compute(ALL)
        S = sum(c["notional"] for c in cases)
        assert abs(tot.value.amount - sum(c["value"] for c in cases)) <= TOL * S
        assert rel_err(tot.risk.GBP_RPI_INFLATION.risk_ladder, np.sum([c["infl_delta"] for c in cases], axis=0), S * 1e-4 * 40) < TOL
        assert rel_err(tot.gamma.GBP_OIS_SONIA.risk_ladder, np.sum([c["disc_gamma"] for c in cases], axis=0), S * 1e-8 * 1600) < TOL
        assert rel_err(tot.gamma.GBP_RPI_INFLATION.risk_ladder, np.sum([c["infl_gamma"] for c in cases], axis=0), S * 1e-8 * 1600) < TOL
