"""Export helpers of the result containers (adrates_b200/results.py: to_dict / to_json / to_csv / df / matrix, Risk.__repr__ /
has_cross_gamma / all_cross_gammas) against the strings the unmodified reference's containers produce
(tests/golden/ref_results_export.json, tests/golden/gen/make_golden_results_export.py)."""
import json

import numpy as np
import pytest

from adrates_b200 import CrossGamma, CurrencyTypes, CurveTypes, Delta, Gamma, Risk, Valuation


@pytest.fixture(scope="module")
def g():
    from tests.conftest import load_golden
    return load_golden("ref_results_export.json")


def _same_json(a: str, b: str):
    """equal documents; `total` (a sum whose order differs between array libraries) to an ulp"""
    x, y = json.loads(a), json.loads(b)
    assert abs(x.pop("total") - y.pop("total")) <= 1e-12 and x == y


def _objects(g):
    i = g["inputs"]
    lad, gam, x = np.array(i["ladder"]), np.array(i["gamma"]), np.array(i["cross"])
    d = Delta(lad, i["tenors"], CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    d2 = Delta(lad[:3] * 2, i["tenors2"], CurrencyTypes.GBP, CurveTypes.USD_GBP_BASIS)
    gm = Gamma(gam, i["tenors"], CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    cg = CrossGamma(x, i["tenors"], i["tenors2"], CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS, CurrencyTypes.GBP)
    return d, d2, gm, cg


def test_valuation_ladder_delta_exports(g):
    v = Valuation(1234.5678, CurrencyTypes.GBP)
    r = g["valuation"]
    assert v.to_dict() == r["to_dict"] and v.to_json() == r["to_json"] and v.to_csv() == r["to_csv"] and repr(v) == r["repr"]
    d = _objects(g)[0]
    r = g["ladder"]
    assert d.ladder.to_dict() == r["to_dict"] and d.ladder.df.to_csv() == r["df_csv"] and repr(d.ladder) == r["repr"]
    r = g["delta"]
    assert d.to_dict() == r["to_dict"] and d.to_json() == r["to_json"] and d.to_csv() == r["to_csv"] and repr(d) == r["repr"]
    assert d.df.shape == (5, 1) and d.df.index.name == "Tenor"


def test_gamma_and_cross_gamma_exports(g, capsys, tmp_path):
    _, _, gm, cg = _objects(g)
    r = g["gamma"]
    assert gm.to_dict == r["to_dict"] and repr(gm) == r["repr"] and gm.to_csv() == r["to_csv"]
    _same_json(gm.to_json(), r["to_json"])
    assert r["matrix"].startswith("error: ValueError")          # the reference formats tenor LABELS as numbers and raises;
    gm.matrix                                                    # here labels print as they are
    shown = capsys.readouterr().out
    assert "1D" not in shown and "1W" in shown and "2Y" in shown          # the all-zero 1D row / column is dropped
    num = Gamma(np.array(g["inputs"]["gamma"]), [0.0027, 0.0192, 0.0833, 1.0, 2.0], CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    num.matrix
    assert capsys.readouterr().out == g["gamma_numeric_tenors"]["matrix"]          # numeric tenors: the reference's table
    r = g["cross"]
    assert cg.to_dict == r["to_dict"] and repr(cg) == r["repr"] and cg.to_csv() == r["to_csv"]
    _same_json(cg.to_json(), r["to_json"])
    cg.matrix
    assert capsys.readouterr().out == r["matrix"]
    path = tmp_path / "g.csv"
    assert gm.to_csv(str(path)) is None and path.read_text() == g["gamma"]["to_csv"]
    assert Gamma(np.array([1.0, 2.0]), ["1Y", "2Y"], CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA).df.values.tolist() == [[1.0, 0.0], [0.0, 2.0]]


def test_risk_repr_and_cross_gamma_lookup(g):
    d, d2, gm, cg = _objects(g)
    r = g["risk"]
    assert repr(Risk([d, d2])) == r["repr"]
    risk = Risk([gm], cross_gammas=[cg])
    assert repr(risk) == r["gamma_repr"]
    assert [risk.has_cross_gamma(CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS),
            risk.has_cross_gamma(CurveTypes.USD_GBP_BASIS, CurveTypes.GBP_OIS_SONIA)] == r["has"]
    allc = risk.all_cross_gammas
    assert [list(k) for k in allc] == r["all_keys"] and allc is not risk._cross_gammas
