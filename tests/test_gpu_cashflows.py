"""Position.compute([VALUE, DELTA, GAMMA, CASHFLOWS]) on the device against the unmodified reference engine
(tests/golden/ref_cashflows.json): the cashflow table (path-A discount factors from cav_curve_df) and, for the trades rolled
on HOLIDAY calendars (UNITED_KINGDOM / TARGET / UNITED_STATES), VALUE / delta / gamma of the engine path to 1e-10."""
import numpy as np
import pytest

from adrates_b200 import RequestTypes
from adrates_b200.error import LibError
from tests.util_cashflows import ALL4, assert_rows_match, golden, make_cal_trade
from tests.util_trades import build_model, rel_err, trade_scales

pytestmark = pytest.mark.gpu


def test_cashflows_and_holiday_calendar_trades_match_the_reference(ref_curves):
    models = {}
    n_cal = 0
    for spec in golden():
        cv = ref_curves[spec["curve"]]
        model = models.setdefault(spec["curve"], build_model(cv))
        pos = make_cal_trade(spec, cv).position(model)
        if "error" in spec:
            with pytest.raises(LibError) as ex:
                pos.compute(ALL4)
            assert str(ex.value) in spec["error"]
            continue
        res = pos.compute(ALL4)
        s_pv, s_d, s_g = trade_scales(spec)
        e = (rel_err(res.value.amount, spec["value"], s_pv), rel_err(np.asarray(res.risk.risk_ladder), np.asarray(spec["delta"]), s_d),
             rel_err(np.asarray(res.gamma.risk_ladder), np.asarray(spec["gamma"]), s_g))
        assert max(e) < 1e-10, (spec["id"], e)
        assert_rows_match(res.cashflows, spec)
        n_cal += spec["cal"] != "WEEKEND"
        only = pos.compute([RequestTypes.CASHFLOWS])
        assert only.value is None and only.risk is None and len(only.cashflows) == len(spec["rows"])
    assert n_cal == 4


def test_bond_and_frn_cashflows_match_the_reference():
    """Position(bond / FRN).compute([..., CASHFLOWS]) on the device against Engine._compute_bond / _compute_frn of the unmodified
    reference (tests/golden/ref_cashflows_credit.json), next to VALUE where the reference offers it (dual-curve notes: VALUE only)."""
    from adrates_b200.credit import BOND_CURVE
    from tests.conftest import load_golden
    from tests.util_bonds import build_bond_model, make_bond, make_frn
    from tests.util_cashflows import assert_credit_rows_match, credit_golden
    g = credit_golden()
    model = build_bond_model(g)
    values = {b["id"]: b["value"] for b in load_golden("ref_bonds.json")["bonds"]}
    for rec in g["bonds"]:
        res = make_bond(rec).position(model).compute([RequestTypes.VALUE, RequestTypes.CASHFLOWS])
        assert_credit_rows_match(res.cashflows, rec, rec["face"])
        assert abs(res.value.amount - values[rec["id"]]) <= 1e-10 * rec["face"]
    for rec in g["frns"]:
        f = make_frn({**rec})
        dual = BOND_CURVE[f._currency] != f._floating_index
        if "error" in rec:
            with pytest.raises(LibError):
                f.position(model).compute([RequestTypes.CASHFLOWS])
            continue
        res = f.position(model).compute([RequestTypes.VALUE, RequestTypes.CASHFLOWS])
        assert_credit_rows_match(res.cashflows, rec, rec["face"])
        assert res.value is not None
        only = f.position(model).compute([RequestTypes.CASHFLOWS])
        assert only.value is None and len(only.cashflows) == len(rec["rows"])
        if dual:
            with pytest.raises(LibError):
                f.position(model).compute([RequestTypes.DELTA, RequestTypes.CASHFLOWS])


def test_yoy_and_collateral_cashflows_match_the_reference():
    """Position(yoy).compute([VALUE, CASHFLOWS]) on the device against Engine._compute_yoy_iis of the unmodified reference
    (tests/golden/ref_cashflows_yoy.json: the fixed leg's rows, or the CPI look-up error of sub-annual / seasoned swaps), and the
    empty table the reference returns for an OIS with cross-currency collateral (engine.py:497-501)."""
    from adrates_b200.position import Position
    from tests.conftest import load_golden
    from tests.util_yoy import make_model, make_swap
    g = load_golden("ref_yoy.json")
    cf_g = {c["id"]: c for c in load_golden("ref_cashflows_yoy.json")["cases"]}
    n_ok = 0
    for name in g["inflation_curves"]:
        model, idx, ic = make_model(g, name)
        for c in (c for c in g["cases"] if c["index"] == name):
            ref = cf_g[c["id"]]["value_cf"]
            pos = Position(make_swap(c, idx), model)
            if "error" in ref:
                with pytest.raises(LibError) as ex:
                    pos.compute([RequestTypes.VALUE, RequestTypes.CASHFLOWS])
                assert "LibError: " + str(ex.value) == ref["error"]
                continue
            res = pos.compute([RequestTypes.VALUE, RequestTypes.CASHFLOWS])
            assert abs(res.value.amount - ref["value"]) <= 1e-10 * c["notional"]
            assert res.risk is None and res.gamma is None
            assert_rows_match(res.cashflows, dict(ref, notional=c["notional"]))
            n_ok += 1
    assert n_ok == 6
    from adrates_b200 import CollateralType
    from tests.util_trades import make_trade
    from tests.util_xccy import build_xccy_model
    cmodel = build_xccy_model(load_golden("ref_xccy.json"), xccy_name="GBP_USD_XCCY")
    t = load_golden("ref_collateral.json")["trades"][0]
    sw = make_trade(dict(t, payment_lag=0), {"name": "GBP_OIS_SONIA", "dc": "ACT_365F"})
    res = sw.position(cmodel).compute([RequestTypes.VALUE, RequestTypes.CASHFLOWS], collateral_type=CollateralType.USD)
    assert len(res.cashflows) == 0 and res.cashflows.currency.name == "GBP" and res.cashflows.total_pv == 0
    assert abs(res.value.amount - t["value"]) <= 1e-10 * max(abs(t["value"]), t["notional"])
    assert sw.position(cmodel).compute([RequestTypes.VALUE], collateral_type=CollateralType.USD).cashflows is None
