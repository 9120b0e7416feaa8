"""CASHFLOWS request, host logic (adrates_b200/cashflows.py) against rows produced by the unmodified reference engine
(tests/golden/ref_cashflows.json: engine.py:190-213 + _extract_leg_cashflows).  The discount factors come from the device in
the product (cav_curve_df); here a TEST DOUBLE stands in for the context so that the row assembly - ordering, signs, leg
names, dead rows, forward rates, the seasoned-swap error - is covered without a GPU.  tests/test_gpu_cashflows.py runs the
same goldens through the real path."""
import numpy as np
import pytest

from adrates_b200 import cashflows as CF
from adrates_b200 import position
from adrates_b200.error import LibError
from tests.util_cashflows import assert_rows_match, golden, make_cal_trade
from tests.util_trades import build_model


class _HostDfContext:
    """test double of _native.Context.curve_df: the host's scalar path-A interpolation (DiscountCurve._node_df)"""
    def __init__(self, curve):
        self.curve = curve

    def curve_df(self, interp_method, node_time, node_df, t):
        assert interp_method == self.curve._interp_type.value
        return np.array([self.curve._node_df(float(u)) for u in np.asarray(t)])


class _Session:
    def __init__(self, curve):
        self.ctx = _HostDfContext(curve)


@pytest.fixture()
def host_df(monkeypatch):
    monkeypatch.setattr(position.CurveSession, "get", classmethod(lambda cls, curve, device=0: _Session(curve)))


def test_cashflow_rows_match_the_reference_engine(ref_curves, host_df):
    models = {}
    seen_error = 0
    for spec in golden():
        cv = ref_curves[spec["curve"]]
        model = models.setdefault(spec["curve"], build_model(cv))
        curve = model.curves[cv["name"]]
        swap = make_cal_trade(spec, cv)
        if "error" in spec:                                  # seasoned swap: the reference raises from the float leg
            with pytest.raises(LibError) as ex:
                CF.ois_cashflows(swap, curve)
            assert str(ex.value) in spec["error"]
            seen_error += 1
            continue
        assert [[d.d(), d.m(), d.y()] for d in swap._fixed_leg._payment_dts] == spec["fixed_payment_dts"]
        assert [[d.d(), d.m(), d.y()] for d in swap._float_leg._payment_dts] == spec["float_payment_dts"]
        cf = CF.ois_cashflows(swap, curve)
        assert_rows_match(cf, spec)
        assert cf.validate()
        assert repr(cf) == spec["repr"]
        first = cf.cashflows[0].to_dict()
        assert set(first) == set(spec["first_row_dict"]) and first["leg_type"] == spec["first_row_dict"]["leg_type"]
        d = cf.to_dict()
        assert d["count"] == len(cf) and d["currency"] == "GBP"
        assert cf.df.shape == (len(cf), 7)
    assert seen_error == 1


def test_cashflows_request_is_rejected_where_it_is_not_offered(ref_curves):
    from adrates_b200 import RequestTypes
    with pytest.raises(NotImplementedError):
        position.request_mask([RequestTypes.VALUE, RequestTypes.CASHFLOWS])
    assert position.request_mask([RequestTypes.VALUE, RequestTypes.CASHFLOWS], allow_cashflows=True) == position.request_mask([RequestTypes.VALUE])


def test_bond_and_frn_cashflow_rows_match_the_reference_engine(host_df):
    """Bonds (bullet, amortising, zero-coupon, seasoned, payment lag) and floating-rate notes (margins, first fixings, seasoned,
    dual-curve) of the bond / FRN goldens: the CASHFLOWS rows of Engine._compute_bond / _compute_frn, row assembly on the host
    with the DF test double (the GPU test runs the same goldens through cav_curve_df)."""
    from tests.util_bonds import build_bond_model, make_bond, make_frn
    from tests.util_cashflows import assert_credit_rows_match, credit_golden
    from adrates_b200.credit import BOND_CURVE
    g = credit_golden()
    model = build_bond_model(g)
    for rec in g["bonds"]:
        b = make_bond(rec)
        cf = CF.bond_cashflows(b, model.curves[b._floating_index.name])
        assert_credit_rows_match(cf, rec, rec["face"])
    n_err = 0
    for rec in g["frns"]:
        f = make_frn({**rec})
        disc = model.curves[BOND_CURVE[f._currency].name]
        idx = model.curves[f._floating_index.name]
        if "error" in rec:
            with pytest.raises(LibError) as ex:
                CF.frn_cashflows(f, disc, idx)
            assert str(ex.value) in rec["error"]
            n_err += 1
            continue
        assert_credit_rows_match(CF.frn_cashflows(f, disc, idx), rec, rec["face"])
    assert n_err == 1 and len(g["bonds"]) == 9 and len(g["frns"]) == 10


def test_yoy_cashflow_rows_match_the_reference_engine(host_df):
    """Year-on-year inflation swaps (engine.py:1355-1406): the reference values the swap on the non-AD path - sub-annual and
    seasoned swaps raise from the CPI look-up before the value date - and then reports the FIXED leg's rows only (its inflation
    leg loop is guarded by an attribute that leg never has).  14 cases run by the unmodified reference
    (tests/golden/gen/make_golden_cashflows_yoy.py), both index conventions."""
    from adrates_b200 import RequestTypes
    from adrates_b200.yoy_engine import compute_yoy
    from tests.conftest import load_golden
    from tests.util_cashflows import assert_rows_match
    from tests.util_yoy import make_model, make_swap
    g = load_golden("ref_yoy.json")
    cf_g = {c["id"]: c for c in load_golden("ref_cashflows_yoy.json")["cases"]}
    n_err = n_ok = 0
    for name in g["inflation_curves"]:
        model, idx, ic = make_model(g, name)
        disc = model.curves.GBP_OIS_SONIA
        for c in (c for c in g["cases"] if c["index"] == name):
            ref = cf_g[c["id"]]["cf_only"]
            assert ("error" in ref) == ("error" in cf_g[c["id"]]["value_cf"])
            swap = make_swap(c, idx)
            if "error" in ref:
                with pytest.raises(LibError) as ex:
                    CF.yoy_cashflows(swap, disc, ic)
                assert "LibError: " + str(ex.value) == ref["error"]
                with pytest.raises(LibError):
                    compute_yoy([swap], model, [RequestTypes.CASHFLOWS])
                n_err += 1
                continue
            cf = CF.yoy_cashflows(swap, disc, ic)
            assert {r["leg_type"] for r in ref["rows"]} == {"Fixed_Pay" if c["fixed_leg"] == "PAY" else "Fixed_Rec"}
            assert_rows_match(cf, dict(ref, notional=c["notional"]))
            assert repr(cf) == ref["repr"] and cf.validate()
            res = compute_yoy([swap], model, [RequestTypes.CASHFLOWS])       # nothing but the table: no device work besides DFs
            assert res.value is None and res.risk is None and res.gamma is None
            assert_rows_match(res.cashflows, dict(ref, notional=c["notional"]))
            n_ok += 1
    assert n_err == 8 and n_ok == 6
    with pytest.raises(NotImplementedError):
        compute_yoy([swap, swap], model, [RequestTypes.CASHFLOWS])
