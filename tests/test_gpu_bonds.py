"""Bonds through Position.compute / Portfolio.compute on the GPU against Engine._compute_bond of the unmodified
reference (tests/golden/ref_bonds.json)."""
import numpy as np
import pytest

from adrates_b200 import Portfolio, RequestTypes
from tests.conftest import load_golden
from tests.util_bonds import build_bond_model, make_bond, make_frn

pytestmark = pytest.mark.gpu
TOL = 1e-10
REQ = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]


def _scales(b):
    T = max(len(b["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[b["freq"]], 1.0)
    return b["face"], b["face"] * 1e-4 * T, b["face"] * 1e-8 * T * T


def test_bond_positions_match_reference():
    g = load_golden("ref_bonds.json")
    m = build_bond_model(g)
    for b in g["bonds"]:
        res = make_bond(b).position(m).compute(REQ)
        s_pv, s_d, s_g = _scales(b)
        assert abs(res.value.amount - b["value"]) <= TOL * max(abs(b["value"]), s_pv), b["id"]
        assert res.risk.tenors == b["tenors"] and res.value.currency.name == b["currency"]
        ref_d, ref_g = np.array(b["delta"]), np.array(b["gamma"])
        assert np.max(np.abs(res.risk.risk_ladder - ref_d) / np.maximum(np.abs(ref_d), s_d)) < TOL, b["id"]
        assert np.max(np.abs(res.gamma.risk_ladder - ref_g) / np.maximum(np.abs(ref_g), s_g)) < TOL, b["id"]
        only_v = make_bond(b).position(m).compute([RequestTypes.VALUE])
        assert only_v.risk is None and only_v.gamma is None


def test_bond_portfolio_is_one_batched_valuation():
    g = load_golden("ref_bonds.json")
    m = build_bond_model(g)
    gbp = [b for b in g["bonds"] if b["currency"] == "GBP"]
    res = Portfolio([make_bond(b).position(m) for b in gbp]).compute(REQ)
    v = sum(b["value"] for b in gbp)
    d = np.sum([b["delta"] for b in gbp], axis=0)
    G = np.sum([b["gamma"] for b in gbp], axis=0)
    scale = sum(abs(b["value"]) for b in gbp)
    assert abs(res.value.amount - v) <= TOL * scale
    assert np.max(np.abs(res.risk.risk_ladder - d)) <= TOL * scale * 1e-4 * 30
    assert np.max(np.abs(res.gamma.risk_ladder - G)) <= TOL * scale * 1e-8 * 900


def test_frn_positions_match_reference():
    """Engine._compute_frn, single-curve case, incl. known first fixing, payment lag (product terms) and a seasoned note."""
    from adrates_b200 import LibError, FRN, Date, FrequencyTypes, DayCountTypes, CurrencyTypes, CurveTypes
    g = load_golden("ref_frn.json")
    m = build_bond_model(g)
    for f in g["frns"]:
        res = make_frn(f).position(m).compute(REQ)
        T = max(len(f["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[f["freq"]], 1.0)
        N = f["face"]
        assert abs(res.value.amount - f["value"]) <= TOL * max(abs(f["value"]), N), f["id"]
        ref_d, ref_g = np.array(f["delta"]), np.array(f["gamma"])
        assert np.max(np.abs(res.risk.risk_ladder - ref_d) / np.maximum(np.abs(ref_d), N * 1e-4 * T)) < TOL, f["id"]
        assert np.max(np.abs(res.gamma.risk_ladder - ref_g) / np.maximum(np.abs(ref_g), N * 1e-8 * T * T)) < TOL, f["id"]
    dual = FRN(Date(30, 4, 2024), "2Y", 0.003, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP,
               CurveTypes.USD_OIS_SOFR)
    with pytest.raises(LibError, match="Dual-curve FRN"):
        dual.position(m).compute(REQ)


def test_dual_curve_frn_value_matches_reference():
    """Index curve != discount curve: VALUE only (DELTA / GAMMA raise like the reference).  With identical pillar dates
    the reference's curve cache resolves the index curve to the discount curve (single-curve value); a SOFR curve
    quoting fewer pillars exercises the genuine two-grid product terms."""
    from adrates_b200 import LibError, Date, DayCountTypes, FrequencyTypes, BusDayAdjustTypes, SwapTypes, InterpTypes
    from adrates_b200.models import Model
    g = load_golden("ref_frn.json")
    m = build_bond_model(g)
    for f in g["dual"]:
        note = make_frn(f)
        res = note.position(m).compute([RequestTypes.VALUE])
        assert abs(res.value.amount - f["value"]) <= 1e-10 * f["face"], f["id"]
        assert res.risk is None and res.gamma is None
        with pytest.raises(LibError, match="Dual-curve FRN delta/gamma not yet implemented"):
            note.position(m).compute([RequestTypes.VALUE, RequestTypes.DELTA])
    m2 = Model(Date(*g["value_dt"]))
    keep = g["dual_distinct_usd_keep"]
    for name, px, tn in (("GBP_OIS_SONIA", g["gbp_px"], g["tenors"]),
                         ("USD_OIS_SOFR", [g["usd_px"][i] for i in keep], [g["tenors"][i] for i in keep])):
        m2.build_curve(name=name, px_list=px, tenor_list=tn, spot_days=0, swap_type=SwapTypes.PAY,
                       fixed_dcc_type=DayCountTypes.ACT_365F, fixed_freq_type=FrequencyTypes.ANNUAL,
                       float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=DayCountTypes.ACT_365F,
                       bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes.LINEAR_ZERO_RATES)
    for f in g["dual_distinct"]:
        res = make_frn(f).position(m2).compute([RequestTypes.VALUE])
        assert abs(res.value.amount - f["value"]) <= 1e-10 * f["face"], f["id"]


def test_portfolio_with_dual_curve_frn_equals_sum_of_positions():
    """Portfolio.compute == sum of Position.compute (reference portfolio.py:39-67) also when a book holds a dual-curve FRN:
    the note is valued on index + discount curves (VALUE), never as a single-curve unit on its index curve, and Greeks
    raise like Position.compute does."""
    from adrates_b200 import LibError
    from adrates_b200.position import Portfolio
    g = load_golden("ref_frn.json")
    m = build_bond_model(g)
    specs = [f for f in g["dual"] + g["frns"] if f["currency"] == "GBP"][:6]         # sterling notes: dual- and single-curve
    assert any(f["index"] != "GBP_OIS_SONIA" for f in specs) and any(f["index"] == "GBP_OIS_SONIA" for f in specs)
    pos = [make_frn(f).position(m) for f in specs]
    total = Portfolio(pos).compute([RequestTypes.VALUE])
    each = sum(p.compute([RequestTypes.VALUE]).value.amount for p in pos)
    assert abs(total.value.amount - each) <= 1e-10 * sum(f["face"] for f in specs)
    assert abs(total.value.amount - sum(f["value"] for f in specs)) <= 1e-10 * sum(f["face"] for f in specs)
    with pytest.raises(LibError, match="Dual-curve FRN delta/gamma not yet implemented"):
        Portfolio(pos).compute([RequestTypes.VALUE, RequestTypes.DELTA])
