"""Batched XCCY book (shared units per schedule, three ladders, per-trade rows expanded on the device) against the
per-trade path `compute_xccy`, which is pinned to the reference engine's outputs (test_gpu_parity.py)."""
import numpy as np
import pytest

from adrates_b200 import RequestTypes, CurveTypes
from adrates_b200.synthetic_xccy import make_xccy_book, XccyBookValuer
from adrates_b200.xccy_engine import compute_xccy
from tests.conftest import load_golden
from tests.util_xccy import build_xccy_model

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("max_offset_bd", [0, 7])
def test_xccy_book_matches_per_trade_path(max_offset_bd):
    g = load_golden("ref_xccy.json")
    m = build_xccy_model(g)
    book = make_xccy_book(m, 600, seed=11, max_offset_bd=max_offset_bd, spot=g["spot_fx"])
    val = XccyBookValuer(book)
    val.value()
    pv, d_dom, d_for, d_bas = val.results()
    rng = np.random.Generator(np.random.PCG64(3))
    for i in rng.choice(book.n_trades, 10, replace=False):
        sw = book.trade(int(i))
        ref = compute_xccy([sw], m, [RequestTypes.VALUE, RequestTypes.DELTA])
        N = book.dom_notional[i]
        T = float((sw._maturity_dt - sw._effective_dt) / 365.0) + 1.0
        assert abs(pv[i] - ref.value.amount) <= TOL * max(abs(ref.value.amount), N), i
        for got, ct in ((d_dom, CurveTypes.USD_OIS_SOFR), (d_for, CurveTypes.GBP_OIS_SONIA), (d_bas, CurveTypes.USD_GBP_BASIS)):
            lad = ref.risk(ct).risk_ladder
            assert np.max(np.abs(got[i, :len(lad)] - lad) / np.maximum(np.abs(lad), N * 1e-4 * T)) < TOL, (i, ct)
            assert not np.any(got[i, len(lad):])
    # portfolio totals of the batched path = sum of its per-trade rows
    val.sync()
    tot = val.agg[0].cpu().numpy()[0] + val.agg[1].cpu().numpy()[0]
    assert abs(tot - pv.sum()) <= 1e-11 * np.abs(pv).sum()
    assert np.max(np.abs(val.agg[2].cpu().numpy()[1:33] - d_bas.sum(0))) <= 1e-11 * np.abs(d_bas).sum()


def test_xccy_book_gammas_match_per_trade_path():
    """Per-trade gamma rows of the batched book (three per-curve blocks + the foreign x basis cross block) against
    compute_xccy([trade], GAMMA), whose blocks are validated by finite differences in tests/test_gpu_xccy_gamma.py."""
    g = load_golden("ref_xccy.json")
    m = build_xccy_model(g)
    book = make_xccy_book(m, 300, seed=5, max_offset_bd=5, spot=g["spot_fx"])
    val = XccyBookValuer(book, gamma=True)
    val.value()
    val.sync()
    nb = len(m.curves.GBP_USD_BASIS.basis_spreads)
    rng = np.random.Generator(np.random.PCG64(1))
    for i in rng.choice(book.n_trades, 6, replace=False):
        sw = book.trade(int(i))
        ref = compute_xccy([sw], m, [RequestTypes.GAMMA]).gamma
        pairs = ((val.gamma_dom, ref(CurveTypes.USD_OIS_SOFR).risk_ladder, 32, 32),
                 (val.gamma_for, ref(CurveTypes.GBP_OIS_SONIA).risk_ladder, 32, 32),
                 (val.gamma_basis, ref(CurveTypes.USD_GBP_BASIS).risk_ladder, nb, nb),
                 (val.gamma_cross, ref.cross_gamma(CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS).risk_matrix, 32, nb))
        N = book.dom_notional[i]
        T = float((sw._maturity_dt - sw._effective_dt) / 365.0) + 1.0
        for got, want, r, c in pairs:
            got = got[int(i)].cpu().numpy()[:r, :c]
            # natural magnitude of a gamma entry: notional x 1e-8 x T^2 (a par floating leg's own-curve gamma is exactly 0 on
            # the per-trade path and round-off of weighted unit rows on the batched one)
            assert np.max(np.abs(got - want)) <= 1e-10 * max(np.abs(want).max(), N * 1e-8 * T * T), i
