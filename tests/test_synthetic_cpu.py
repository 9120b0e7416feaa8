"""Synthetic book generator, its two flat layouts and the C oracle, cross-checked on CPU."""
import numpy as np

from oracle import cavour_oracle as orc
from oracle import c_oracle
from adrates_b200.curves import OISCurve
from adrates_b200.dates import Date
from adrates_b200.global_types import InterpTypes, SwapTypes
from adrates_b200.synthetic import make_book, flatten_book, reference_leg_tables
from tests.flat_eval import eval_flat
from tests.util_trades import METHOD, make_calibration_swaps, leg_arrays, rel_err

TOL = 1e-10


def _setup(ref_curves, n):
    cv = ref_curves["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    book = make_book(curve, n, seed=20240430)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    return cv, curve, book, (plan["times"], d, J, C)


def test_c_oracle_matches_python_oracle(ref_curves):
    cv, curve, book, tables = _setup(ref_curves, 40)
    leg_tabs = reference_leg_tables(book)
    trades = dict(sched=book.sched, coupon=book.coupon, notional=book.notional, spread=book.spread,
                  fixed_sign=book.fixed_sign)
    pv_d, dl_d, gm_d = c_oracle.ois_batch(tables, METHOD[cv["interp"]], leg_tabs, trades, dense=True)
    pv_s, dl_s, gm_s = c_oracle.ois_batch(tables, METHOD[cv["interp"]], leg_tabs, trades, dense=False, n_threads=2)
    vd = Date(*cv["value_dt"])
    from adrates_b200 import OIS, FrequencyTypes, DayCountTypes, CurveTypes, CurrencyTypes, BusDayAdjustTypes
    for i in range(book.n_trades):
        s = book.schedules[book.sched[i]]
        sw = OIS(s._effective_dt, s._termination_dt,
                 SwapTypes.RECEIVE if book.fixed_sign[i] > 0 else SwapTypes.PAY, float(book.coupon[i]),
                 FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurveTypes.GBP_OIS_SONIA, CurrencyTypes.GBP,
                 notional=float(book.notional[i]), float_freq_type=FrequencyTypes.ANNUAL,
                 float_dc_type=DayCountTypes.ACT_365F, bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        fixed, floating = leg_arrays(sw, vd)
        v, dl, gm = orc.ois_analytics(tables, METHOD[cv["interp"]], fixed, floating)
        N = book.notional[i]
        for pv_c, dl_c, gm_c in ((pv_d, dl_d, gm_d), (pv_s, dl_s, gm_s)):
            assert rel_err(pv_c[i], v, N) < TOL
            assert rel_err(dl_c[i], dl, N * 1e-4) < TOL
            assert rel_err(gm_c[i], gm, N * 1e-8) < TOL


def test_both_flat_layouts_match_c_oracle(ref_curves):
    cv, curve, book, tables = _setup(ref_curves, 300)
    leg_tabs = reference_leg_tables(book)
    trades = dict(sched=book.sched, coupon=book.coupon, notional=book.notional, spread=book.spread,
                  fixed_sign=book.fixed_sign)
    pv_c, dl_c, gm_c = c_oracle.ois_batch(tables, METHOD[cv["interp"]], leg_tabs, trades, dense=False)
    _, d, J, C = tables
    for dedup in (True, False):
        flat = flatten_book(book, dedup=dedup, max_group=4)
        pv, dl, gm = eval_flat(flat, d, J, C)
        N = book.notional
        assert np.max(np.abs(pv - pv_c) / np.maximum(np.abs(pv_c), N)) < TOL
        assert np.max(np.abs(dl - dl_c) / np.maximum(np.abs(dl_c), (N * 1e-4)[:, None])) < TOL
        assert np.max(np.abs(gm - gm_c) / np.maximum(np.abs(gm_c), (N * 1e-8)[:, None, None])) < TOL


def test_book_statistics(ref_curves):
    cv, curve, book, _ = _setup(ref_curves, 20000)
    assert 0.47 < np.mean(book.fixed_sign > 0) < 0.53
    assert book.coupon.min() >= 0.005 and book.coupon.max() <= 0.09
    assert 1e5 <= book.notional.min() and book.notional.max() <= 1e8
    flat = flatten_book(book, dedup=False)
    mean_terms = flat.n_terms / flat.n_trades
    assert 24.0 < mean_terms < 29.0        # (1 + mean of U{1..50}) merged cashflow times per trade
