"""GPU parity tests proper: the CUDA path, called through the C ABI, against the committed
golden vectors (unmodified reference engine) and the CPU oracle.  Tolerance: 1e-10 relative
(north_star), with the scale-aware metric of SURVEY R3."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import cavour_oracle as orc  # noqa: E402
from adrates_b200 import _native, RequestTypes  # noqa: E402
from adrates_b200.curves import OISCurve  # noqa: E402
from adrates_b200.dates import Date  # noqa: E402
from adrates_b200.flatten import Flattener  # noqa: E402
from adrates_b200.global_types import InterpTypes  # noqa: E402
from adrates_b200.position import Portfolio, CurveSession  # noqa: E402
from tests.conftest import GOLDEN  # noqa: E402
from tests.util_trades import (METHOD, build_model, leg_arrays, make_calibration_swaps, make_trade, rel_err,  # noqa: E402
                               trade_scales)

TOL = 1e-10
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
MASK = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA


def _curve(cv):
    vd, swaps = make_calibration_swaps(cv)
    return OISCurve(vd, swaps, InterpTypes[cv["interp"]])


def _run_flat(ctx, flat, mask=MASK):
    n = flat.n_trades
    pv = torch.zeros(n, dtype=torch.float64, device="cuda")
    dl = torch.zeros(n, 32, dtype=torch.float64, device="cuda")
    gm = torch.zeros(n, 32, 32, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.portfolio_upload(flat)
    agg = ctx.portfolio_value_host(mask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr())
    ctx.sync()
    return pv.cpu().numpy(), dl.cpu().numpy(), gm.cpu().numpy(), agg.copy()


@pytest.mark.parametrize("key", ["gbp_readme_lzr", "usd_dec24_lzr", "gbp_semi_lzr"])
def test_bootstrap_tables_match_reference_ad(ref_curves, key):
    cv = ref_curves[key]
    curve = _curve(cv)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    d, J, H = ctx.curve_read()
    ref = np.load(os.path.join(GOLDEN, f"ref_tables_{key}.npz"))
    assert np.max(np.abs(d - np.array(cv["pathB_dfs"]))) <= 4e-16
    assert rel_err(J, ref["jac"], 1.0) < 1e-12
    assert rel_err(H, ref["hess"], 1.0) < 1e-12
    ctx.close()


@pytest.mark.parametrize("dedup", [True, False])
def test_trades_match_reference_engine(ref_curves, ref_trades, dedup):
    for key, cv in ref_curves.items():
        specs = [s for s in ref_trades if s["curve"] == key]
        if not specs:
            continue
        curve = _curve(cv)
        ctx = _native.Context(0)
        ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
        fl = Flattener(curve)
        for s in specs:
            fl.add_trade(make_trade(s, cv))
        pv, dl, gm, agg = _run_flat(ctx, fl.finalize(dedup=dedup))
        for i, s in enumerate(specs):
            s_pv, s_d, s_g = trade_scales(s)
            e = (rel_err(pv[i], s["value"], s_pv), rel_err(dl[i], s["delta"], s_d), rel_err(gm[i], s["gamma"], s_g))
            assert max(e) < TOL, (s["id"], dedup, e)
            assert np.allclose(gm[i], gm[i].T, rtol=1e-10, atol=1e-14)
        # portfolio totals = sum of trades (Portfolio.compute semantics)
        tot = np.concatenate([[pv.sum()], dl.sum(0), gm.sum(0).reshape(-1)])
        scale = np.concatenate([[np.abs(pv).sum()], np.abs(dl).sum(0), np.abs(gm).sum(0).reshape(-1)]) + 1e-300
        assert np.max(np.abs(agg - tot) / scale) < 1e-12
        ctx.close()


def test_position_compute_api_matches_notebook_and_readme(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    nb = next(t for t in ref_trades if t["id"] == "nb_1w_par")
    res = make_trade(nb, cv).position(model).compute(ALL)
    assert abs(res.value.amount - 4.672529030358419e-11) < 1e-9        # notebook cell 36 (abs on 1e6 notional)
    assert abs(res.risk.ladder.data["1W"] - 1.9158970567491282) < 1e-12   # '1D'/'1W' pillars share the '1W' label
    assert abs(res.risk.risk_ladder[1] - 1.9158970567491282) < 1e-12   # cell 40
    assert abs(res.gamma.value.amount - (-7.34132e-06)) < 5e-12       # cell 44
    assert len(res.risk.ladder.data) == 31 and len(res.risk.tenors) == 32
    rd = next(t for t in ref_trades if t["id"] == "readme_10y")
    res = make_trade(rd, cv).position(model).compute(ALL)
    s_pv, s_d, s_g = trade_scales(rd)
    assert rel_err(res.value.amount, rd["value"], s_pv) < TOL
    assert rel_err(res.risk.risk_ladder, rd["delta"], s_d) < TOL
    assert rel_err(res.gamma.risk_ladder, rd["gamma"], s_g) < TOL
    assert res.risk.tenors == rd["delta_tenors"]
    only_v = make_trade(rd, cv).position(model).compute([RequestTypes.VALUE])
    assert only_v.risk is None and only_v.gamma is None and rel_err(only_v.value.amount, rd["value"], s_pv) < TOL


def test_portfolio_compute_sums_positions(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    specs = [s for s in ref_trades if s["curve"] == "gbp_readme_lzr"]
    port = Portfolio([make_trade(s, cv).position(model) for s in specs])
    res = port.compute(ALL)
    assert rel_err(res.value.amount, sum(s["value"] for s in specs), 1e7) < TOL
    assert rel_err(res.risk.risk_ladder, np.sum([s["delta"] for s in specs], axis=0), 1e4) < TOL
    assert rel_err(res.gamma.risk_ladder, np.sum([s["gamma"] for s in specs], axis=0), 1e1) < TOL


def test_portfolio_of_objects_uses_tile_plan_and_matches_oracle(ref_curves):
    """Portfolio.compute over 300 OIS objects on ~150 distinct schedules: the public path attaches a tile plan
    (tensor-core Greeks kernel) above position.TILE_MIN_UNITS; totals against the C oracle."""
    from oracle import c_oracle
    from adrates_b200 import OIS, SwapTypes, FrequencyTypes, DayCountTypes, CurveTypes, CurrencyTypes, BusDayAdjustTypes
    from adrates_b200 import position as pos_mod
    from adrates_b200.synthetic import make_book, reference_leg_tables
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    book = make_book(curve, 300, seed=21, max_offset_bd=40)
    swaps = []
    for i in range(book.n_trades):
        t = book.schedules[book.sched[i]]
        swaps.append(OIS(t._effective_dt, t._termination_dt, SwapTypes.RECEIVE if book.fixed_sign[i] > 0 else SwapTypes.PAY,
                         float(book.coupon[i]), FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurveTypes.GBP_OIS_SONIA,
                         CurrencyTypes.GBP, notional=float(book.notional[i]), float_freq_type=FrequencyTypes.ANNUAL,
                         float_dc_type=DayCountTypes.ACT_365F, bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING))
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    trades = dict(sched=book.sched, coupon=book.coupon, notional=book.notional, spread=book.spread, fixed_sign=book.fixed_sign)
    pv_c, dl_c, gm_c = c_oracle.ois_batch((plan["times"], d, J, C), METHOD[cv["interp"]], reference_leg_tables(book),
                                          trades, dense=False)
    results = []
    for min_units in (64, 10 ** 9):          # with and without the tile plan
        pos_mod.TILE_MIN_UNITS = min_units
        results.append(Portfolio([s.position(model) for s in swaps]).compute(ALL))
    pos_mod.TILE_MIN_UNITS = 64
    scale = np.abs(pv_c).sum()
    for res in results:
        assert abs(res.value.amount - pv_c.sum()) <= 1e-11 * scale
        assert np.max(np.abs(res.risk.risk_ladder - dl_c.sum(0))) <= 1e-11 * np.abs(dl_c).sum()
        assert np.max(np.abs(res.gamma.risk_ladder - gm_c.sum(0))) <= 1e-11 * np.abs(gm_c).sum()


def test_reference_property_tests_hold(ref_curves):
    """tests/test_refit_curves.py:152-231 (every calibration swap reprices to |PV| <= 1e-5 through
    Position.compute) and tests/test_ois_request_types.py:841-905 (PAY + RECEIVE = 0 within 1e-10)."""
    from adrates_b200 import (OIS, SwapTypes, FrequencyTypes, DayCountTypes, CurveTypes, CurrencyTypes,
                              BusDayAdjustTypes)
    for key in ["gbp_dec24_lzr", "gbp_semi_lzr", "gbp_quarterly_lzr", "usd_dec24_lzr", "gbp_readme_ff"]:
        cv = ref_curves[key]
        model = build_model(cv)
        vd = Date(*cv["value_dt"])
        dc, fq = DayCountTypes[cv["dc"]], FrequencyTypes[cv["freq"]]

        def mk(tenor, px, side):
            return OIS(effective_dt=vd, term_dt_or_tenor=tenor, fixed_leg_type=side, fixed_coupon=px / 100,
                       fixed_freq_type=fq, fixed_dc_type=dc, floating_index=CurveTypes[cv["name"]],
                       currency=CurrencyTypes[cv["name"][:3]], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                       float_freq_type=fq, float_dc_type=dc)
        for tenor, px in zip(cv["tenors"], cv["px"]):
            res = mk(tenor, px, SwapTypes.PAY).position(model).compute(ALL)
            assert abs(res.value.amount) <= 1e-5, (key, tenor, res.value.amount)
            g = res.gamma.risk_ladder
            assert g.shape == (32, 32) and np.allclose(g, g.T, rtol=1e-10, atol=1e-14)
        a = mk("10Y", 4.4, SwapTypes.PAY).position(model).compute(ALL)
        b = mk("10Y", 4.4, SwapTypes.RECEIVE).position(model).compute(ALL)
        assert abs(a.value.amount + b.value.amount) < 1e-10
        assert np.max(np.abs(a.risk.risk_ladder + b.risk.risk_ladder)) < 1e-10


def test_delta_gamma_explain_scenario_pnl(ref_curves):
    """tests/test_ois_request_types.py:429-474, 577-641: AD delta vs central finite difference through
    Model.scenario (1bp, rel err < 1e-4) and second-order Taylor on +-100bp."""
    from adrates_b200 import OIS, SwapTypes, FrequencyTypes, DayCountTypes, CurveTypes, CurrencyTypes, BusDayAdjustTypes
    cv = ref_curves["gbp_dec24_lzr"]
    model = build_model(cv)
    vd = Date(*cv["value_dt"])
    swap = OIS(vd, "10Y", SwapTypes.PAY, 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurveTypes.GBP_OIS_SONIA,
               CurrencyTypes.GBP, bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, float_freq_type=FrequencyTypes.ANNUAL,
               float_dc_type=DayCountTypes.ACT_365F)
    base = swap.position(model).compute(ALL)
    up = swap.position(model.scenario("GBP_OIS_SONIA", 0.01)).compute([RequestTypes.VALUE]).value.amount
    dn = swap.position(model.scenario("GBP_OIS_SONIA", -0.01)).compute([RequestTypes.VALUE]).value.amount
    fd = (up - dn) / 2.0
    assert abs(fd - base.risk.value.amount) / abs(fd) < 1e-4
    for shock_bp in (100.0, -100.0):
        pv_s = swap.position(model.scenario("GBP_OIS_SONIA", shock_bp / 100)).compute([RequestTypes.VALUE]).value.amount
        actual = pv_s - base.value.amount
        taylor1 = base.risk.value.amount * shock_bp
        taylor2 = taylor1 + 0.5 * base.gamma.value.amount * shock_bp ** 2
        assert abs(taylor2 - actual) < abs(taylor1 - actual)
        assert abs(taylor2 - actual) / abs(actual) < 0.05


def test_df_ad_matches_reference(ref_curves):
    for key, cv in ref_curves.items():
        model = build_model(cv)
        curve = getattr(model.curves, cv["name"])
        got = curve.df_ad(cv["df_ad_t"])
        assert rel_err(got, cv["df_ad"], 1.0) < 1e-13, key
        assert abs(float(curve.df_ad(5.0)) - cv["df_ad"][5]) < 1e-13
        # non-AD df() on path A (host) also pinned
        from adrates_b200.dates import DayCountTypes
        for dmy, ref in zip(cv["df_dates"], cv["df"]):
            assert abs(curve.df(Date(*dmy), DayCountTypes[cv["dc"]]) - ref) < 1e-14


def test_scenarios_match_oracle_rebootstrap(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
    specs = [s for s in ref_trades if s["curve"] == "gbp_readme_lzr"]
    vd = Date(*cv["value_dt"])
    swaps = [make_trade(s, cv) for s in specs]
    rng = np.random.default_rng(7)
    S = 5
    shocked = np.array(cv["swap_rates"])[None, :] + rng.normal(0, 1e-3, (S, 32))
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    for dedup in (True, False):
        fl = Flattener(curve)
        for sw in swaps:
            fl.add_trade(sw)
        flat = fl.finalize(dedup=dedup)
        ctx.portfolio_upload(flat)
        pnl = torch.zeros(S, flat.n_trades, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.scenarios(shocked, pnl.data_ptr())
        got = pnl.cpu().numpy()
        for s in range(S):
            d = orc.bootstrap_dfs(shocked[s], plan)
            for i, sw in enumerate(swaps):
                fixed, floating = leg_arrays(sw, vd)
                ref = orc.ois_value_only(plan["times"], d, METHOD[cv["interp"]], fixed, floating)
                assert abs(got[s, i] - ref) <= TOL * max(abs(ref), specs[i]["notional"]), (s, specs[i]["id"])
    ctx.close()


def test_scenario_df_cache_matches_direct_path(ref_curves, monkeypatch):
    """The DF-cache scenario path (one exp per distinct query and scenario) against the direct path (one exp per
    term and scenario) on a book with shared dates, and against the oracle's re-bootstrap for a few cells."""
    from adrates_b200.synthetic import make_book, flatten_book, shocked_rate_scenarios
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    book = make_book(curve, 1500, seed=5, max_offset_bd=30)
    flat = flatten_book(book, dedup=True)
    S = 37
    shocked = shocked_rate_scenarios(curve, S)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CAV_SCEN_DFCACHE", flag)
        ctx = _native.Context(0)
        ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
        ctx.portfolio_upload(flat)
        pnl = torch.zeros(S, flat.n_trades, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.scenarios(shocked, pnl.data_ptr())
        ctx.sync()
        out.append(pnl.cpu().numpy())
        ctx.close()
    assert np.array_equal(out[0], out[1])          # same arithmetic in the same order
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    vd = Date(*cv["value_dt"])
    for s_, i in ((0, 0), (11, 700), (36, 1499)):
        d = orc.bootstrap_dfs(shocked[s_], plan)
        tmpl = book.schedules[book.sched[i]]
        fixed, floating = leg_arrays(tmpl, vd)       # unit-notional, unit-coupon, PAY-fixed template
        fixed["payments"] = np.asarray(fixed["payments"]) * book.coupon[i] * book.notional[i]
        floating["notionals"] = np.asarray(floating["notionals"]) * book.notional[i]
        sign = -book.fixed_sign[i]                   # template pays fixed
        ref = sign * orc.ois_value_only(plan["times"], d, METHOD[cv["interp"]], fixed, floating)
        assert abs(out[1][s_, i] - ref) <= TOL * max(abs(ref), book.notional[i]), (s_, i)


def test_error_behaviour():
    ctx = _native.Context(0)
    from adrates_b200.error import LibError
    with pytest.raises(LibError):
        ctx.portfolio_value_host(MASK)          # no curve / portfolio yet
    ctx.close()


@pytest.mark.parametrize("dedup,compact", [(True, True), (False, True), (False, False)])
def test_synthetic_book_matches_c_oracle(ref_curves, dedup, compact):
    """3000 trades of the BASELINE config-2/3 book (homogeneous tiles: exercises the shared-row path of the
    tiled units kernel) against the C oracle, per trade and in total."""
    from oracle import c_oracle
    from adrates_b200.synthetic import make_book, flatten_book, reference_leg_tables
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    book = make_book(curve, 3000, seed=11, max_offset_bd=6)      # few offsets -> many trades per schedule
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    trades = dict(sched=book.sched, coupon=book.coupon, notional=book.notional, spread=book.spread,
                  fixed_sign=book.fixed_sign)
    pv_c, dl_c, gm_c = c_oracle.ois_batch((plan["times"], d, J, C), METHOD[cv["interp"]], reference_leg_tables(book),
                                          trades, dense=False)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    flat = flatten_book(book, dedup=dedup, compact=compact)
    assert flat.tile_plan is not None and (int(flat.tile_plan.tile_mask.min()) != 0xFFFFFFFF) == compact
    pv, dl, gm, agg = _run_flat(ctx, flat)
    N = book.notional
    assert np.max(np.abs(pv - pv_c) / np.maximum(np.abs(pv_c), N)) < TOL
    assert np.max(np.abs(dl - dl_c) / np.maximum(np.abs(dl_c), (N * 1e-4)[:, None])) < TOL
    assert np.max(np.abs(gm.reshape(-1, 32, 32) - gm_c) / np.maximum(np.abs(gm_c), (N * 1e-8)[:, None, None])) < TOL
    tot = np.concatenate([[pv_c.sum()], dl_c.sum(0), gm_c.sum(0).reshape(-1)])
    scale = np.concatenate([[np.abs(pv_c).sum()], np.abs(dl_c).sum(0), np.abs(gm_c).sum(0).reshape(-1)]) + 1e-300
    assert np.max(np.abs(agg - tot) / scale) < 1e-11
    ctx.close()


def test_tile_mask_too_small_is_rejected(ref_curves):
    """A tile plan whose active-pillar masks miss part of the tables' support must fail loudly, not drop Greeks."""
    from adrates_b200.error import LibError
    from adrates_b200.synthetic import make_book, flatten_book
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    flat = flatten_book(make_book(curve, 200, seed=5, max_offset_bd=6), dedup=False)
    # drop the highest active pillar of the first tile (it stays in the smallest size class: tiles remain ordered)
    m0 = int(flat.tile_plan.tile_mask[0])
    flat.tile_plan.tile_mask[0] = np.uint32(m0 & ~(1 << (m0.bit_length() - 1)))
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    with pytest.raises(LibError, match="active-pillar masks"):
        _run_flat(ctx, flat)
    ctx.close()


def test_malformed_tile_plans_are_rejected(ref_curves):
    """cav_portfolio_set_tiles checks everything the tiled kernel dereferences; each defect has its own message."""
    import copy
    from adrates_b200.error import LibError
    from adrates_b200.synthetic import make_book, flatten_book
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    flat = flatten_book(make_book(curve, 4000, seed=6), dedup=True)
    tp0 = flat.tile_plan
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    flat_nt = copy.copy(flat)
    flat_nt.tile_plan = None
    ctx.portfolio_upload(flat_nt)

    def broken(field, fn):
        tp = copy.copy(tp0)
        a = np.array(getattr(tp0, field), copy=True)
        fn(a)
        setattr(tp, field, a)
        return tp

    lens = np.diff(flat.unit_offsets)
    full = int(np.flatnonzero(np.asarray(tp0.tile_units).reshape(-1, 16)[:, 1] >= 0)[0])   # a tile with >= 2 units
    first = int(tp0.tile_units[16 * full])
    other = int(np.flatnonzero(lens != lens[first])[0])          # a unit with another number of terms
    cases = [
        (broken("tile_units", lambda a: a.__setitem__(0, flat.n_units)), "unit id out of range"),
        (broken("tile_units", lambda a: a.__setitem__(0, -2)), "unit id out of range"),
        (broken("tile_units", lambda a: a.__setitem__(0, -1)), "exactly one tile"),
        (broken("tile_units", lambda a: a.__setitem__(16 * full + 1, other)), "differ in length"),
        (broken("tile_units", lambda a: a.__setitem__(16 * full + 1, first)), "exactly one tile"),     # listed twice, one omitted
        (broken("tile_kcount", lambda a: a.__setitem__(len(a) - 1, len(tp0.k_row) + 1)), "K range out of bounds"),
        (broken("tile_kstart", lambda a: a.__setitem__(0, -1)), "K range out of bounds"),
        (broken("k_row", lambda a: a.__setitem__(0, 10 ** 6)), "bad K row"),
        (broken("k_coef", lambda a: a.__setitem__(0, 6)), "bad K row"),
        (broken("perm", lambda a: a.__setitem__(0, a[1])), "not a permutation"),
    ]
    for tp, msg in cases:
        with pytest.raises(LibError, match=msg):
            ctx.portfolio_set_tiles(tp)
    ctx.portfolio_set_tiles(tp0)                                  # the context still takes the good plan
    agg = ctx.portfolio_value_host(MASK)
    ctx.portfolio_upload(flat_nt)                                 # ... which gives what the plan-less kernel gives
    agg0 = ctx.portfolio_value_host(MASK)
    scale = np.abs(agg0) + 1e-6 * np.abs(agg0).max()
    assert np.max(np.abs(agg - agg0) / scale) < 1e-10
    ctx.close()


def test_delta_chain_gemm_matches_fused_path(ref_curves):
    """The DMMA chain-rule GEMM (delta = Q * 1e-4 J/d) against the fused per-cashflow chain, both layouts."""
    from adrates_b200.synthetic import make_book, flatten_book
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    book = make_book(curve, 2500, seed=3)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    for flat in (flatten_book(book, dedup=True), flatten_book(book, dedup=False, sort_units=False)):
        pv, dl, gm, agg = _run_flat(ctx, flat, _native.REQ_VALUE | _native.REQ_DELTA)
        pv2 = torch.zeros(flat.n_trades, dtype=torch.float64, device="cuda")
        dl2 = torch.zeros(flat.n_trades, 32, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ms, flops = ctx.portfolio_delta_gemm(pv2.data_ptr(), dl2.data_ptr())
        ctx.sync()
        N = book.notional
        assert flops == 2.0 * flat.n_units * 272 * 32 and ms > 0
        assert np.max(np.abs(pv2.cpu().numpy() - pv) / np.maximum(np.abs(pv), N)) < TOL
        assert np.max(np.abs(dl2.cpu().numpy() - dl) / np.maximum(np.abs(dl), (N * 1e-4)[:, None])) < TOL
    ctx.close()


def test_xccy_position_and_portfolio_match_reference():
    """Engine._compute_xccy VALUE + three delta ladders (goldens from the unmodified reference engine)."""
    from tests.conftest import load_golden
    from tests.util_xccy import build_xccy_model, make_xccy_trade
    from adrates_b200 import Position, CurveTypes
    g = load_golden("ref_xccy.json")
    m = build_xccy_model(g)
    trades = [make_xccy_trade(t) for t in g["trades"]]
    tot_v, tot_b = 0.0, None
    for sw, t in zip(trades, g["trades"]):
        res = Position(sw, m).compute([RequestTypes.VALUE, RequestTypes.DELTA])
        N, T = t["domestic_notional"], float(t["tenor"][:-1])
        assert abs(res.value.amount - t["value"]) <= TOL * max(abs(t["value"]), N), t["id"]
        for key, ct in (("USD_OIS_SOFR", CurveTypes.USD_OIS_SOFR), ("GBP_OIS_SONIA", CurveTypes.GBP_OIS_SONIA),
                        ("USD_GBP_BASIS", CurveTypes.USD_GBP_BASIS)):
            ref = np.array(t["deltas"][key]["ladder"])
            got = res.risk(ct).risk_ladder
            assert got.shape == ref.shape and res.risk(ct).tenors == t["deltas"][key]["tenors"]
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), N * 1e-4 * T)) < TOL, (t["id"], key)
        tot_v += t["value"]
        b = np.array(t["deltas"]["USD_GBP_BASIS"]["ladder"])
        tot_b = b if tot_b is None else tot_b + b
        only_v = Position(sw, m).compute([RequestTypes.VALUE])
        assert only_v.risk is None and abs(only_v.value.amount - t["value"]) <= TOL * max(abs(t["value"]), N)
    port = Portfolio([Position(sw, m) for sw in trades]).compute([RequestTypes.VALUE, RequestTypes.DELTA])
    assert abs(port.value.amount - tot_v) <= TOL * 1e8
    assert np.max(np.abs(port.risk.USD_GBP_BASIS.risk_ladder - tot_b)) <= TOL * 1e8 * 1e-4 * 10
    # GAMMA: three per-curve blocks and the foreign x basis cross block (validated in tests/test_gpu_xccy_gamma.py); the sum of
    # the per-trade results is the portfolio's
    one = [Position(sw, m).compute([RequestTypes.GAMMA]).gamma for sw in trades]
    both = Portfolio([Position(sw, m) for sw in trades]).compute([RequestTypes.GAMMA]).gamma
    for ct in (CurveTypes.USD_OIS_SOFR, CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS):
        tot = sum(g(ct).risk_ladder for g in one)
        assert np.max(np.abs(both(ct).risk_ladder - tot)) <= 1e-12 * np.abs(tot).max()
    with pytest.raises(NotImplementedError):
        Position(trades[0], m).compute([RequestTypes.CASHFLOWS])


def test_small_curve_seasoned_swap_and_empty_portfolio():
    """Edge cases against the pinned oracle: a 6-pillar curve (ladders narrower than a warp), a seasoned swap whose
    first accrual started before the value date (past cashflows masked, DF(t<0) = 1), and an empty portfolio."""
    from adrates_b200 import (Model, OIS, Date, SwapTypes, FrequencyTypes, DayCountTypes, CurveTypes, CurrencyTypes,
                              BusDayAdjustTypes, InterpTypes)
    from adrates_b200.flatten import FlatPortfolio
    vd = Date(17, 12, 2024)
    m = Model(vd)
    tenors, px = ["6M", "1Y", "2Y", "5Y", "10Y", "30Y"], [5.1, 5.0, 4.7, 4.3, 4.1, 4.0]
    m.build_curve(name="GBP_OIS_SONIA", px_list=px, tenor_list=tenors, fixed_dcc_type=DayCountTypes.ACT_365F,
                  float_dc_type=DayCountTypes.ACT_365F, interp_type=InterpTypes.FLAT_FWD_RATES)
    curve = m.curves.GBP_OIS_SONIA
    plan = orc.plan_path_b(curve.swap_times, curve.year_fracs)
    d, J, C = orc.bootstrap_tables(curve.swap_rates, plan)
    for eff, tenor in ((Date(17, 12, 2024), "7Y"), (Date(20, 3, 2023), "6Y"), (Date(3, 2, 2025), "18M")):
        sw = OIS(eff, tenor, SwapTypes.RECEIVE, 0.043, FrequencyTypes.SEMI_ANNUAL, DayCountTypes.ACT_365F,
                 CurveTypes.GBP_OIS_SONIA, CurrencyTypes.GBP, notional=3e6, float_spread=0.001,
                 float_freq_type=FrequencyTypes.QUARTERLY, float_dc_type=DayCountTypes.ACT_365F,
                 bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
        res = sw.position(m).compute(ALL)
        fixed, floating = leg_arrays(sw, vd)
        v, dl, gm = orc.ois_analytics((plan["times"], d, J, C), 1, fixed, floating)
        assert res.risk.risk_ladder.shape == (6,) and res.gamma.risk_ladder.shape == (6, 6)
        assert rel_err(res.value.amount, v, 3e6) < TOL
        assert rel_err(res.risk.risk_ladder, dl, 3e6 * 1e-4) < TOL
        assert rel_err(res.gamma.risk_ladder, gm, 3e6 * 1e-8) < TOL
    ctx = CurveSession.get(curve).ctx
    z64, zf, zi = np.zeros(1, dtype=np.int64), np.zeros(0), np.zeros(0, dtype=np.int32)
    ctx.portfolio_upload(FlatPortfolio(0, 0, z64, 2, zf, zf, zi, 0, 1, zf, 0, z64, zi, None, zf))
    assert np.all(ctx.portfolio_value_host(MASK) == 0.0)
    assert Portfolio([]).compute(ALL).value is None


def test_upload_rejects_bad_indices(ref_curves):
    from adrates_b200.error import LibError
    from adrates_b200.synthetic import make_book, flatten_book
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=1)
    flat = flatten_book(make_book(curve, 50, seed=1), dedup=True)
    flat.node = flat.node.copy()
    flat.node[3] = 10_000                      # beyond the curve grid
    with pytest.raises(LibError, match="node index out of range"):
        ctx.portfolio_upload(flat)
    with pytest.raises(LibError):
        ctx.portfolio_value_host(_native.REQ_VALUE)     # the rejected portfolio must not be valued
    ctx.close()


def test_ois_with_cross_currency_collateral_matches_reference():
    from tests.conftest import load_golden
    from tests.util_xccy import build_xccy_model
    from tests.util_trades import make_trade
    from adrates_b200 import CollateralType, CurveTypes
    g, cg = load_golden("ref_xccy.json"), load_golden("ref_collateral.json")
    m = build_xccy_model(g, xccy_name="GBP_USD_XCCY")
    fx = m.curves.GBP_USD_XCCY._spot_fx
    for t in cg["trades"]:
        sw = make_trade(dict(t, payment_lag=0), {"name": "GBP_OIS_SONIA", "dc": "ACT_365F"})
        res = sw.position(m).compute([RequestTypes.VALUE, RequestTypes.DELTA], collateral_type=CollateralType.USD)
        N, T = t["notional"] / fx, float(t["tenor"][:-1])
        assert res.value.currency.name == "USD"
        assert abs(res.value.amount - t["value"]) <= TOL * max(abs(t["value"]), N), t["id"]
        for key, ct in (("GBP_OIS_SONIA", CurveTypes.GBP_OIS_SONIA), ("USD_GBP_BASIS", CurveTypes.USD_GBP_BASIS)):
            r = np.array(t["deltas"][key]["ladder"])
            got = res.risk(ct).risk_ladder
            assert np.max(np.abs(got - r) / np.maximum(np.abs(r), N * 1e-4 * T)) < TOL, (t["id"], key)
        same = sw.position(m).compute([RequestTypes.VALUE], collateral_type=CollateralType.GBP)   # natural currency
        assert same.value.currency.name == "GBP"
        with pytest.raises(NotImplementedError):
            sw.position(m).compute([RequestTypes.GAMMA], collateral_type=CollateralType.USD)


def test_curve_rebuild_from_device_rates(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    shocked = np.array(curve.swap_rates) + 1e-3
    rd = torch.tensor(shocked, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.curve_rebuild_dev(rd.data_ptr())
    d1, J1, H1 = ctx.curve_read()
    ctx.curve_build(curve._interp_type.value, shocked, curve.path_b_plan(), order=2)
    d2, J2, H2 = ctx.curve_read()
    assert np.array_equal(d1, d2) and np.array_equal(J1, J2) and np.array_equal(H1, H2)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    assert np.max(np.abs(d1 - orc.bootstrap_dfs(shocked, plan))) < 4e-16
    ctx.close()
