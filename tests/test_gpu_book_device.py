"""Device-side book flattener (cav_book_from_arrays) against the host flattener batch.OISBook.flatten + tiles.plan_tiles:
flat arrays bit for bit (random books, every convention the device path accepts, and the reference's own 1750 schedules),
the tile plan entry by entry, per-trade results bit-identical through either path, the reference's error messages, and the
BASELINE-size book (1M trades)."""
import time

import numpy as np
import pytest

from adrates_b200 import RequestTypes, _native
from adrates_b200 import batch as B
from adrates_b200.curves import OISCurve
from adrates_b200.dates import (BusDayAdjustTypes, CalendarTypes, Date, DateGenRuleTypes, DayCountTypes, FrequencyTypes)
from adrates_b200.error import LibError
from adrates_b200.global_types import InterpTypes
from adrates_b200.position import CurveSession
from adrates_b200.tiles import node_support_masks, plan_tiles
from tests.test_book_core_cpu import CONVS, _random_book
from tests.util_trades import make_calibration_swaps

pytestmark = pytest.mark.gpu
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]


def _curve(cv):
    vd, swaps = make_calibration_swaps(cv)
    return OISCurve(vd, swaps, InterpTypes[cv["interp"]])


def _assert_flat_equal(got, ref, exact_weights=False):
    assert (got.n_units, got.n_terms, got.n_trades, got.n_groups, got.n_pairs, got.n_comp) == \
           (ref.n_units, ref.n_terms, ref.n_trades, ref.n_groups, ref.n_pairs, ref.n_comp)
    for name in ("unit_offsets", "amt", "weight", "node", "comp_weight", "group_offsets", "group_units", "out_index"):
        assert np.array_equal(getattr(got, name), getattr(ref, name)), name
    if exact_weights:
        assert np.array_equal(got.unit_weight, ref.unit_weight)
    else:   # the device sums a unit's trade weights lane-strided + butterfly (fixed order), numpy's bincount sequentially
        scale = np.zeros(ref.n_units)
        ids = np.repeat(ref.group_units.reshape(-1, ref.n_comp), np.diff(ref.group_offsets), axis=0)
        for k in range(ref.n_comp):
            scale += np.bincount(ids[:, k], weights=np.abs(ref.comp_weight.reshape(-1, ref.n_comp)[:, k]), minlength=ref.n_units)
        assert np.all(np.abs(got.unit_weight - ref.unit_weight) <= 1e-14 * np.maximum(scale, 1.0))


def _assert_tiles_equal(got, ref, G):
    assert got["n_tiles"] == ref.n_tiles
    assert np.array_equal(got["perm"], ref.perm)
    for name in ("tile_units", "tile_kcount", "tile_kstart", "tile_npos", "tile_mask"):
        assert np.array_equal(got[name], getattr(ref, name)), name

    def canon(rows, pairs):
        rows = np.asarray(rows).astype(np.int64)
        pr = np.asarray(pairs).reshape(-1, 2)
        is_pair = rows >= 3 * G
        out = rows.copy()
        out[is_pair] = 10 * G + pr[rows[is_pair] - 3 * G, 0] * 4096 + pr[rows[is_pair] - 3 * G, 1]
        return out
    assert np.array_equal(canon(got["k_row"], got["pairs"]), canon(ref.k_row, ref.pairs))
    two = ref.k_coef2 >= 0
    desc = ref.k_pos | (ref.k_coef << 8) | (np.where(two, ref.k_pos2, 0) << 16) | (np.where(two, ref.k_coef2, 7) << 24)
    assert np.array_equal(got["k_desc"], desc)
    cb = got["class_begin"]
    assert cb[0] == 0 and cb[-1] == ref.n_tiles and np.all(np.diff(cb) >= 0)


@pytest.mark.parametrize("conv", list(CONVS))
@pytest.mark.parametrize("spread", [False, True])
@pytest.mark.parametrize("name", ["gbp_readme_lzr", "gbp_semi_lzr"])
def test_device_flatten_equals_host_flatten(ref_curves, conv, spread, name):
    curve = _curve(ref_curves[name])
    rng = np.random.default_rng(31)
    book = B.OISBook.from_arrays(curve, **_random_book(curve, 3000, rng, spread), **CONVS[conv])
    sess = CurveSession.get(curve, 0)
    assert book.upload(sess.ctx, tiles=True) == "device"
    info = sess.ctx.book_info()
    assert info["built"] == 1
    ref = book.flatten(dedup=True, tiles=True)
    _assert_flat_equal(sess.ctx.book_read(), ref)
    assert ref.tile_plan is not None and info["n_tiles"] == ref.tile_plan.n_tiles
    _assert_tiles_equal(sess.ctx.book_read_tiles(), ref.tile_plan, curve.path_b_plan().n_nodes)


def test_device_flatten_on_the_reference_schedules(ref_curves, ref_schedules):
    """The (effective, termination) pairs of the 1750 schedules the unmodified reference generated, one book per
    (frequency, roll convention, generation rule): the device-built flat arrays equal the host-built ones bit for bit;
    pairs the reference rejects are rejected with its message."""
    curve = _curve(ref_curves["gbp_readme_lzr"])
    sess = CurveSession.get(curve, 0)
    classes = {}
    for r in ref_schedules["schedules"]:
        classes.setdefault((r["freq"], r["bd"], r["dg"]), []).append(r)
    n_ok = n_bad = 0
    for (freq, bd, dg), rs in classes.items():
        conv = dict(fixed_freq_type=FrequencyTypes[freq], fixed_dc_type=DayCountTypes.ACT_365F, float_freq_type=FrequencyTypes[freq],
                    float_dc_type=DayCountTypes.ACT_360, bd_type=BusDayAdjustTypes[bd], dg_type=DateGenRuleTypes[dg])
        eff = np.array([Date(*r["eff"])._n for r in rs], dtype=np.int64)
        term = np.array([Date(*r["term"])._n for r in rs], dtype=np.int64)
        good = np.ones(len(rs), dtype=bool)
        for i in range(len(rs)):                    # the host flattener decides which pairs make a valid OIS
            try:
                B.OISBook.from_arrays(curve, eff[i:i + 1], termination=term[i:i + 1], fixed_sign=1.0, fixed_coupon=0.03,
                                      **conv).flatten(tiles=False)
            except LibError as ex:
                good[i] = False
                one = B.OISBook(curve, eff[i:i + 1], term[i:i + 1], np.ones(1), np.full(1, 0.03), np.full(1, 1e6), **conv)
                with pytest.raises(LibError) as dev_ex:
                    one.upload(sess.ctx, tiles=False)
                assert str(dev_ex.value) == str(ex), (rs[i], str(dev_ex.value), str(ex))
                n_bad += 1
        k = int(good.sum())
        rng = np.random.default_rng(k)
        book = B.OISBook.from_arrays(curve, eff[good], termination=term[good], fixed_sign=np.where(rng.random(k) < 0.5, 1.0, -1.0),
                                     fixed_coupon=rng.uniform(0.01, 0.06, k), notional=rng.uniform(1e5, 1e7, k),
                                     float_spread=rng.normal(0, 1e-3, k), **conv)
        assert book.upload(sess.ctx, tiles=True) == "device"
        ref = book.flatten(dedup=True, tiles=True)
        _assert_flat_equal(sess.ctx.book_read(), ref)
        if ref.tile_plan is not None:
            _assert_tiles_equal(sess.ctx.book_read_tiles(), ref.tile_plan, curve.path_b_plan().n_nodes)
        n_ok += k
    assert n_ok + n_bad == len(ref_schedules["schedules"]) >= 1750 and n_ok > 1500


def test_results_are_bit_identical_through_either_flattener(ref_curves):
    import torch
    curve = _curve(ref_curves["gbp_readme_lzr"])
    rng = np.random.default_rng(7)
    book = B.OISBook.from_arrays(curve, **_random_book(curve, 5000, rng, True), **CONVS["annual_act365"])
    res_d, rows_d = book.compute(ALL)
    res_h, rows_h = book.compute(ALL, device_flatten=False)
    for k in ("pv", "delta", "gamma"):
        assert torch.equal(rows_d[k], rows_h[k]), k
    # totals: sums of unit results weighted by unit_weight, whose summation order differs between the two flatteners
    gross = float(np.sum(np.abs(book.notional)))
    assert abs(res_d.value.amount - res_h.value.amount) <= 1e-13 * gross
    assert np.max(np.abs(res_d.risk.risk_ladder - res_h.risk.risk_ladder)) <= 1e-13 * gross * 1e-4 * 40
    assert np.max(np.abs(res_d.gamma.risk_ladder - res_h.gamma.risk_ladder)) <= 1e-13 * gross * 1e-8 * 1600
    # PV + delta only (no tile plan), and scenario values through the device-built book
    r2, rows2 = book.compute([RequestTypes.VALUE, RequestTypes.DELTA])
    r2h, rows2h = book.compute([RequestTypes.VALUE, RequestTypes.DELTA], device_flatten=False)
    assert torch.equal(rows2["pv"], rows2h["pv"]) and torch.equal(rows2["delta"], rows2h["delta"])
    assert np.max(np.abs(r2.risk.risk_ladder - r2h.risk.risk_ladder)) <= 1e-13 * gross * 1e-4 * 40
    rates = np.array(curve.swap_rates)[None, :] + rng.normal(0, 1e-3, (8, len(curve.swap_rates)))
    sv = book.scenario_values(rates)
    from adrates_b200.scenarios import scenario_values_flat
    assert torch.equal(sv, scenario_values_flat(curve, book.flatten(tiles=False), rates))


def test_device_flatten_errors_and_fallbacks(ref_curves):
    curve = _curve(ref_curves["gbp_readme_lzr"])
    sess = CurveSession.get(curve, 0)
    vd = curve._value_dt._n
    n = 6000                                        # above EAGER_CHECK_MAX: validated by the device flattener
    eff = np.full(n, vd, dtype=np.int64)
    term = eff + 365
    term[4321] = vd - 30
    book = B.OISBook.from_arrays(curve, eff, termination=term, fixed_sign=1.0, fixed_coupon=0.03)
    with pytest.raises(LibError, match="Start date after maturity date"):
        book.compute([RequestTypes.VALUE])
    # a payment lag, a fully matured book and private units are flattened on the host (same call, same results)
    lag = B.OISBook.from_arrays(curve, eff, tenor_years=5, fixed_sign=1.0, fixed_coupon=0.03, payment_lag=2)
    assert lag.device_conv() is None and lag.upload(sess.ctx, tiles=False) == "host"
    old = B.OISBook.from_arrays(curve, np.full(5000, vd - 4000), tenor_years=2, fixed_sign=1.0, fixed_coupon=0.03)
    assert old.upload(sess.ctx, tiles=False) == "host"
    assert old.compute([RequestTypes.VALUE])[0].value.amount == 0.0
    ok = B.OISBook.from_arrays(curve, eff, tenor_years=5, fixed_sign=1.0, fixed_coupon=0.03)
    assert ok.upload(sess.ctx, dedup=False) == "host"
    # the plan of a device-built book is the library's own: a host plan on top of it is refused
    assert ok.upload(sess.ctx) == "device"
    with pytest.raises(LibError, match="flattened on the device"):
        sess.ctx.portfolio_set_tiles(ok.flatten().tile_plan)


def test_device_flatten_at_baseline_size(ref_curves):
    """BASELINE configs 2/3: the 1M-trade synthetic book from per-trade arrays.  Counts and arrays equal the host
    flattener's; the device path takes milliseconds where the host path takes seconds."""
    from adrates_b200.synthetic import make_array_book
    curve = _curve(ref_curves["gbp_readme_lzr"])
    book = make_array_book(curve, 1_000_000)
    sess = CurveSession.get(curve, 0)
    assert book.upload(sess.ctx) == "device"
    sess.ctx.sync()
    t0 = time.perf_counter()
    for _ in range(3):
        book.upload(sess.ctx)
    sess.ctx.sync()
    dev_s = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    ref = book.flatten(dedup=True, tiles=True)
    host_s = time.perf_counter() - t0
    print(f"device flatten {dev_s * 1e3:.2f} ms, host flatten {host_s:.2f} s for 1M trades")
    got = sess.ctx.book_read()
    _assert_flat_equal(got, ref)
    _assert_tiles_equal(sess.ctx.book_read_tiles(), ref.tile_plan, curve.path_b_plan().n_nodes)
    assert dev_s < 0.05 and dev_s < host_s / 20
    res, rows = book.compute(ALL)
    res_h, rows_h = book.compute(ALL, device_flatten=False)
    import torch
    assert torch.equal(rows["gamma"][::997], rows_h["gamma"][::997]) and torch.equal(rows["pv"], rows_h["pv"])
    assert abs(res.value.amount - res_h.value.amount) <= 1e-12 * max(abs(res_h.value.amount), 1.0)


def test_narrow_inputs_build_the_same_book_on_the_device():
    """int32 day serials + int8 sides (CAV_BOOK_DATES_I32 | CAV_BOOK_SIGN_I8) against int64 / float64 inputs: the device
    flattener must build bit-identical flat arrays; also with explicit int32 termination dates."""
    import numpy as np
    from adrates_b200 import _native
    from adrates_b200.batch import OISBook
    from adrates_b200.market_data import readme_model
    from adrates_b200.synthetic import make_array_book
    curve = readme_model().curves.GBP_OIS_SONIA
    wide = make_array_book(curve, 20_000, seed=3)
    conv = dict(fixed_freq_type=wide.fixed_freq_type, fixed_dc_type=wide.fixed_dc_type, float_freq_type=wide.float_freq_type,
                float_dc_type=wide.float_dc_type, bd_type=wide.bd_type)
    narrow = OISBook.from_arrays(curve, wide.effective.astype(np.int32), tenor_years=wide._tenor, fixed_sign=wide.fixed_sign.astype(np.int8),
                                 fixed_coupon=wide.coupon, notional=wide.notional, **conv)
    narrow_t = OISBook.from_arrays(curve, wide.effective.astype(np.int32), termination=wide.termination.astype(np.int32),
                                   fixed_sign=wide.fixed_sign.astype(np.int8), fixed_coupon=wide.coupon, notional=wide.notional, **conv)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    flats = []
    for b in (wide, narrow, narrow_t):
        assert b.upload(ctx) == "device"
        flats.append(ctx.book_read())
    for k in ("unit_offsets", "amt", "weight", "node", "comp_weight", "group_offsets", "group_units", "out_index", "unit_weight"):
        for f in flats[1:]:
            assert np.array_equal(getattr(flats[0], k), getattr(f, k)), k


@pytest.mark.parametrize("cal", ["UNITED_KINGDOM", "UNITED_STATES", "TARGET"])
def test_device_flatten_on_a_holiday_calendar(ref_curves, cal):
    """Books rolled on a holiday calendar are flattened on the device too (cav_book_set_holidays: the schedule kernels walk the
    non-business-day bitmap): flat arrays and tile plan equal the host flattener's bit for bit, which the reference's own
    holiday calendars and schedules pin (tests/test_calendars_cpu.py); the per-trade results follow."""
    curve = _curve(ref_curves["gbp_readme_lzr"])
    rng = np.random.default_rng(43)
    conv = dict(CONVS["semi_vs_quarterly"], bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
    arrays = _random_book(curve, 3000, rng, True)
    book = B.OISBook.from_arrays(curve, **arrays, cal_type=CalendarTypes[cal], **conv)
    sess = CurveSession.get(curve, 0)
    assert book.upload(sess.ctx, tiles=True) == "device"
    ref = book.flatten(dedup=True, tiles=True)
    _assert_flat_equal(sess.ctx.book_read(), ref)
    _assert_tiles_equal(sess.ctx.book_read_tiles(), ref.tile_plan, curve.path_b_plan().n_nodes)
    weekend = B.OISBook.from_arrays(curve, **arrays, **conv).flatten(dedup=True, tiles=False)
    assert not np.array_equal(weekend.amt, ref.amt)                    # the holidays moved coupon dates
    tot_d, rows_d = book.compute(ALL)
    tot_h, rows_h = book.compute(ALL, device_flatten=False)
    for k in ("pv", "delta", "gamma"):
        assert np.array_equal(rows_d[k].cpu().numpy(), rows_h[k].cpu().numpy()), k
    # the WEEKEND book afterwards is flattened without the bitmap again
    plain = B.OISBook.from_arrays(curve, **arrays, **conv)
    assert plain.upload(sess.ctx, tiles=False) == "device"
    _assert_flat_equal(sess.ctx.book_read(), weekend)


def test_holiday_calendar_needs_its_bitmap_and_its_range(ref_curves):
    curve = _curve(ref_curves["gbp_readme_lzr"])
    sess = CurveSession.get(curve, 0)
    vd = curve._value_dt._n
    conv = _native.BookConv(vd, 12, 12, 7, 7, CalendarTypes.TARGET.value, 3, 2, 0, 0)
    sess.ctx.book_set_holidays(None)
    args = dict(effective=np.array([vd], dtype=np.int64), tenor=np.array([5], dtype=np.int32), fixed_sign=np.ones(1),
                coupon=np.full(1, 0.03), notional=np.full(1, 1e6))
    with pytest.raises(LibError) as ex:
        sess.ctx.book_from_arrays(conv, **args)
    assert ex.value.code == _native.E_UNSUPPORTED
    from adrates_b200 import holidays as H
    sess.ctx.book_set_holidays(H.table(CalendarTypes.TARGET))
    sess.ctx.book_from_arrays(conv, **args)
    far = dict(args, effective=np.array([Date(1, 6, 2199)._n], dtype=np.int64))
    with pytest.raises(LibError) as ex:
        sess.ctx.book_from_arrays(conv, **far)
    assert "holiday bitmap" in str(ex.value)
