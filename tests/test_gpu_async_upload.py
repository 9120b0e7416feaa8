"""Pipelined upload (cav_set_async_upload): per-trade arrays stream in on a side copy stream in chunks while the
units kernel runs.  Results must be bit-identical to the synchronous upload, also when books of the same size
replace each other back to back (a stale chunk or a missed event wait would show up as rows of the previous
book)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from adrates_b200 import _native  # noqa: E402
from adrates_b200.synthetic import flatten_book, make_book  # noqa: E402
from tests.util_trades import build_model  # noqa: E402

MASK = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA


def _pinned(flat):
    fp = copy.copy(flat)
    keep = []
    for k in ("unit_offsets", "amt", "weight", "node", "comp_weight", "group_offsets", "group_units", "out_index",
              "unit_weight"):
        a = getattr(flat, k)
        if a is not None:
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(t)
            setattr(fp, k, t.numpy())
    fp._keep = keep
    return fp


def _value(ctx, flat, n, mask=MASK):
    pv = torch.zeros(n, dtype=torch.float64, device="cuda")
    dl = torch.zeros(n, 32, dtype=torch.float64, device="cuda")
    gm = torch.zeros(n, 32, 32, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.portfolio_upload(flat)
    agg = ctx.portfolio_value_host(mask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr()).copy()
    ctx.sync()
    return pv, dl, gm, agg


def test_async_upload_is_bit_identical_and_race_free(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    curve = build_model(cv).curves.GBP_OIS_SONIA
    n = 120_000
    books = [make_book(curve, n, seed=s) for s in (1, 2, 3)]
    flats = [_pinned(flatten_book(b, dedup=True)) for b in books]
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    ref = [_value(ctx, f, n) for f in flats]
    ctx.set_async_upload(True)
    for rep in range(3):
        for f, r in zip(flats, ref):
            got = _value(ctx, f, n)
            for a, b in zip(got[:3], r[:3]):
                assert torch.equal(a, b)
            assert np.array_equal(got[3], r[3])
    # PV + delta only (row-table path waits for every chunk), then totals only (no per-trade rows)
    pv, dl, _, agg = _value(ctx, flats[0], n, _native.REQ_VALUE | _native.REQ_DELTA)
    ctx.set_async_upload(False)
    pv_s, dl_s, _, agg_s = _value(ctx, flats[0], n, _native.REQ_VALUE | _native.REQ_DELTA)
    ctx.set_async_upload(True)
    assert torch.equal(pv, pv_s) and torch.equal(dl, dl_s) and np.array_equal(agg, agg_s)
    ctx.portfolio_upload(flats[1])
    agg = ctx.portfolio_value_host(MASK)
    assert np.array_equal(agg, ref[1][3])
    # back-to-back uploads without a valuation in between, then the last one is valued
    ctx.portfolio_upload(flats[0])
    ctx.portfolio_upload(flats[2])
    got = _value(ctx, flats[2], n)
    assert torch.equal(got[2], ref[2][2])
    # a small book (fewer than 16 groups) takes the synchronous path
    small = flatten_book(make_book(curve, 40, seed=9), dedup=True)
    a = _value(ctx, small, 40)
    ctx.set_async_upload(False)
    b = _value(ctx, small, 40)
    assert torch.equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    ctx.close()


def test_bad_per_trade_indices_are_rejected_before_any_kernel_reads_them(ref_curves):
    """Synchronous upload: the upload call itself fails.  Pipelined upload: the per-trade arrays are checked on the
    host after the units kernel has been launched, so the error comes from the valuation (or cav_sync) and the
    portfolio is discarded - no expansion kernel ever sees the bad index."""
    from adrates_b200.error import LibError
    cv = ref_curves["gbp_readme_lzr"]
    curve = build_model(cv).curves.GBP_OIS_SONIA
    n = 60_000
    good = _pinned(flatten_book(make_book(curve, n, seed=4), dedup=True))
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    ref = _value(ctx, good, n)
    for field, value, msg in (("out_index", n, "out_index"), ("out_index", -1, "out_index"),
                              ("group_units", good.n_units, "unit id")):
        bad = copy.copy(good)
        arr = getattr(good, field).copy()
        arr[len(arr) // 2] = value
        setattr(bad, field, arr)
        ctx.set_async_upload(False)
        with pytest.raises(LibError, match=msg):
            ctx.portfolio_upload(bad)
        with pytest.raises(LibError):
            ctx.portfolio_value_host(MASK)
        ctx.set_async_upload(True)
        bad = _pinned(bad)
        ctx.portfolio_upload(bad)                       # accepted: only the unit arrays are checked here
        with pytest.raises(LibError, match=msg):
            ctx.portfolio_value_host(MASK)
        with pytest.raises(LibError):
            ctx.portfolio_value_host(MASK)              # ... and the portfolio is gone
        ctx.portfolio_upload(bad)
        with pytest.raises(LibError, match=msg):
            ctx.sync()                                  # cav_sync settles the check too
        got = _value(ctx, good, n)                      # the context is still usable
        assert torch.equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
    ctx.close()
