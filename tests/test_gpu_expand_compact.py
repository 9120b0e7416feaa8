"""Gamma expansion from compact unit rows (k_expand_c: the tile kernels leave every unit's gamma as the packed triangle over
its tile's active pillars, the expansion gathers it from L2) against the full-row path (k_expand over 8 KB unit rows): the
per-trade gamma / delta / PV rows and the portfolio totals must be bit-identical - the same staged values reach the same
multiply-adds in the same order.  Switch: CAV_EXPAND_COMPACT (read on every valuation)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _value(ctx, n, mask=7):
    pv = torch.zeros(n, dtype=torch.float64, device="cuda")
    dl = torch.zeros(n, 32, dtype=torch.float64, device="cuda")
    gm = torch.zeros(n, 32, 32, dtype=torch.float64, device="cuda")
    agg = ctx.portfolio_value_host(mask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr()).copy()
    ctx.sync()
    return pv.cpu().numpy(), dl.cpu().numpy(), gm.cpu().numpy(), agg


@pytest.mark.parametrize("n,source", [(30_000, "device"), (5_000, "host"), (777, "device")])
def test_compact_expansion_is_bit_identical_to_full_rows(n, source):
    from adrates_b200 import _native
    from adrates_b200.market_data import readme_model
    from adrates_b200.synthetic import flatten_book, make_array_book, make_book

    curve = readme_model().curves.GBP_OIS_SONIA
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    if source == "device":
        make_array_book(curve, n, seed=5).upload(ctx)            # device flattener + device tile plan
    else:
        ctx.portfolio_upload(flatten_book(make_book(curve, n, seed=5), dedup=True))      # host flattener, pipelined chunk upload
    old = os.environ.get("CAV_EXPAND_COMPACT")
    try:
        os.environ["CAV_EXPAND_COMPACT"] = "0"
        full = _value(ctx, n)
        os.environ["CAV_EXPAND_COMPACT"] = "1"
        comp = _value(ctx, n)
        comp_gamma_only = _value(ctx, n, mask=4)
    finally:
        if old is None:
            os.environ.pop("CAV_EXPAND_COMPACT", None)
        else:
            os.environ["CAV_EXPAND_COMPACT"] = old
    assert np.abs(full[2]).max() > 0
    for a, b in zip(full, comp):
        assert a.tobytes() == b.tobytes()
    assert comp_gamma_only[2].tobytes() == full[2].tobytes()
    # symmetric rows, and zero outside the tile's active pillars exactly where the full path writes zeros
    g = comp[2]
    assert np.array_equal(g, np.swapaxes(g, 1, 2))
