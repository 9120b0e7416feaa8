"""BASELINE.json full size (1M synthetic SONIA OIS, configs[2]) through size-independent properties, plus the
limits of the C ABI.  The oracle finishes only samples at this size, so the checks are: a sample of trades against
the C oracle, checksum of checksums (portfolio totals = sum of the per-trade rows), symmetry of every gamma matrix,
PAY + RECEIVE = 0 (linearity in the trade weights), and dedup layout = private layout on a slice."""
import numpy as np
import pytest
import torch

from oracle import cavour_oracle as orc, c_oracle
from adrates_b200 import _native
from adrates_b200.error import LibError
from adrates_b200.synthetic import make_book, flatten_book, reference_leg_tables
from tests.test_gpu_parity import _curve, METHOD

pytestmark = pytest.mark.gpu
MASK = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA


def test_one_million_trades_properties(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    n = 1_000_000
    book = make_book(curve, n, seed=20240430)
    flat = flatten_book(book, dedup=True)
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    ctx.portfolio_upload(flat)
    pv = torch.empty(n, dtype=torch.float64, device="cuda")
    dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
    gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    agg = ctx.portfolio_value_host(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr()).copy()
    ctx.sync()
    # (1) sample against the C oracle
    k = 512
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    tr = dict(sched=book.sched[:k], coupon=book.coupon[:k], notional=book.notional[:k], spread=book.spread[:k],
              fixed_sign=book.fixed_sign[:k])
    pv_c, dl_c, gm_c = c_oracle.ois_batch((plan["times"], d, J, C), METHOD[cv["interp"]], reference_leg_tables(book), tr,
                                          dense=False)
    N = book.notional[:k]
    assert np.max(np.abs(pv[:k].cpu().numpy() - pv_c) / np.maximum(np.abs(pv_c), N)) < 1e-10
    assert np.max(np.abs(dl[:k].cpu().numpy() - dl_c) / np.maximum(np.abs(dl_c), (N * 1e-4)[:, None])) < 1e-10
    assert np.max(np.abs(gm[:k].cpu().numpy() - gm_c) / np.maximum(np.abs(gm_c), (N * 1e-8)[:, None, None])) < 1e-10
    # (2) checksum of checksums: totals = sum of rows (different summation orders: 1e-11 of the absolute mass)
    assert abs(agg[0] - float(pv.sum())) <= 1e-11 * float(pv.abs().sum())
    assert np.max(np.abs(agg[1:33] - dl.sum(0).cpu().numpy())) <= 1e-11 * float(dl.abs().sum())
    assert np.max(np.abs(agg[33:] - gm.sum(0).reshape(-1).cpu().numpy())) <= 1e-11 * float(gm.abs().sum())
    # (3) every gamma matrix is symmetric to the last bit (both halves are written from one packed triangle)
    for s in range(0, n, 125_000):
        blk = gm[s:s + 125_000]
        assert bool(torch.equal(blk, blk.transpose(1, 2)))
    # (4) linearity: the same book with every trade's side flipped cancels the first one exactly
    book.fixed_sign = -book.fixed_sign
    ctx.portfolio_upload(flatten_book(book, dedup=True))
    pv2 = torch.empty_like(pv)
    agg2 = ctx.portfolio_value_host(MASK, pv2.data_ptr(), None, None).copy()
    ctx.sync()
    assert bool(torch.equal(pv2, -pv))
    assert np.max(np.abs(agg + agg2)) <= 1e-11 * np.max(np.abs(agg))
    ctx.close()


def test_dedup_and_private_layouts_agree_on_a_slice(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    n = 60_000
    book = make_book(curve, n, seed=20240430)
    outs = []
    for dedup in (True, False):
        ctx = _native.Context(0)
        ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
        ctx.portfolio_upload(flatten_book(book, dedup=dedup))
        pv = torch.empty(n, dtype=torch.float64, device="cuda")
        dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
        gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.portfolio_value(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), None)
        ctx.sync()
        outs.append((pv, dl, gm))
        ctx.close()
    N = torch.from_numpy(book.notional).cuda()
    assert float(((outs[0][0] - outs[1][0]).abs() / N).max()) < 1e-12
    assert float(((outs[0][1] - outs[1][1]).abs() / (N * 1e-4)[:, None]).max()) < 1e-11
    assert float(((outs[0][2] - outs[1][2]).abs() / (N * 1e-8)[:, None, None]).max()) < 1e-10


def test_abi_limits(ref_curves):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    ctx = _native.Context(0)
    plan = curve.path_b_plan()
    with pytest.raises(LibError, match="more than 32 par-rate pillars"):
        ctx.curve_build(curve._interp_type.value, list(curve.swap_rates) + [0.04], plan, order=2)
    with pytest.raises(LibError, match="Invalid interpolation scheme"):
        ctx.curve_build(7, curve.swap_rates, plan, order=2)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, plan, order=0)     # DFs only
    flat = flatten_book(make_book(curve, 50, seed=1), dedup=True, tiles=False)
    ctx.portfolio_upload(flat)
    with pytest.raises(LibError, match="jacobian"):
        ctx.portfolio_value_host(MASK)
    bad = flatten_book(make_book(curve, 50, seed=1), dedup=True, tiles=False)
    bad.node = bad.node.copy()
    bad.node[0] = 10_000
    with pytest.raises(LibError, match="node index out of range"):
        ctx.portfolio_upload(bad)
    with pytest.raises(LibError):
        ctx.portfolio_value_host(_native.REQ_VALUE)       # the rejected upload left no valid portfolio
    ctx.close()
