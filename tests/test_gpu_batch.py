"""Array-based books on the GPU: OISBook.compute (batch.py -> C ABI) against the object-based Portfolio.compute,
the numpy evaluation of the same flat arrays, and the C oracle on the bench book."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import cavour_oracle as orc  # noqa: E402
from adrates_b200 import RequestTypes, batch as B  # noqa: E402
from adrates_b200.dates import Date  # noqa: E402
from adrates_b200.global_types import CurrencyTypes, CurveTypes, SwapTypes  # noqa: E402
from adrates_b200.position import Portfolio  # noqa: E402
from adrates_b200.trades import OIS  # noqa: E402
from tests.flat_eval import eval_flat  # noqa: E402
from tests.test_batch_cpu import CONVS, _random_book  # noqa: E402
from tests.util_trades import build_model  # noqa: E402

ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
TOL = 1e-10


def _scaled(got, ref, scale):
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale)))


@pytest.mark.parametrize("conv", list(CONVS))
def test_book_compute_matches_flat_arrays_and_portfolio(ref_curves, conv):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    rng = np.random.default_rng(23)
    n = 150
    spec = _random_book(curve, n, rng, spread=(conv != "annual_act365"))
    book = B.OISBook.from_arrays(curve, **spec, **CONVS[conv])
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    c = CONVS[conv]
    sp = np.broadcast_to(spec["float_spread"], (n,))
    positions = []
    for i in range(n):
        eff = Date._of(int(spec["effective"][i]))
        positions.append(OIS(eff, eff.add_tenor(f"{int(spec['tenor_months'][i])}M"),
                             SwapTypes.RECEIVE if spec["fixed_sign"][i] > 0 else SwapTypes.PAY,
                             float(spec["fixed_coupon"][i]), c["fixed_freq_type"], c["fixed_dc_type"],
                             CurveTypes.GBP_OIS_SONIA, CurrencyTypes.GBP, float(spec["notional"][i]),
                             c.get("payment_lag", 0), float(sp[i]), c["float_freq_type"], c["float_dc_type"],
                             bd_type=c["bd_type"]).position(model))
    ref_tot = Portfolio(positions).compute(ALL)
    for dedup in ((True,) if conv == "lagged" else (True, False)):
        res, rows = book.compute(ALL, dedup=dedup)
        exp = eval_flat(book.flatten(dedup=dedup, tiles=False), d, J, C)
        N = 1e8
        assert _scaled(rows["pv"].cpu().numpy(), exp[0], N) < TOL
        assert _scaled(rows["delta"].cpu().numpy(), exp[1], N * 1e-4 * 40) < TOL
        assert _scaled(rows["gamma"].cpu().numpy(), exp[2], N * 1e-8 * 1600) < TOL
        assert abs(res.value.amount - ref_tot.value.amount) <= TOL * n * N
        assert _scaled(res.risk.risk_ladder, ref_tot.risk.risk_ladder, n * N * 1e-4) < TOL
        assert _scaled(res.gamma.risk_ladder, ref_tot.gamma.risk_ladder, n * N * 1e-8 * 40) < TOL
        assert res.value.currency == CurrencyTypes.GBP and res.risk.tenors == ref_tot.risk.tenors


def test_array_bench_book_matches_c_oracle_and_object_book(ref_curves):
    """200k trades of the BASELINE book built without any trade object: rows equal those of the object-built book
    (same draws), a sample equals the C oracle."""
    from adrates_b200.position import CurveSession
    from adrates_b200.synthetic import flatten_book, make_array_book, make_book, reference_leg_tables
    from oracle import c_oracle
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    n = 200_000
    arr = make_array_book(curve, n)
    res, rows = arr.compute(ALL)
    obj = make_book(curve, n)
    sess = CurveSession.get(curve, 0)
    sess.ctx.portfolio_upload(flatten_book(obj, dedup=True))
    pv = torch.empty(n, dtype=torch.float64, device="cuda")
    dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
    gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda")
    agg = sess.ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr())
    sess.ctx.sync()
    # same trades, different unit order (tile composition differs): equal up to FP64 reassociation
    assert float(((rows["pv"] - pv).abs() / pv.abs().clamp_min(1e5)).max()) < 1e-12
    assert float(((rows["delta"] - dl).abs() / dl.abs().clamp_min(1e5 * 1e-4)).max()) < 1e-12
    assert float(((rows["gamma"] - gm).abs() / gm.abs().clamp_min(1e5 * 1e-8)).max()) < 1e-12
    assert abs(res.value.amount - agg[0]) <= 1e-12 * float(pv.abs().sum())
    # private layout of the array book: same rows to 1e-10
    res_p, rows_p = arr.compute(ALL, dedup=False)
    N = 1e8
    assert _scaled(rows_p["pv"].cpu().numpy(), pv.cpu().numpy(), N) < TOL
    assert _scaled(rows_p["delta"][:20000].cpu().numpy(), dl[:20000].cpu().numpy(), N * 1e-4 * 50) < TOL
    assert _scaled(rows_p["gamma"][:4000].cpu().numpy(), gm[:4000].cpu().numpy(), N * 1e-8 * 2500) < TOL
    # C oracle on a sample of the same trades (inputs from the OBJECT schedules, i.e. an independent date path)
    m = 300
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    tr = dict(sched=obj.sched[:m], coupon=obj.coupon[:m], notional=obj.notional[:m], spread=obj.spread[:m],
              fixed_sign=obj.fixed_sign[:m])
    o_pv, o_dl, o_gm = c_oracle.ois_batch((plan["times"], d, J, C), 4 if cv["interp"] == "LINEAR_ZERO_RATES" else 1,
                                          reference_leg_tables(obj), tr, dense=False)
    assert _scaled(rows["pv"][:m].cpu().numpy(), o_pv, N) < TOL
    assert _scaled(rows["delta"][:m].cpu().numpy(), o_dl, N * 1e-4 * 50) < TOL
    assert _scaled(rows["gamma"][:m].cpu().numpy(), o_gm, N * 1e-8 * 2500) < TOL


def test_portfolio_compute_array_route_equals_object_route(ref_curves):
    """Portfolio.compute over >= 512 vanilla OIS objects takes the array route (position._value_as_arrays); the
    totals equal those of the object flattener, for a book mixing two convention sets."""
    from adrates_b200 import position as P
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    rng = np.random.default_rng(5)
    derivs = []
    for conv in ("annual_act365", "semi_vs_quarterly"):
        c = CONVS[conv]
        n = 330
        spec = _random_book(curve, n, rng, spread=(conv != "annual_act365"))
        sp = np.broadcast_to(spec["float_spread"], (n,))
        for i in range(n):
            eff = Date._of(int(spec["effective"][i]))
            derivs.append(OIS(eff, eff.add_tenor(f"{int(spec['tenor_months'][i])}M"),
                              SwapTypes.RECEIVE if spec["fixed_sign"][i] > 0 else SwapTypes.PAY,
                              float(spec["fixed_coupon"][i]), c["fixed_freq_type"], c["fixed_dc_type"],
                              CurveTypes.GBP_OIS_SONIA, CurrencyTypes.GBP, float(spec["notional"][i]), 0, float(sp[i]),
                              c["float_freq_type"], c["float_dc_type"], bd_type=c["bd_type"]))
    assert len(derivs) >= P.ARRAY_ROUTE_MIN and all(P._ois_conventions(d) is not None for d in derivs)
    fast = Portfolio([d.position(model) for d in derivs]).compute(ALL)
    slow = P.value_positions(derivs, curve, ALL, dedup=True)          # explicit dedup -> object flattener
    S = len(derivs) * 1e8
    assert abs(fast.value.amount - slow.value.amount) <= TOL * S
    assert _scaled(fast.risk.risk_ladder, slow.risk.risk_ladder, S * 1e-4) < TOL
    assert _scaled(fast.gamma.risk_ladder, slow.gamma.risk_ladder, S * 1e-8 * 40) < TOL
    # anything the array route cannot express exactly falls back to objects
    derivs[7]._float_leg._notional_array = [1.0]
    assert P._ois_conventions(derivs[7]) is None


def _dist_worker(rank, world, port, out):
    import os
    import sys
    import json
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adrates_b200.synthetic import make_array_book
    from tests.util_trades import build_model as bm
    cv = json.load(open(os.path.join(root, "tests", "golden", "ref_curves.json")))["gbp_readme_lzr"]
    curve = bm(cv).curves.GBP_OIS_SONIA
    book = make_array_book(curve, 5000, seed=21)
    res, rows, (lo, hi) = book.compute_distributed(ALL, device=0)       # both ranks share the one test GPU
    # the same book dealt by schedule: same totals, the shard's rows are the rows of its index set
    res_s, rows_s, idx = book.compute_distributed(ALL, device=0, shard="schedule")
    tot_s = np.concatenate([[res_s.value.amount], res_s.risk.risk_ladder, res_s.gamma.risk_ladder.reshape(-1)])
    np.save(f"{out}_sched_{rank}.npy", np.concatenate([tot_s, [float(len(idx)), float(rows_s["pv"].sum())]]))
    np.save(f"{out}_{rank}.npy", np.concatenate([[res.value.amount], res.risk.risk_ladder, res.gamma.risk_ladder.reshape(-1),
                                                 [lo, hi, float(rows["pv"].sum())]]))
    dist.destroy_process_group()


def test_compute_distributed_two_ranks_equal_single_process(ref_curves, tmp_path):
    """OISBook.compute_distributed: two processes (gloo group, one GPU) value their shards and all-reduce the totals;
    every rank returns the totals of the whole book."""
    import torch.multiprocessing as mp
    from adrates_b200.synthetic import make_array_book
    out = str(tmp_path / "dist")
    mp.spawn(_dist_worker, args=(2, 29533, out), nprocs=2, join=True)
    curve = build_model(ref_curves["gbp_readme_lzr"]).curves.GBP_OIS_SONIA
    book = make_array_book(curve, 5000, seed=21)
    res, rows = book.compute(ALL)
    ref = np.concatenate([[res.value.amount], res.risk.risk_ladder, res.gamma.risk_ladder.reshape(-1)])
    a, b = np.load(out + "_0.npy"), np.load(out + "_1.npy")
    assert np.array_equal(a[:-3], b[:-3])                                 # same totals on every rank
    assert np.max(np.abs(a[:-3] - ref) / np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))) < 1e-10
    assert (a[-3], b[-2]) == (0, 5000) and a[-2] == b[-3]                 # shards are contiguous and cover the book
    assert abs(a[-1] + b[-1] - float(rows["pv"].sum())) <= 1e-10 * float(rows["pv"].abs().sum())
    sa, sb = np.load(out + "_sched_0.npy"), np.load(out + "_sched_1.npy")
    assert np.array_equal(sa[:-2], sb[:-2])
    assert np.max(np.abs(sa[:-2] - ref) / np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))) < 1e-10
    assert sa[-2] + sb[-2] == 5000 and min(sa[-2], sb[-2]) > 1500
    assert abs(sa[-1] + sb[-1] - float(rows["pv"].sum())) <= 1e-10 * float(rows["pv"].abs().sum())
