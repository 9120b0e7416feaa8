"""N > 1 path on CPU: two gloo ranks each own a shard of the book, evaluate it (numpy restatement of the
kernels, tests/flat_eval.py) and all-reduce the portfolio totals; the result equals the unsharded total."""
import json
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adrates_b200.parallel import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_balanced_and_covering():
    rng = np.random.default_rng(0)
    cost = rng.integers(1, 52, 10000)
    for world in (1, 2, 4, 8):
        b = shard_bounds(cost, world)
        assert b[0][0] == 0 and b[-1][1] == 10000 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sums = [cost[lo:hi].sum() for lo, hi in b]
        assert max(sums) - min(sums) <= 2 * 51
    assert shard_bounds([], 2) == [(0, 0), (0, 0)]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cavour_oracle as orc
    from adrates_b200.curves import OISCurve
    from adrates_b200.global_types import InterpTypes
    from adrates_b200.synthetic import make_book, flatten_book, Book
    from adrates_b200.parallel import all_reduce_totals
    from tests.flat_eval import eval_flat
    from tests.util_trades import make_calibration_swaps
    cv = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_curves.json")))["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    book = make_book(curve, 120, seed=5)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    cost = np.array([len(book.schedules[s]._fixed_leg._payment_dts) for s in book.sched])
    lo, hi = shard_bounds(cost, world)[rank]
    mine = Book(curve, book.schedules, book.sched[lo:hi], book.coupon[lo:hi], book.notional[lo:hi],
                book.fixed_sign[lo:hi], book.spread[lo:hi])
    pv, dl, gm = eval_flat(flatten_book(mine, dedup=True), d, J, C)
    totals = torch.from_numpy(np.concatenate([[pv.sum()], np.pad(dl.sum(0), (0, 0)), gm.sum(0).reshape(-1)]))
    all_reduce_totals(totals)
    if rank == 0:
        pv_a, dl_a, gm_a = eval_flat(flatten_book(book, dedup=True), d, J, C)
        ref = np.concatenate([[pv_a.sum()], dl_a.sum(0), gm_a.sum(0).reshape(-1)])
        np.save(out, np.stack([totals.numpy(), ref]))
    dist.destroy_process_group()


def test_two_rank_gloo_totals_match_single_process(tmp_path):
    out = str(tmp_path / "tot.npy")
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, ref = np.load(out)
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 1e-12 * scale


def test_array_book_shards_cover_the_book_and_add_up():
    """OISBook.shard: contiguous, covering, balanced by coupon count; the shards' totals (numpy evaluation of their
    flat layouts) add up to the totals of the whole book."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import cavour_oracle as orc
    from adrates_b200.curves import OISCurve
    from adrates_b200.global_types import InterpTypes
    from adrates_b200.synthetic import make_array_book
    from tests.flat_eval import eval_flat
    from tests.util_trades import make_calibration_swaps
    cv = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_curves.json")))["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    book = make_array_book(curve, 300, seed=3)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)

    def totals(b):
        pv, dl, gm = eval_flat(b.flatten(tiles=False), d, J, C)
        return np.concatenate([[pv.sum()], dl.sum(0), gm.sum(0).reshape(-1)])
    ref = totals(book)
    for world in (2, 3, 8):
        shards = [book.shard(r, world) for r in range(world)]
        assert sum(s.n_trades for s in shards) == 300
        assert np.array_equal(np.concatenate([s.notional for s in shards]), book.notional)
        costs = [float(np.sum(s.termination - s.effective)) for s in shards]
        assert max(costs) - min(costs) <= 2 * 51 * 366
        tot = sum(totals(s) for s in shards)
        assert np.max(np.abs(tot - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_schedule_shards_partition_the_book_and_keep_schedules_whole():
    """OISBook.shard_by_schedule: the index sets of the ranks partition the book, no (effective, termination) schedule is
    split over two ranks, coupon counts are balanced, and the trades of a shard keep their original order."""
    from adrates_b200.batch import OISBook, add_weekdays
    from adrates_b200.market_data import readme_model
    curve = readme_model().curves.GBP_OIS_SONIA
    rng = np.random.Generator(np.random.PCG64(3))
    n = 4000
    eff = add_weekdays(np.full(n, curve._value_dt._n), rng.integers(0, 40, n))
    book = OISBook.from_arrays(curve, eff, tenor_years=rng.integers(1, 31, n).astype(np.int32), fixed_sign=np.where(rng.random(n) < 0.5, 1.0, -1.0),
                               fixed_coupon=rng.uniform(0.01, 0.05, n), notional=rng.uniform(1e5, 1e7, n))
    for world in (1, 2, 4, 8):
        parts = [book.shard_by_schedule(r, world) for r in range(world)]
        idx = np.concatenate([p[1] for p in parts])
        assert np.array_equal(np.sort(idx), np.arange(n))
        owner = {}
        for r, (b, ix) in enumerate(parts):
            assert np.all(np.diff(ix) > 0) and np.array_equal(b.notional, book.notional[ix])
            for k in set(zip(b.effective.tolist(), b.termination.tolist())):
                assert owner.setdefault(k, r) == r
        if world > 1:
            costs = [float(np.sum(b.termination - b.effective)) for b, _ in parts]
            assert max(costs) <= 1.25 * (sum(costs) / world)
