"""The rules of the device-side book flattener (adrates_b200/csrc/cav_book_core.h, compiled for the host by
tests/native_book.py) against (1) the reference's own schedule / day-count rows, (2) the numpy restatement batch.py on
random inputs, (3) curves.plan_queries, (4) batch.OISBook.flatten (unit arrays bit for bit) and (5) tiles.plan_tiles.
The CUDA kernels of cav_book.cu call the same inline functions; their parallel plumbing is covered by the -m gpu tests."""
import numpy as np
import pytest

from adrates_b200 import batch as B
from adrates_b200.curves import OISCurve, plan_queries
from adrates_b200.dates import (BusDayAdjustTypes, CalendarTypes, Date, DateGenRuleTypes, DayCountTypes, FrequencyTypes,
                                annual_frequency)
from adrates_b200.error import LibError
from adrates_b200.global_types import InterpTypes
from adrates_b200.tiles import node_support_masks, plan_tiles
from tests import native_book as nb
from tests.util_trades import make_calibration_swaps

DCS = ["ACT_365F", "ACT_360", "THIRTY_E_360", "THIRTY_360_BOND", "THIRTY_E_360_ISDA", "ACT_ACT_ISDA", "THIRTY_E_PLUS_360",
       "SIMPLE", "ZERO"]


def _ser(dmy_list):
    return np.array([Date(*x)._n for x in dmy_list], dtype=np.int64)


def _random_dates(rng, n):
    lo, hi = Date(1, 1, 1990)._n, Date(31, 12, 2090)._n
    base = rng.integers(lo, hi, n)
    y = rng.integers(1992, 2090, n // 4)
    eom = B.ordinal(B.days_in_month(np.full_like(y, 2), y), np.full_like(y, 2), y)
    return np.concatenate([base, eom, eom - 1, B.add_months(eom, rng.integers(0, 12, eom.shape[0]), eom=True)])


def test_date_rules_match_reference_rows(ref_schedules):
    rows = ref_schedules["daycounts"]
    n1, n2 = _ser([r["d1"] for r in rows]), _ser([r["d2"] for r in rows])
    d, m, y = nb.ymd(n1)
    assert [[int(a), int(b), int(c)] for a, b, c in zip(d, m, y)] == [r["d1"] for r in rows]
    for dc in DCS[:-1]:
        assert np.array_equal(nb.year_frac(n1, n2, DayCountTypes[dc].value), [r[dc] for r in rows]), dc     # bit-exact
    back = lambda n: [[int(a), int(b), int(c)] for a, b, c in zip(*B.ymd(n))]  # noqa: E731
    assert back(nb.add_months(n1, -7)) == [r["add_months_m7"] for r in rows]
    assert back(nb.adjust(n1, BusDayAdjustTypes.MODIFIED_FOLLOWING.value)) == [r["adj_mf"] for r in rows]
    assert back(nb.adjust(n1, BusDayAdjustTypes.MODIFIED_PRECEDING.value)) == [r["adj_mp"] for r in rows]


def test_date_rules_match_batch_on_random_dates():
    rng = np.random.default_rng(5)
    n = _random_dates(rng, 800)
    for a, b in zip(nb.ymd(n), B.ymd(n)):
        assert np.array_equal(a, b)
    for bd in BusDayAdjustTypes:
        assert np.array_equal(nb.adjust(n, bd.value), B.adjust(n, bd)), bd
        assert np.array_equal(nb.adjust(n, bd.value, CalendarTypes.NONE.value), n)
    for mm in (-25, -12, -1, 0, 1, 6, 12, 13, 600):
        assert np.array_equal(nb.add_months(n, mm), B.add_months(n, mm)), mm
        assert np.array_equal(nb.add_months(n, mm, eom=True), B.add_months(n, mm, eom=True)), mm
    for c in (0, 1, 2, 4, 5, 30, 50):
        assert np.array_equal(nb.add_tenor(n, c, True), B.add_tenor(n, c, "Y")), c
        assert np.array_equal(nb.add_tenor(n, c, False), B.add_tenor(n, c, "M")), c
    n2 = n + rng.integers(-300, 20000, n.shape[0])
    for dc in DCS:
        assert np.array_equal(nb.year_frac(n, n2, DayCountTypes[dc].value), B.year_frac(n, n2, DayCountTypes[dc])), dc


def test_schedules_match_reference_rows(ref_schedules):
    """All 1750 schedules the unmodified reference generated (dates or LibError)."""
    n = 0
    for r in ref_schedules["schedules"]:
        step = int(12 / annual_frequency(FrequencyTypes[r["freq"]]))
        got = nb.schedule(Date(*r["eff"])._n, Date(*r["term"])._n, step, CalendarTypes.WEEKEND.value,
                          BusDayAdjustTypes[r["bd"]].value, DateGenRuleTypes[r["dg"]].value)
        if r["dates"] == "ERR:LibError":
            assert isinstance(got, int) and got < 0, r
        else:
            assert not isinstance(got, int), (r, got)
            assert got.tolist() == [Date(*x)._n for x in r["dates"]], r
        n += 1
    assert n == len(ref_schedules["schedules"]) >= 1750


@pytest.mark.parametrize("dg", list(DateGenRuleTypes))
@pytest.mark.parametrize("eom", [False, True])
def test_schedules_match_batch_on_random_dates(dg, eom):
    rng = np.random.default_rng(21 + eom)
    eff = _random_dates(rng, 120)
    for freq in (FrequencyTypes.ANNUAL, FrequencyTypes.SEMI_ANNUAL, FrequencyTypes.TRI_ANNUAL, FrequencyTypes.QUARTERLY,
                 FrequencyTypes.MONTHLY):
        step = int(12 / annual_frequency(freq))
        for bd in BusDayAdjustTypes:
            span = rng.integers(1, 12000, eff.shape[0])
            span[::7] = rng.integers(1, 40, span[::7].shape[0])
            for e, t in zip(eff, eff + span):
                got = nb.schedule(e, t, step, CalendarTypes.WEEKEND.value, bd.value, dg.value, eom)
                try:
                    ref = B.roll_schedules([e], [t], freq, CalendarTypes.WEEKEND, bd, dg, True, eom).of(0).tolist()
                except LibError:
                    assert isinstance(got, int) and got < 0
                else:
                    assert not isinstance(got, int) and got.tolist() == ref, (freq, bd, dg, eom, e, t)


def _curve(cv):
    vd, swaps = make_calibration_swaps(cv)
    return OISCurve(vd, swaps, InterpTypes[cv["interp"]])


@pytest.mark.parametrize("name", ["gbp_readme_lzr", "gbp_semi_lzr"])
@pytest.mark.parametrize("interp", [InterpTypes.LINEAR_ZERO_RATES, InterpTypes.FLAT_FWD_RATES])
def test_bracket_planner_matches_plan_queries(ref_curves, name, interp):
    x = _curve(ref_curves[name]).path_b_plan().node_time
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.uniform(-1.0, 62.0, 4000), x, x + 5e-11, x - 5e-11, x + 2e-10, x - 2e-10, x + 1e-12,
                        [0.0, 1e-15, 1e-13, 49.99999, 50.0, 50.01, 70.0, -0.5]])
    a, b, wa, wb = plan_queries(t, x, interp)
    wref = np.stack([wa, wb], 1)
    nref = np.stack([a, b], 1).astype(np.int32)
    nref[wref == 0.0] = 0                       # as OISBook._plan stores them
    ga, gb, gwa, gwb = nb.plan_queries(t, x, interp == InterpTypes.LINEAR_ZERO_RATES)
    assert np.array_equal(np.stack([ga, gb], 1), nref)
    assert np.array_equal(gwa, wa) and np.array_equal(gwb, wb)          # bit-exact weights


CONVS = {
    "annual_act365": dict(fixed_freq_type=FrequencyTypes.ANNUAL, fixed_dc_type=DayCountTypes.ACT_365F,
                          float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=DayCountTypes.ACT_365F,
                          bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING),
    "semi_vs_quarterly": dict(fixed_freq_type=FrequencyTypes.SEMI_ANNUAL, fixed_dc_type=DayCountTypes.THIRTY_E_360,
                              float_freq_type=FrequencyTypes.QUARTERLY, float_dc_type=DayCountTypes.ACT_360,
                              bd_type=BusDayAdjustTypes.FOLLOWING),
    "forward_eom_isda": dict(fixed_freq_type=FrequencyTypes.SEMI_ANNUAL, fixed_dc_type=DayCountTypes.ACT_ACT_ISDA,
                             float_freq_type=FrequencyTypes.SEMI_ANNUAL, float_dc_type=DayCountTypes.ACT_365F,
                             bd_type=BusDayAdjustTypes.PRECEDING, dg_type=DateGenRuleTypes.FORWARD),
}


def _random_book(curve, n, rng, spread=False):
    vd = curve._value_dt._n
    off = rng.integers(-400, 300, n)
    off[: n // 3] = 0
    eff = np.full(n, vd) + off
    months = rng.integers(1, 480, n)
    months[: n // 2] = 12 * rng.integers(1, 40, n // 2)
    months[::5] = months[1::5][: months[::5].shape[0]]
    eff[::5] = eff[1::5][: eff[::5].shape[0]]
    return dict(effective=eff, tenor_months=months, fixed_sign=np.where(rng.random(n) < 0.5, 1.0, -1.0),
                fixed_coupon=rng.uniform(0.01, 0.07, n), notional=np.exp(rng.uniform(11, 18, n)),
                float_spread=np.where(rng.random(n) < 0.5, rng.normal(0, 0.002, n), 0.0) if spread else 0.0)


def book_conv9(book):
    return nb.conv9(book.curve._value_dt._n, int(12 / annual_frequency(book.fixed_freq_type)),
                    int(12 / annual_frequency(book.float_freq_type)), book.fixed_dc_type.value, book.float_dc_type.value,
                    book.cal_type.value, book.bd_type.value, book.dg_type.value, 0)


@pytest.mark.parametrize("conv", list(CONVS))
@pytest.mark.parametrize("spread", [False, True])
def test_class_units_equal_batch_flatten_bit_for_bit(ref_curves, conv, spread):
    """Unit arrays (offsets, amounts, bracket weights and nodes) of the schedule classes of a random book: the shared core
    against batch.OISBook.flatten(dedup=True), which is pinned to the object layer and through it to the reference."""
    curve = _curve(ref_curves["gbp_readme_lzr"])
    rng = np.random.default_rng(17)
    book = B.OISBook.from_arrays(curve, **_random_book(curve, 400, rng, spread), **CONVS[conv])
    flat = book.flatten(dedup=True, tiles=False)
    eff, term, cls_of = book.schedule_classes()
    with_spread = np.bincount(cls_of[book.spread != 0.0], minlength=eff.shape[0]) > 0
    x = curve.path_b_plan().node_time
    err, off, amt, weight, node, has3, uid3, _ = nb.flatten_classes(
        book_conv9(book), eff, term, with_spread if spread else None, x, curve._interp_type == InterpTypes.LINEAR_ZERO_RATES)
    assert err == 0
    assert np.array_equal(off, flat.unit_offsets)
    assert np.array_equal(amt, flat.amt)
    assert np.array_equal(weight, flat.weight)
    assert np.array_equal(node, flat.node)
    # trades of a class point at (annuity, floating[, spread annuity]) units in class order
    S = eff.shape[0]
    K = flat.n_comp
    ids = np.stack([np.where(has3[k * S:(k + 1) * S][cls_of], uid3[k * S:(k + 1) * S][cls_of], 0) for k in range(K)], 1)
    order = flat.out_index
    got_ids = np.repeat(flat.group_units.reshape(-1, K), np.diff(flat.group_offsets), axis=0)
    assert np.array_equal(got_ids, ids[order])


def test_class_walk_reports_errors():
    cv = nb.conv9(Date(30, 4, 2024)._n, 12, 12, 7, 7, 2, 3, 2)
    x = np.array([0.0, 1.0, 2.0])
    e = Date(30, 4, 2024)._n
    err, *_ = nb.flatten_classes(cv, [e], [e], None, x, True)
    assert err & 2                                  # effective date not before termination
    err, off, amt, *_ = nb.flatten_classes(cv, [e - 4000], [e - 3000], None, x, True)
    assert err == 0 and off[-1] == 0                # matured: no live terms, no units


@pytest.mark.parametrize("conv", ["annual_act365", "semi_vs_quarterly"])
def test_tile_plan_equals_host_planner(ref_curves, conv):
    """K rows, tiles, masks, permutation and class order of the shared core against tiles.plan_tiles (pair rows are numbered
    in node order here and in first-seen order there: compared through the node pair they stand for)."""
    curve = _curve(ref_curves["gbp_readme_lzr"])
    rng = np.random.default_rng(23)
    book = B.OISBook.from_arrays(curve, **_random_book(curve, 1500, rng, spread=True), **CONVS[conv])
    flat = book.flatten(dedup=True, tiles=False)
    plan = curve.path_b_plan()
    G = plan.n_nodes
    support = node_support_masks(plan.node_swap, plan.node_prev, plan.node_acc)
    ref = plan_tiles(flat, G, support=support)
    got = nb.plan_tiles(flat.unit_offsets, flat.weight, flat.node, G, support)
    assert got["n_tiles"] == ref.n_tiles
    assert np.array_equal(got["perm"], ref.perm)
    assert np.array_equal(got["tile_units"], ref.tile_units)
    assert np.array_equal(got["tile_kcount"], ref.tile_kcount)
    assert np.array_equal(got["tile_kstart"], ref.tile_kstart)
    assert np.array_equal(got["tile_npos"], ref.tile_npos)
    assert np.array_equal(got["tile_mask"], ref.tile_mask)

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
    assert sorted(map(tuple, got["pairs"].reshape(-1, 2))) == sorted(map(tuple, ref.pairs.reshape(-1, 2)))
