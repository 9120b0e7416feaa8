"""Synthetic SONIA OIS portfolios of BASELINE.json configs 2-4.

Book: tenor ~ U{1..50} years, annual/annual ACT/365F; effective date = value date for 50 %
of trades (cashflows land on grid nodes) and value date + U{1..250} business days for the
rest (exercises the bracket path); fixed coupon ~ N(par(tenor), 50 bp) clipped to
[0.5 %, 9 %]; notional ~ logU[1e5, 1e8]; PAY/RECEIVE 50/50; payment lag 0; spread 0;
numpy Generator(PCG64) with the given seed.

Schedules are built once per distinct (tenor, start offset) with the ordinary OIS class
(<= 50 + 50*250 objects); trades are arrays over those schedules.  `flatten_book` lays the
book out either with shared schedule units (dedup=True, the product default) or with one
private unit per trade (dedup=False, the general irregular-trade form).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .curves import OISCurve
from .dates import BusDayAdjustTypes, DayCountTypes, FrequencyTypes, times_from_dates
from .flatten import FlatPortfolio, assemble, group_trades, ois_components, _merge_single_df_terms, _Unit
from .global_types import CurrencyTypes, CurveTypes, SwapTypes
from .trades import OIS


@dataclass
class Book:
    curve: OISCurve
    schedules: list            # OIS objects with coupon 1, notional 1, PAY fixed
    sched: np.ndarray          # int32 [N] schedule id per trade
    coupon: np.ndarray         # f64 [N]
    notional: np.ndarray       # f64 [N]
    fixed_sign: np.ndarray     # f64 [N] +1 receive fixed, -1 pay fixed
    spread: np.ndarray         # f64 [N]

    @property
    def n_trades(self) -> int:
        return int(self.sched.shape[0])


def make_book(curve: OISCurve, n_trades: int, seed: int = 20240430, max_offset_bd: int = 250,
              index=CurveTypes.GBP_OIS_SONIA, currency=CurrencyTypes.GBP,
              dc=DayCountTypes.ACT_365F) -> Book:
    rng = np.random.Generator(np.random.PCG64(seed))
    tenor = rng.integers(1, 51, n_trades)
    on_grid = rng.random(n_trades) < 0.5
    offset = np.where(on_grid, 0, rng.integers(1, max_offset_bd + 1, n_trades))
    par = np.interp(tenor.astype(np.float64), np.array(curve.swap_times), np.array(curve.swap_rates))
    coupon = np.clip(par + rng.normal(0.0, 0.005, n_trades), 0.005, 0.09)
    notional = np.exp(rng.uniform(np.log(1e5), np.log(1e8), n_trades))
    fixed_sign = np.where(rng.random(n_trades) < 0.5, 1.0, -1.0)
    key = tenor.astype(np.int64) * (max_offset_bd + 1) + offset
    uniq, sched = np.unique(key, return_inverse=True)
    vd = curve._value_dt
    starts = {}
    schedules = []
    for k in uniq:
        T, off = int(k // (max_offset_bd + 1)), int(k % (max_offset_bd + 1))
        if off not in starts:
            starts[off] = vd if off == 0 else vd.add_weekdays(off)
        schedules.append(OIS(effective_dt=starts[off], term_dt_or_tenor=f"{T}Y", fixed_leg_type=SwapTypes.PAY,
                             fixed_coupon=1.0, fixed_freq_type=FrequencyTypes.ANNUAL, fixed_dc_type=dc,
                             floating_index=index, currency=currency, notional=1.0,
                             float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                             bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING))
    return Book(curve, schedules, sched.astype(np.int32), coupon, notional, fixed_sign, np.zeros(n_trades))


def make_array_book(curve: OISCurve, n_trades: int, seed: int = 20240430, max_offset_bd: int = 250,
                    dc=DayCountTypes.ACT_365F):
    """The same book as make_book (identical random draws) as an array-based batch.OISBook: no trade or schedule
    objects at all - effective dates by closed-form weekday stepping, schedules rolled as arrays."""
    from .batch import OISBook, add_weekdays
    rng = np.random.Generator(np.random.PCG64(seed))
    tenor = rng.integers(1, 51, n_trades)
    on_grid = rng.random(n_trades) < 0.5
    offset = np.where(on_grid, 0, rng.integers(1, max_offset_bd + 1, n_trades))
    par = np.interp(tenor.astype(np.float64), np.array(curve.swap_times), np.array(curve.swap_rates))
    coupon = np.clip(par + rng.normal(0.0, 0.005, n_trades), 0.005, 0.09)
    notional = np.exp(rng.uniform(np.log(1e5), np.log(1e8), n_trades))
    fixed_sign = np.where(rng.random(n_trades) < 0.5, 1.0, -1.0)
    eff = add_weekdays(np.full(n_trades, curve._value_dt._n), offset)
    return OISBook.from_arrays(curve, eff, tenor_years=tenor, fixed_sign=fixed_sign, fixed_coupon=coupon,
                               notional=notional, fixed_freq_type=FrequencyTypes.ANNUAL, fixed_dc_type=dc,
                               float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                               bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)


def flatten_book(book: Book, dedup: bool = True, max_group: int = 256, sort_units: bool = True,
                 tiles: bool = True, compact: bool = True) -> FlatPortfolio:
    flat = _flatten_book(book, dedup, max_group, sort_units)
    if not tiles:
        return flat
    plan = book.curve.path_b_plan()
    return flat.with_tiles(plan.n_nodes, plan if compact else None)


def _flatten_book(book: Book, dedup: bool, max_group: int, sort_units: bool) -> FlatPortfolio:
    vd = book.curve._value_dt
    comps = [ois_components(s, vd) for s in book.schedules]   # unit-notional, unit-coupon, PAY fixed
    for c in comps:
        kinds = [k[0][0] for k in c]
        if kinds != ["A", "F"]:
            raise ValueError("synthetic schedules must decompose into an annuity and a floating unit")
    n = book.n_trades
    # trade = fixed_sign*N*c * annuity(schedule) - fixed_sign*N * float(schedule)
    # (schedule objects were built PAY fixed with N = c = 1: comps weights are -1 (annuity) and +1 (float))
    wA = book.fixed_sign * book.notional * book.coupon
    wF = -book.fixed_sign * book.notional
    if dedup:
        # unit order = all annuity units (schedule order), then all floating units: neighbouring units then
        # bracket the same grid nodes, which the tiled units kernel exploits (shared table rows per tile)
        S = len(comps)
        units = [_merge_single_df_terms(c[0][1]) for c in comps] + [_merge_single_df_terms(c[1][1]) for c in comps]
        ids = np.stack([book.sched, S + book.sched], axis=1).astype(np.int32)
        ws = np.stack([wA, wF], axis=1)
        base = assemble(book.curve, units, [[(0, 1.0)]], 1, direct=True)   # plan the unit terms once
        return group_trades(len(units), base.unit_offsets, base.n_pairs, base.amt, base.weight, base.node, ids, ws,
                            max_group)
    # private unit per trade: template = merged terms of (annuity, float) with separate amount vectors
    tmpl_units, a_vecs, f_vecs = [], [], []
    for c in comps:
        ua, uf = c[0][1], c[1][1]
        times = sorted({q[0][0] for q in ua.times} | {q[0][0] for q in uf.times})
        pos = {t: i for i, t in enumerate(times)}
        a = np.zeros(len(times))
        f = np.zeros(len(times))
        for q, amt in zip(ua.times, ua.amts):
            a[pos[q[0][0]]] += amt
        for q, amt in zip(uf.times, uf.amts):
            f[pos[q[0][0]]] += amt
        tmpl_units.append(_Unit([((t, 1.0),) for t in times], [1.0] * len(times)))
        a_vecs.append(a)
        f_vecs.append(f)
    tm = assemble(book.curve, tmpl_units, [[(0, 1.0)]], 1, direct=True)
    a_all, f_all = np.concatenate(a_vecs), np.concatenate(f_vecs)
    # units in schedule order (neighbours bracket the same nodes); rows go back to trade order via out_index
    order = np.argsort(book.sched, kind="stable") if sort_units else np.arange(n)
    sched_sorted = book.sched[order]
    wA, wF = wA[order], wF[order]
    cnt = np.diff(tm.unit_offsets)[sched_sorted]
    unit_offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=unit_offsets[1:])
    n_terms = int(unit_offsets[-1])
    src = np.repeat(tm.unit_offsets[:-1][sched_sorted] - unit_offsets[:-1], cnt) + np.arange(n_terms, dtype=np.int64)
    amt = np.repeat(wA, cnt) * a_all[src] + np.repeat(wF, cnt) * f_all[src]
    weight = tm.weight.reshape(-1, 2)[src].reshape(-1)
    node = tm.node.reshape(-1, 2)[src].reshape(-1)
    return FlatPortfolio(n, n_terms, unit_offsets, 2, amt, np.ascontiguousarray(weight), np.ascontiguousarray(node),
                         n, 1, np.ones(n), n, np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32),
                         order.astype(np.int64) if sort_units else None, np.ones(n))


def reference_leg_tables(book: Book):
    """Per-schedule leg arrays in the shape the reference engine extracts them
    (engine.py:2519-2527, 2858-2877) - input of the CPU oracle baseline."""
    vd = book.curve._value_dt
    fo, lo = [0], [0]
    f_pay_t, f_alpha, l_start_t, l_end_t, l_pay_t, l_alpha = [], [], [], [], [], []
    for s in book.schedules:
        fl, ft = s._fixed_leg, s._float_leg
        f_pay_t += [times_from_dates(d, vd, fl._dc_type) for d in fl._payment_dts]
        f_alpha += list(fl._year_fracs)
        l_start_t += [times_from_dates(d, vd, ft._dc_type) for d in ft._start_accrued_dts]
        l_end_t += [times_from_dates(d, vd, ft._dc_type) for d in ft._end_accrued_dts]
        l_pay_t += [times_from_dates(d, vd, ft._dc_type) for d in ft._payment_dts]
        l_alpha += list(ft._year_fracs)
        fo.append(len(f_pay_t))
        lo.append(len(l_pay_t))
    return dict(fo=np.array(fo), f_pay_t=np.array(f_pay_t), f_alpha=np.array(f_alpha), lo=np.array(lo),
                l_start_t=np.array(l_start_t), l_end_t=np.array(l_end_t), l_pay_t=np.array(l_pay_t),
                l_alpha=np.array(l_alpha))


def shocked_rate_scenarios(curve: OISCurve, n_scen: int, seed: int = 7, vol_bp: float = 10.0):
    """BASELINE config 4: per-pillar shocks ~ N(0, (10bp)^2) with correlation exp(-|ln(T_i/T_j)|), added to the
    quotes like a Model.scenario dict-shock (models.py:541-545).  Returns par rates in decimal [S, R]."""
    T = np.array(curve.swap_times)
    corr = np.exp(-np.abs(np.log(T[:, None] / T[None, :])))
    chol = np.linalg.cholesky(corr + 1e-12 * np.eye(len(T)))
    rng = np.random.Generator(np.random.PCG64(seed))
    z = rng.standard_normal((n_scen, len(T))) @ chol.T
    return np.array(curve.swap_rates)[None, :] + z * (vol_bp * 1e-4)
