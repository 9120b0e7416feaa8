"""Cashflow-schedule flattener: trades -> SoA arrays for the CUDA valuation kernels.

Replaces the per-trade host preparation inside the reference engine
(cavour/market/position/engine.py:2519-2539 fixed leg, :2858-2897 floating leg):
payment/start/end dates -> year fractions with the leg's own day count
(times_from_dates, cavour/utils/helpers.py:154-197), past-cashflow masks (`>` for fixed,
`>=` for floating: engine.py:2430 vs :2695), leg signs.

Layout (see include/adrates_b200.h, cav_portfolio_upload):
  unit   = set of terms  amt * prod_m DF-bracket  on the engine grid; each term carries
           n_pairs (node, weight) pairs produced by curves.plan_queries
  trade  = sum_k comp_weight[k] * unit[k]
Vanilla fixed-float swaps are linear in (coupon*notional, notional, spread*notional) once
the schedule is fixed, so with dedup=True trades that share a schedule share its units
(annuity, floating, spread-annuity) and differ only in three weights; the Greeks of a unit
are computed once and expanded per trade by a bandwidth-bound kernel.  dedup=False gives
every trade a private unit (the general form for irregular trades).

Floating coupons with payment_lag == 0 pay on their accrual end date, so
((DF(s)/DF(e) - 1)/a + spread) * a * N * DF(e) = N (DF(s) - DF(e)) + spread a N DF(e):
two single-DF terms.  With a payment lag the product DF(s) DF(p)/DF(e) is kept as one
6-pair term.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .curves import OISCurve, plan_queries
from .dates import times_from_dates
from .error import LibError
from .global_types import InstrumentTypes, SwapTypes


@dataclass
class FlatPortfolio:
    n_units: int
    n_terms: int
    unit_offsets: np.ndarray      # int64 [n_units+1]
    n_pairs: int
    amt: np.ndarray               # f64 [n_terms]
    weight: np.ndarray            # f64 [n_terms*n_pairs]
    node: np.ndarray              # i32 [n_terms*n_pairs]
    n_trades: int
    n_comp: int
    comp_weight: np.ndarray       # f64 [n_trades*n_comp]
    n_groups: int
    group_offsets: np.ndarray     # int64 [n_groups+1]
    group_units: np.ndarray       # i32 [n_groups*n_comp]
    out_index: Optional[np.ndarray]   # int64 [n_trades] or None
    unit_weight: np.ndarray       # f64 [n_units]
    tile_plan: Optional[object] = None   # adrates_b200.tiles.TilePlan: enables the tensor-core Greeks kernel

    def h2d_bytes(self) -> int:
        arrs = [self.unit_offsets, self.amt, self.weight, self.node, self.comp_weight, self.group_offsets,
                self.group_units, self.unit_weight] + ([self.out_index] if self.out_index is not None else [])
        tp = self.tile_plan
        if tp is not None:
            arrs += [tp.tile_units, tp.tile_kstart, tp.tile_kcount, tp.k_row, tp.k_pos, tp.k_coef, tp.k_pos2, tp.k_coef2, tp.pairs, tp.tile_mask]
        return int(sum(a.nbytes for a in arrs))

    def with_tiles(self, n_nodes: int, plan=None) -> "FlatPortfolio":
        """Attach a tile plan (adrates_b200/tiles.py) when every unit consists of single-DF terms.
        plan: the curve's PathBPlan; its dependency structure gives the active par-rate pillars of each tile
        (column compaction of the tile GEMM).  Without it every pillar is treated as active."""
        from .tiles import plan_tiles, node_support_masks
        if self.n_pairs == 2 and self.n_units > 0:
            support = None if plan is None else node_support_masks(plan.node_swap, plan.node_prev, plan.node_acc)
            tp = plan_tiles(self, n_nodes, support=support)
            if len(tp.leftover_units) == 0:
                self.tile_plan = tp
        return self


# a DF query list: times on the curve, each with an exponent (+1 / -1) inside one term
@dataclass
class _Unit:
    times: list        # per term: tuple of (time, exponent) pairs, 1 or 3 entries
    amts: list


def _leg_sign(leg) -> float:
    return +1.0 if leg._leg_type == SwapTypes.RECEIVE else -1.0


def _times(dts, value_dt, dc_type):
    """Year fractions of a list of dates as a list of Python floats (one vectorised call for the simple day counts)."""
    return [float(x) for x in times_from_dates(list(dts), value_dt, dc_type)] if len(dts) else []


def ois_components(swap, value_dt):
    """[(key, unit, weight)] for one OIS (engine.py:153-189 = fixed leg + floating leg)."""
    if getattr(swap, "derivative_type", None) != InstrumentTypes.OIS_SWAP:
        raise LibError(f"{getattr(swap, 'derivative_type', type(swap))} not yet implemented")
    out = []
    fl, ft = swap._fixed_leg, swap._float_leg
    # ---- fixed leg: sum_{t_i > t_val} alpha_i N c DF(t_i) (+ principal, 0 for OIS) ----
    val_t = times_from_dates(value_dt, value_dt, fl._dc_type)
    t_pay = _times(fl._payment_dts, value_dt, fl._dc_type)
    live = [i for i, t in enumerate(t_pay) if t > val_t]
    if live:
        times = tuple(t_pay[i] for i in live)
        alphas = tuple(fl._year_fracs[i] for i in live)
        unit = _Unit([((t, 1.0),) for t in times], list(alphas))
        out.append((("A", times, alphas), unit, _leg_sign(fl) * fl._notional * fl._cpn))
        if fl._principal != 0.0:
            unit_p = _Unit([((times[-1], 1.0),)], [1.0])
            out.append((("P", times[-1]), unit_p, _leg_sign(fl) * fl._principal))
    # ---- floating leg ----
    val_t = times_from_dates(value_dt, value_dt, ft._dc_type)
    tp = _times(ft._payment_dts, value_dt, ft._dc_type)
    ts = _times(ft._start_accrued_dts, value_dt, ft._dc_type)
    te = _times(ft._end_accrued_dts, value_dt, ft._dc_type)
    al = list(ft._year_fracs)
    live = [i for i, t in enumerate(tp) if t >= val_t]
    notional = ft._notional
    if ft._notional_array and any(n != notional for n in ft._notional_array):
        raise LibError("amortising notionals need a private unit (dedup=False path)")
    if live:
        sgn = _leg_sign(ft)
        fwd_terms, fwd_amts = [], []
        for i in live:
            if al[i] > 0:
                if tp[i] == te[i]:
                    fwd_terms += [((ts[i], 1.0),), ((te[i], 1.0),)]
                    fwd_amts += [1.0, -1.0]
                else:
                    fwd_terms += [((ts[i], 1.0), (te[i], -1.0), (tp[i], 1.0)), ((tp[i], 1.0),)]
                    fwd_amts += [1.0, -1.0]
        if fwd_terms:
            key = ("F", tuple(ts[i] for i in live), tuple(te[i] for i in live), tuple(tp[i] for i in live))
            out.append((key, _Unit(fwd_terms, fwd_amts), sgn * notional))
        if ft._spread != 0.0:
            key = ("S", tuple(tp[i] for i in live), tuple(al[i] for i in live))
            out.append((key, _Unit([((tp[i], 1.0),) for i in live], [al[i] for i in live]), sgn * notional * ft._spread))
        if ft._principal != 0.0:
            out.append((("P", tp[live[-1]]), _Unit([((tp[live[-1]], 1.0),)], [1.0]), sgn * ft._principal))
    return out


def _merge_single_df_terms(unit: _Unit):
    """Sum amounts of single-DF terms that hit the same time; drop exact zeros."""
    single = {}
    multi_t, multi_a = [], []
    for q, a in zip(unit.times, unit.amts):
        if len(q) == 1 and q[0][1] == 1.0:
            single[q[0][0]] = single.get(q[0][0], 0.0) + a
        else:
            multi_t.append(q)
            multi_a.append(a)
    ts = sorted(t for t, a in single.items() if a != 0.0)
    return _Unit([((t, 1.0),) for t in ts] + multi_t, [single[t] for t in ts] + multi_a)


def bond_components(bond, value_dt):
    """Engine._compute_bond (engine.py:505-575): `_price_fixed_leg_jax` with payments = coupon payments, principal =
    face value on the LAST payment date, leg_sign = +1 (investor side), times in the bond's day count; flows on or
    before the value date are worth 0."""
    ts = _times(bond._payment_dts, value_dt, bond._dc_type)
    live = [i for i, t in enumerate(ts) if t > 0.0]
    terms = [((ts[i], 1.0),) for i in live]
    amts = [float(bond._coupon_payments[i]) for i in live]
    if ts and ts[-1] > 0.0:
        terms.append(((ts[-1], 1.0),))
        amts.append(float(bond._face_value))
    if not terms:
        return []
    key = ("B", tuple(t[0][0] for t in terms), tuple(amts))
    return [(key, _Unit(terms, amts), 1.0)]


def frn_components(frn, value_dt):
    """Engine._compute_frn, single-curve case (engine.py:763-860): `_float_leg_jax` with spreads = quoted margin,
    notionals = face value, leg_sign = +1, an optional known first fixing on coupon 0, plus the face value
    discounted from the (adjusted) maturity date; times in the FRN's day count, coupons with payment time >= 0."""
    dc = frn._dc_type
    tp = _times(frn._payment_dts, value_dt, dc)
    ts = _times(frn._start_accrued_dts, value_dt, dc)
    te = _times(frn._end_accrued_dts, value_dt, dc)
    N, m = float(frn._face_value), float(frn._quoted_margin)
    terms, amts = [], []
    for i, al in enumerate(frn._year_fracs):
        if not tp[i] >= 0.0:
            continue
        if i == 0 and frn._first_fixing_rate is not None:
            terms.append(((tp[i], 1.0),))
            amts.append((float(frn._first_fixing_rate) + m) * al * N)
            continue
        if al > 0:
            if tp[i] == te[i]:
                terms += [((ts[i], 1.0),), ((te[i], 1.0),)]
            else:
                terms += [((ts[i], 1.0), (te[i], -1.0), (tp[i], 1.0)), ((tp[i], 1.0),)]
            amts += [N, -N]
        if m != 0.0:
            terms.append(((tp[i], 1.0),))
            amts.append(m * al * N)
    tm = float(times_from_dates(frn._maturity_dt, value_dt, dc))
    if tm > 0.0:
        terms.append(((tm, 1.0),))
        amts.append(N)
    if not terms:
        return []
    return [(("N", id(frn)), _Unit(terms, amts), 1.0)]


def trade_components(derivative, value_dt):
    kind = getattr(derivative, "derivative_type", None)
    if kind == InstrumentTypes.BOND:
        return bond_components(derivative, value_dt)
    if kind == InstrumentTypes.FRN:
        return frn_components(derivative, value_dt)
    return ois_components(derivative, value_dt)


class Flattener:
    """Collects trades on one curve and emits a FlatPortfolio."""

    def __init__(self, curve: OISCurve):
        self.curve = curve
        self.value_dt = curve._value_dt
        self._trades = []   # list of [(key, unit, weight)]

    def add_trade(self, swap):
        self._trades.append(trade_components(swap, self.value_dt))

    def add_components(self, comps):
        self._trades.append(comps)

    def finalize(self, dedup: bool = True, max_group: int = 256) -> FlatPortfolio:
        n_trades = len(self._trades)
        units, unit_ids = [], {}
        trade_comps = []
        if dedup:
            for comps in self._trades:
                row = []
                for key, unit, w in comps:
                    uid = unit_ids.get(key)
                    if uid is None:
                        uid = unit_ids[key] = len(units)
                        units.append(_merge_single_df_terms(unit))
                    row.append((uid, w))
                trade_comps.append(row)
        else:
            for comps in self._trades:
                big = _Unit([], [])
                for _, unit, w in comps:
                    big.times += unit.times
                    big.amts += [a * w for a in unit.amts]
                units.append(_merge_single_df_terms(big))
                trade_comps.append([(len(units) - 1, 1.0)])
        return assemble(self.curve, units, trade_comps, n_trades, direct=not dedup, max_group=max_group)


def assemble(curve: OISCurve, units, trade_comps, n_trades, direct=False, max_group=256) -> FlatPortfolio:
    """Plan every DF query of every unit in one vectorised call and lay out the arrays."""
    plan = curve.path_b_plan()
    n_pairs = 2
    for u in units:
        if any(len(q) > 1 for q in u.times):
            n_pairs = 6
            break
    slots = n_pairs // 2
    counts = np.array([len(u.amts) for u in units], dtype=np.int64)
    unit_offsets = np.zeros(len(units) + 1, dtype=np.int64)
    np.cumsum(counts, out=unit_offsets[1:])
    n_terms = int(unit_offsets[-1])
    amt = np.empty(n_terms)
    q_time = np.zeros((n_terms, slots))
    q_exp = np.zeros((n_terms, slots))
    k = 0
    for u in units:
        for q, a in zip(u.times, u.amts):
            amt[k] = a
            for s, (t, e) in enumerate(q):
                q_time[k, s] = t
                q_exp[k, s] = e
            k += 1
    a_idx, b_idx, wa, wb = plan_queries(q_time.reshape(-1), plan.node_time, curve._interp_type)
    ex = q_exp.reshape(-1)
    weight = np.stack([wa * ex, wb * ex], axis=1).reshape(n_terms, n_pairs)
    node = np.stack([a_idx, b_idx], axis=1).reshape(n_terms, n_pairs).astype(np.int32)
    node[weight == 0.0] = 0
    return layout_trades(len(units), unit_offsets, n_pairs, amt, weight.reshape(-1), node.reshape(-1), trade_comps,
                         n_trades, direct, max_group)


def layout_trades(n_units, unit_offsets, n_pairs, amt, weight, node, trade_comps, n_trades, direct, max_group):
    n_terms = int(unit_offsets[-1])
    if direct:
        comp_weight = np.ones(n_trades)
        group_offsets = np.arange(n_trades + 1, dtype=np.int64)
        group_units = np.arange(n_trades, dtype=np.int32)
        return FlatPortfolio(n_units, n_terms, unit_offsets, n_pairs, amt, weight, node, n_trades, 1, comp_weight,
                             n_trades, group_offsets, group_units, None, np.ones(n_units))
    n_comp = max([len(r) for r in trade_comps] + [1])
    if n_comp > 4:
        raise LibError("a trade decomposes into more than 4 units")
    ids = np.zeros((n_trades, n_comp), dtype=np.int32)
    ws = np.zeros((n_trades, n_comp))
    for t, row in enumerate(trade_comps):
        for k, (uid, w) in enumerate(row):
            ids[t, k], ws[t, k] = uid, w
        for k in range(len(row), n_comp):
            ids[t, k] = row[0][0] if row else 0   # unused slot: weight 0 on an existing unit
    return group_trades(n_units, unit_offsets, n_pairs, amt, weight, node, ids, ws, max_group)


def group_trades(n_units, unit_offsets, n_pairs, amt, weight, node, ids, ws, max_group=256, run_key=None) -> FlatPortfolio:
    """Sort trades by their unit ids so that a CTA of the expansion kernel serves one run.
    run_key (int array per trade, optional): trades are ordered by this key instead (stable) and a run is a stretch of equal
    keys - the schedule class of array books, which is also how the device flattener (cav_book_from_arrays) lays trades out;
    every trade of a run must carry the same unit ids."""
    n_trades, n_comp = ids.shape
    n_terms = int(unit_offsets[-1])
    if not n_trades:
        order = np.zeros(0, dtype=np.int64)
    elif run_key is not None:
        order = np.argsort(run_key, kind="stable")
    else:
        order = np.lexsort(tuple(ids[:, k] for k in reversed(range(n_comp))))
    ids_s, ws_s = ids[order], ws[order]
    if n_trades:
        new_run = np.ones(n_trades, dtype=bool)
        if run_key is not None:
            key_s = np.asarray(run_key)[order]
            new_run[1:] = key_s[1:] != key_s[:-1]
        else:
            new_run[1:] = np.any(ids_s[1:] != ids_s[:-1], axis=1)
        run_start = np.flatnonzero(new_run)
        run_end = np.append(run_start[1:], n_trades)
        per_run = -(-(run_end - run_start) // max_group)            # groups per run of equal unit ids
        first = np.cumsum(per_run) - per_run
        owner = np.repeat(np.arange(run_start.shape[0]), per_run)
        starts = run_start[owner] + (np.arange(int(per_run.sum())) - first[owner]) * max_group
        group_offsets = np.append(starts, n_trades).astype(np.int64)
        group_units = ids_s[starts].astype(np.int32).reshape(-1)
    else:
        group_offsets = np.zeros(1, dtype=np.int64)
        group_units = np.zeros(0, dtype=np.int32)
    unit_weight = np.zeros(n_units)
    for k in range(n_comp):
        unit_weight += np.bincount(ids_s[:, k], weights=ws_s[:, k], minlength=n_units)
    return FlatPortfolio(n_units, n_terms, unit_offsets, n_pairs, amt, weight, node, n_trades, n_comp,
                         np.ascontiguousarray(ws_s.reshape(-1)), len(group_offsets) - 1, group_offsets, group_units,
                         order.astype(np.int64), unit_weight)
