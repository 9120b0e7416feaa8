"""Synthetic cross-currency basis-swap books (BASELINE.json config 5, XCCY half) and their batched valuation.

Book: GBP/USD basis swaps, tenor ~ U{1..30} years, effective date = value date (or value
date + U{1..max_offset_bd} business days for half of the trades when max_offset_bd > 0), foreign leg quarterly
ACT/365F on SONIA, domestic leg annual ACT/360 on SOFR (conventions of the reference's
tests/test_engine_basis_swap.py:82-97 as mirrored in tests/util_xccy.py), foreign basis spread ~ N(-10 bp, 15 bp),
foreign notional ~ logU[1e5, 1e8], domestic notional = foreign notional x spot; numpy Generator(PCG64), seed 11.

A basis swap is linear in (domestic notional, domestic notional x spread, foreign notional, foreign notional x
spread) once its dates are fixed, so the book shares *units* per distinct (tenor, start): the float-plus-exchange
unit and the spread annuity of each leg.  The domestic units live on the domestic OIS grid, the foreign units on
the stacked [foreign OIS grid ; XCCY nodes] grid of xccy_engine.py, and the three ladders of
Engine._compute_xccy (engine.py:1411-1765) are three valuations of the same flat book against the Jacobian blocks
J_dom, [J_for ; 0] and [0 ; J_basis].  Per-trade rows are expanded on the device (k_expand_rows).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _native
from .dates import BusDayAdjustTypes, DayCountTypes, FrequencyTypes, times_from_dates
from .flatten import FlatPortfolio, _Unit, assemble, group_trades
from .global_types import CurrencyTypes, CurveTypes
from .trades import XccyBasisSwap
from .xccy_engine import XccySession, _flatten_stacked, _sign, domestic_leg_unit, foreign_leg_terms


@dataclass
class XccyBook:
    model: object
    templates: list             # XccyBasisSwap with unit notionals and zero spreads, one per distinct schedule
    sched: np.ndarray           # int32 [N] template of each trade
    dom_notional: np.ndarray    # f64 [N]
    for_notional: np.ndarray
    dom_spread: np.ndarray
    for_spread: np.ndarray

    @property
    def n_trades(self) -> int:
        return int(self.sched.shape[0])

    def trade(self, i: int) -> XccyBasisSwap:
        """The i-th trade as an ordinary object (for the per-trade path / parity checks)."""
        t = self.templates[int(self.sched[i])]
        return _swap(t._effective_dt, t._termination_dt, float(self.dom_notional[i]), float(self.for_notional[i]),
                     float(self.dom_spread[i]), float(self.for_spread[i]))


def _swap(effective_dt, end, n_dom, n_for, s_dom, s_for) -> XccyBasisSwap:
    return XccyBasisSwap(effective_dt=effective_dt, term_dt_or_tenor=end, domestic_notional=n_dom,
                         foreign_notional=n_for, domestic_spread=s_dom, foreign_spread=s_for,
                         domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.QUARTERLY,
                         domestic_dc_type=DayCountTypes.ACT_360, foreign_dc_type=DayCountTypes.ACT_365F,
                         domestic_floating_index=CurveTypes.USD_OIS_SOFR, foreign_floating_index=CurveTypes.GBP_OIS_SONIA,
                         domestic_currency=CurrencyTypes.USD, foreign_currency=CurrencyTypes.GBP,
                         domestic_bd_type=BusDayAdjustTypes.FOLLOWING, foreign_bd_type=BusDayAdjustTypes.FOLLOWING)


def make_xccy_book(model, n_trades: int, seed: int = 11, max_tenor: int = 30, max_offset_bd: int = 0,
                   spot: float = 1.25) -> XccyBook:
    rng = np.random.Generator(np.random.PCG64(seed))
    tenor = rng.integers(1, max_tenor + 1, n_trades)
    offset = np.zeros(n_trades, dtype=np.int64)
    if max_offset_bd > 0:
        offset = np.where(rng.random(n_trades) < 0.5, 0, rng.integers(1, max_offset_bd + 1, n_trades))
    for_notional = np.exp(rng.uniform(np.log(1e5), np.log(1e8), n_trades))
    for_spread = rng.normal(-10e-4, 15e-4, n_trades)
    key = tenor.astype(np.int64) * (max_offset_bd + 1) + offset
    uniq, sched = np.unique(key, return_inverse=True)
    vd = model.value_dt
    templates = []
    for k in uniq:
        T, off = int(k // (max_offset_bd + 1)), int(k % (max_offset_bd + 1))
        eff = vd if off == 0 else vd.add_weekdays(off)
        templates.append(_swap(eff, f"{T}Y", 1.0, 1.0, 0.0, 0.0))
    return XccyBook(model, templates, sched.astype(np.int32), for_notional * spot, for_notional,
                    np.zeros(n_trades), for_spread)


def _spread_unit_domestic(swap, value_dt) -> _Unit:
    leg = swap._domestic_leg
    sg = _sign(leg)
    ts = [times_from_dates(d, value_dt, leg._dc_type) for d in leg._payment_dts]
    keep = [i for i, t in enumerate(ts) if t >= 0.0]
    return _Unit([((ts[i], 1.0),) for i in keep], [sg * leg._year_fracs[i] for i in keep])


def _spread_terms_foreign(swap, value_dt, xccy_curve):
    leg = swap._foreign_leg
    sg, fx = _sign(leg), xccy_curve._spot_fx
    out = []
    for pay, al in zip(leg._payment_dts, leg._year_fracs):
        tp = times_from_dates(pay, value_dt, xccy_curve._dc_type)
        if tp >= 0.0:
            out.append((sg * al / fx, None, None, tp))
    return out


def flatten_xccy_book(book: XccyBook, max_group: int = 256):
    """(flat_domestic on the domestic OIS grid, flat_foreign on the stacked grid); unit 2s = float + exchange of
    template s per unit notional, unit 2s + 1 = its spread annuity per unit (notional x spread)."""
    m = book.model
    vd = m.value_dt
    t0 = book.templates[0]
    dom = getattr(m.curves, t0._domestic_floating_index.name)
    forn = getattr(m.curves, t0._foreign_floating_index.name)
    xc = getattr(m.curves, f"{t0._foreign_currency.name}_{t0._domestic_currency.name}_BASIS")
    S = len(book.templates)
    ids = np.stack([2 * book.sched, 2 * book.sched + 1], axis=1).astype(np.int32)
    # domestic
    units = []
    for t in book.templates:
        units += [domestic_leg_unit(t, vd), _spread_unit_domestic(t, vd)]
    base = assemble(dom, units, [[(0, 1.0)]], 1, direct=True)
    ws = np.stack([book.dom_notional, book.dom_notional * book.dom_spread], axis=1)
    flat_dom = group_trades(2 * S, base.unit_offsets, base.n_pairs, base.amt, base.weight, base.node, ids, ws, max_group)
    # foreign (stacked grid, product terms)
    lists = []
    for t in book.templates:
        lists += [foreign_leg_terms(t, vd, xc), _spread_terms_foreign(t, vd, xc)]
    fb = _flatten_stacked(lists, forn, xc)
    ws = np.stack([book.for_notional, book.for_notional * book.for_spread], axis=1)
    flat_for = group_trades(2 * S, fb.unit_offsets, fb.n_pairs, fb.amt, fb.weight, fb.node, ids, ws, max_group)
    return flat_dom, flat_for, (dom, forn, xc)


class XccyBookValuer:
    """Uploads a flattened XCCY book once and revalues it: per-trade PV and the three delta ladders on the device, and with
    gamma=True also the three per-curve gamma matrices and the foreign x basis cross-gamma matrix of every trade
    (engine.py:1769-1967; four more 32x32 rows per trade)."""

    def __init__(self, book: XccyBook, device: int = 0, stream=None, gamma: bool = False):
        import torch
        from .position import CurveSession
        self.flat_dom, self.flat_for, (dom, forn, xc) = flatten_xccy_book(book)
        self.n = book.n_trades
        self.gamma = gamma
        self.dsess = CurveSession.get(dom, device)
        self.xs = XccySession.get(forn, xc, device)
        if gamma:
            self.ctx_for, self.ctx_basis, self.ctx_cross = self.xs.second_order()
        else:
            self.ctx_for, self.ctx_basis, self.ctx_cross = self.xs.ctx_for, self.xs.ctx_basis, None
        self.ctxs = tuple(c for c in (self.dsess.ctx, self.ctx_for, self.ctx_basis, self.ctx_cross) if c is not None)
        if stream is not None:
            for c in self.ctxs:
                c.set_stream(stream)
        self.dsess.ctx.portfolio_upload(self.flat_dom)
        for c in self.ctxs[1:]:
            c.portfolio_upload(self.flat_for)
        dev = torch.device("cuda", device)
        f64 = dict(dtype=torch.float64, device=dev)
        self.pv_dom, self.pv_for, self.pv_tmp = (torch.empty(self.n, **f64) for _ in range(3))
        self.delta_dom, self.delta_for, self.delta_basis = (torch.empty(self.n, 32, **f64) for _ in range(3))
        self.agg = [torch.zeros(_native.NOUT, **f64) for _ in range(4)]
        if gamma:
            self.delta_tmp = torch.empty(self.n, 32, **f64)
            self.gamma_dom, self.gamma_for, self.gamma_basis, self.gamma_cross = (torch.empty(self.n, 32, 32, **f64) for _ in range(4))

    def value(self):
        """PV (domestic + foreign/spot) and the domestic / foreign / basis ladders (and gammas) of every trade."""
        M = _native.REQ_VALUE | _native.REQ_DELTA | (_native.REQ_GAMMA if self.gamma else 0)
        gp = (lambda t: t.data_ptr()) if self.gamma else (lambda t: None)
        g = (self.gamma_dom, self.gamma_for, self.gamma_basis, self.gamma_cross) if self.gamma else (None,) * 4
        self.dsess.ctx.portfolio_value(M, self.pv_dom.data_ptr(), self.delta_dom.data_ptr(), gp(g[0]), self.agg[0].data_ptr())
        self.ctx_for.portfolio_value(M, self.pv_for.data_ptr(), self.delta_for.data_ptr(), gp(g[1]), self.agg[1].data_ptr())
        self.ctx_basis.portfolio_value(M, self.pv_tmp.data_ptr(), self.delta_basis.data_ptr(), gp(g[2]), self.agg[2].data_ptr())
        if self.gamma:
            self.ctx_cross.portfolio_value(M, self.pv_tmp.data_ptr(), self.delta_tmp.data_ptr(), gp(g[3]), self.agg[3].data_ptr())

    def sync(self):
        for c in self.ctxs:
            c.sync()

    def results(self):
        self.sync()
        return ((self.pv_dom + self.pv_for).cpu().numpy(), self.delta_dom.cpu().numpy(), self.delta_for.cpu().numpy(),
                self.delta_basis.cpu().numpy())
