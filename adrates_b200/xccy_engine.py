"""Cross-currency basis swap valuation on the CUDA path (VALUE + three delta ladders).

Replaces Engine._compute_xccy (cavour/market/position/engine.py:1411-1765):
    total = PV_dom(domestic OIS grid) + PV_for(foreign OIS grid for forwards, XCCY nodes for discounting)/spot_fx
    delta_dom   = grad_dom_dfs  @ J_dom      * 1e-4
    delta_for   = grad_for_dfs  @ J_for      * 1e-4 / spot_fx      (XCCY DFs held fixed)
    delta_basis = grad_xccy_dfs @ J_basis    * 1e-4 / spot_fx      (both OIS curves held fixed)
The domestic leg is an ordinary single-curve unit.  The foreign leg lives on a *stacked* node grid
[foreign OIS engine grid ; XCCY curve nodes]: a coupon is the 3-bracket product term
N*DF_f(s)/DF_f(e)*DF_x(p), and the two foreign-leg ladders are the same valuation run against two
Jacobian blocks ([J_for ; 0] and [0 ; J_basis]) uploaded with cav_curve_set_tables.

GAMMA (engine.py:1769-1967): three diagonal blocks, each `J^T H_pv J + sum_k grad_k C_k` of one curve with the other two
held fixed - the domestic leg on the domestic OIS grid, the foreign leg on the stacked grid against the second-order
blocks ([J_for ; 0], [C_for ; 0]) and ([0 ; J_basis], [0 ; H_basis]) - i.e. the same GAMMA valuation the OIS path runs,
three times.  `H_basis` = XccyCurve._hess_basis (second-order forward mode through the XCCY bootstrap).
Cross-gamma foreign OIS x basis (engine.py:1884-1955): `sum_i grad_xccy[i] * mixed[i, k, j] * d(foreign node DF j)/d(rate l)`.
The reference contracts `_mixed_hess_foreign_basis` (indexed by the foreign curve's 65 path-A nodes) with the ENGINE-grid
Jacobian (263 rows) and raises on the shape mismatch (reproduced when the goldens were generated), so its XCCY GAMMA request
never returns.  Here the contraction uses the Jacobian of those same path-A nodes (OISCurve.path_a_jacobian), which is what
the formula calls for; the result is flagged `beyond_reference` and validated by finite differences instead of goldens.
The contraction itself is a fourth GAMMA valuation with J = 0 and the per-node matrices `sum_j mixed[i,k,j] J_A[j,l]` in
place of the curve Hessian.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .curves import plan_queries
from .dates import times_from_dates, to_tenor
from .error import LibError
from .flatten import FlatPortfolio, Flattener, _Unit
from .global_types import CurveTypes, RequestTypes, SwapTypes
from .results import AnalyticsResult, CrossGamma, Delta, Gamma, Risk, Valuation


def _sign(leg) -> float:
    return +1.0 if leg._leg_type == SwapTypes.RECEIVE else -1.0


def domestic_leg_unit(swap, value_dt) -> _Unit:
    """Floating leg with notional exchange on its own curve (engine.py:2675-2728), single-DF terms."""
    leg = swap._domestic_leg
    dc = leg._dc_type
    sg, N = _sign(leg), leg._notional
    t = lambda d: times_from_dates(d, value_dt, dc)  # noqa: E731
    times, amts = [], []
    for i, pay in enumerate(leg._payment_dts):
        tp, ts, te, al = t(pay), t(leg._start_accrued_dts[i]), t(leg._end_accrued_dts[i]), leg._year_fracs[i]
        if not tp >= 0.0:
            continue
        if al > 0:
            if tp == te:
                times += [((ts, 1.0),), ((te, 1.0),)]
                amts += [sg * N, -sg * N]
            else:
                times += [((ts, 1.0), (te, -1.0), (tp, 1.0)), ((tp, 1.0),)]
                amts += [sg * N, -sg * N]
        if leg._spread != 0.0:
            times.append(((tp, 1.0),))
            amts.append(sg * leg._spread * al * N)
    if leg._notional_exchange:
        te_, tm_ = t(swap._effective_dt), t(swap._maturity_dt)
        if te_ >= 0.0:
            times.append(((te_, 1.0),))
            amts.append(-sg * N)
        if tm_ >= 0.0:
            times.append(((tm_, 1.0),))
            amts.append(sg * N)
    return _Unit(times, amts)


def foreign_leg_terms(swap, value_dt, xccy_curve):
    """[(amount, t_start|None, t_end|None, t_pay)]: start/end on the foreign OIS grid (foreign leg day count),
    pay on the XCCY nodes (XCCY day count), amounts in domestic currency (divided by the curve's spot_fx)."""
    leg = swap._foreign_leg
    fx = xccy_curve._spot_fx
    sg, N = _sign(leg), leg._notional
    tf = lambda d: times_from_dates(d, value_dt, leg._dc_type)          # noqa: E731
    tx = lambda d: times_from_dates(d, value_dt, xccy_curve._dc_type)   # noqa: E731
    out = []
    for i, pay in enumerate(leg._payment_dts):
        tp, al = tx(pay), leg._year_fracs[i]
        if not tp >= 0.0:
            continue
        if al > 0:
            out.append((sg * N / fx, tf(leg._start_accrued_dts[i]), tf(leg._end_accrued_dts[i]), tp))
            out.append((-sg * N / fx, None, None, tp))
        if leg._spread != 0.0:
            out.append((sg * leg._spread * al * N / fx, None, None, tp))
    if leg._notional_exchange:
        te_, tm_ = tx(swap._effective_dt), tx(swap._maturity_dt)
        if te_ >= 0.0:
            out.append((-sg * N / fx, None, None, te_))
        if tm_ >= 0.0:
            out.append((sg * N / fx, None, None, tm_))
    return out


def flatten_foreign_legs(swaps, value_dt, foreign_curve, xccy_curve) -> FlatPortfolio:
    return _flatten_stacked([foreign_leg_terms(sw, value_dt, xccy_curve) for sw in swaps], foreign_curve, xccy_curve)


def _flatten_stacked(term_lists, foreign_curve, xccy_curve) -> FlatPortfolio:
    """One private unit per trade on the stacked grid; 6 (node, weight) pairs per term:
    [start bracket (+), end bracket (-), pay bracket (+ on XCCY nodes, offset by the OIS grid size)]."""
    plan_f = foreign_curve.path_b_plan()
    Gf = plan_f.n_nodes
    terms, offsets = [], [0]
    for tl in term_lists:
        terms += tl
        offsets.append(len(terms))
    n = len(terms)
    amt = np.array([t[0] for t in terms], dtype=np.float64)
    has_fwd = np.array([t[1] is not None for t in terms], dtype=bool)
    ts = np.array([t[1] if t[1] is not None else 0.0 for t in terms], dtype=np.float64)
    te = np.array([t[2] if t[2] is not None else 0.0 for t in terms], dtype=np.float64)
    tp = np.array([t[3] for t in terms], dtype=np.float64)
    weight = np.zeros((n, 6))
    node = np.zeros((n, 6), dtype=np.int32)
    if n:
        a, b, wa, wb = plan_queries(ts, plan_f.node_time, foreign_curve._interp_type)
        weight[:, 0], weight[:, 1], node[:, 0], node[:, 1] = wa * has_fwd, wb * has_fwd, a, b
        a, b, wa, wb = plan_queries(te, plan_f.node_time, foreign_curve._interp_type)
        weight[:, 2], weight[:, 3], node[:, 2], node[:, 3] = -wa * has_fwd, -wb * has_fwd, a, b
        a, b, wa, wb = plan_queries(tp, np.asarray(xccy_curve._times, dtype=np.float64), xccy_curve._interp_type)
        weight[:, 4], weight[:, 5], node[:, 4], node[:, 5] = wa, wb, a + Gf, b + Gf
        node[weight == 0.0] = 0
    N = len(term_lists)
    return FlatPortfolio(N, n, np.array(offsets, dtype=np.int64), 6, amt, weight.reshape(-1), node.reshape(-1),
                         N, 1, np.ones(N), N, np.arange(N + 1, dtype=np.int64), np.arange(N, dtype=np.int32), None,
                         np.ones(N))


class XccySession:
    """Device tables of the stacked (foreign OIS + XCCY) grid, one context per Jacobian block."""
    _cache = {}

    @classmethod
    def get(cls, foreign_curve, xccy_curve, device=0):
        key = (device, id(foreign_curve), id(xccy_curve))
        if key not in cls._cache:
            if len(cls._cache) >= 8:
                # the oldest session leaves the cache; its contexts are freed with its last reference (a caller may hold it)
                cls._cache.pop(next(iter(cls._cache)))
            cls._cache[key] = XccySession(foreign_curve, xccy_curve, device)
        return cls._cache[key]

    def __init__(self, foreign_curve, xccy_curve, device):
        from .position import CurveSession
        fsess = CurveSession.get(foreign_curve, device)
        d_f, J_f, _ = fsess.ctx.curve_read(jac=True, hess=False)       # engine grid of the foreign OIS curve
        d_x = np.asarray(xccy_curve._dfs, dtype=np.float64)
        J_b = np.asarray(xccy_curve._jac_basis, dtype=np.float64)
        Gf, Rf, Gx, Rb = d_f.shape[0], J_f.shape[1], d_x.shape[0], J_b.shape[1]
        d = np.concatenate([d_f, d_x])
        self.ctx_for = _native.Context(device)
        self.ctx_for.curve_set_tables(d, np.vstack([J_f, np.zeros((Gx, Rf))]))
        self.ctx_basis = _native.Context(device)
        self.ctx_basis.curve_set_tables(d, np.vstack([np.zeros((Gf, Rb)), J_b]))
        self.n_for, self.n_basis = Rf, Rb
        self._foreign_curve, self._xccy_curve, self._device = foreign_curve, xccy_curve, device
        self._stacked = (d, J_f, J_b, Gf, Gx)
        self.ctx_for2 = self.ctx_basis2 = self.ctx_cross = None

    def second_order(self):
        """Contexts with second-order tables (built on the first GAMMA request): foreign OIS block, basis block, and the
        cross block (J = 0, per-node matrices of the foreign x basis contraction: rows = foreign par rates, columns = basis
        pillars)."""
        if self.ctx_for2 is None:
            from .position import CurveSession
            d, J_f, J_b, Gf, Gx = self._stacked
            Rf, Rb = self.n_for, self.n_basis
            _, _, C_f = CurveSession.get(self._foreign_curve, self._device).ctx.curve_read(jac=False, hess=True)
            xc = self._xccy_curve
            H_b = np.asarray(xc._hess_basis, dtype=np.float64)
            self.ctx_for2 = _native.Context(self._device)
            self.ctx_for2.curve_set_tables(d, np.vstack([J_f, np.zeros((Gx, Rf))]), np.concatenate([C_f, np.zeros((Gx, Rf, Rf))]))
            self.ctx_basis2 = _native.Context(self._device)
            self.ctx_basis2.curve_set_tables(d, np.vstack([np.zeros((Gf, Rb)), J_b]), np.concatenate([np.zeros((Gf, Rb, Rb)), H_b]))
            # cross block: A[i, l, k] = sum_j mixed[i, k, j] * J_A[j, l]   (i: XCCY node, k: basis pillar, l: foreign par rate)
            J_A = self._foreign_curve.path_a_jacobian()                         # [nf, Rf]
            mixed = np.asarray(xc._mixed_hess_foreign_basis, dtype=np.float64)   # [Gx, Rb, nf]
            A = np.einsum("ikj,jl->ilk", mixed, J_A)                             # [Gx, Rf, Rb]
            R = max(Rf, Rb)
            Apad = np.zeros((Gf + Gx, R, R))
            Apad[Gf:, :Rf, :Rb] = A
            self.ctx_cross = _native.Context(self._device)
            self.ctx_cross.curve_set_tables(d, np.zeros((Gf + Gx, R)), Apad)
        return self.ctx_for2, self.ctx_basis2, self.ctx_cross


def compute_xccy(derivatives, model, request_list, device=0) -> AnalyticsResult:
    reqs = set(request_list)
    if RequestTypes.CASHFLOWS in reqs:
        raise NotImplementedError("CASHFLOWS on XCCY swaps raises NameError in the reference (engine.py:1986)")
    from .position import CurveSession
    d0 = derivatives[0]
    curves = model.curves
    try:
        dom = getattr(curves, d0._domestic_floating_index.name)
        forn = getattr(curves, d0._foreign_floating_index.name)
    except AttributeError as ex:
        raise LibError(str(ex))
    name = f"{d0._foreign_currency.name}_{d0._domestic_currency.name}_BASIS"
    try:
        xc = getattr(curves, name)
    except AttributeError:
        raise LibError(f"XCCY curve {name} not found in model.")
    vd = model.value_dt
    want_delta, want_gamma = RequestTypes.DELTA in reqs, RequestTypes.GAMMA in reqs
    mask = _native.REQ_VALUE | (_native.REQ_DELTA if want_delta else 0)
    gmask = mask | (_native.REQ_GAMMA if want_gamma else 0)
    # domestic legs: ordinary single-curve units on the domestic OIS curve
    dsess = CurveSession.get(dom, device)
    fl = Flattener(dom)
    for sw in derivatives:
        fl.add_components([(("XD", id(sw)), domestic_leg_unit(sw, vd), 1.0)])
    dsess.ctx.portfolio_upload(fl.finalize(dedup=False))
    agg_dom = dsess.ctx.portfolio_value_host(gmask)
    # foreign legs on the stacked grid
    xs = XccySession.get(forn, xc, device)
    flat = flatten_foreign_legs(derivatives, vd, forn, xc)
    if want_gamma:
        ctx_for, ctx_basis, ctx_cross = xs.second_order()
    else:
        ctx_for, ctx_basis, ctx_cross = xs.ctx_for, xs.ctx_basis, None
    ctx_for.portfolio_upload(flat)
    agg_for = ctx_for.portfolio_value_host(gmask)
    value = delta = gamma = None
    ccy = d0._domestic_currency
    Rd, Rf, Rb = len(dom.swap_rates), xs.n_for, xs.n_basis
    t_dom, t_for, t_bas = to_tenor(dom.swap_times), to_tenor(forn.swap_times), to_tenor(xc.swap_times)
    if RequestTypes.VALUE in reqs:
        value = Valuation(float(agg_dom[0] + agg_for[0]), ccy)
    if want_delta or want_gamma:
        ctx_basis.portfolio_upload(flat)
        agg_bas = ctx_basis.portfolio_value_host(gmask)
    if want_delta:
        delta = Risk([
            Delta(np.array(agg_dom[1:1 + Rd]), t_dom, ccy, d0._domestic_floating_index),
            Delta(np.array(agg_for[1:1 + Rf]), t_for, ccy, d0._foreign_floating_index),
            Delta(np.array(agg_bas[1:1 + Rb]), t_bas, ccy, CurveTypes.USD_GBP_BASIS),
        ])
    if want_gamma:
        ctx_cross.portfolio_upload(flat)
        agg_x = ctx_cross.portfolio_value_host(_native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA)
        block = lambda agg, r, c: np.array(agg[33:].reshape(32, 32)[:r, :c])      # noqa: E731
        cross = CrossGamma(block(agg_x, Rf, Rb), t_for, t_bas, d0._foreign_floating_index, CurveTypes.USD_GBP_BASIS, ccy)
        gamma = Risk([
            Gamma(block(agg_dom, Rd, Rd), t_dom, ccy, d0._domestic_floating_index),
            Gamma(block(agg_for, Rf, Rf), t_for, ccy, d0._foreign_floating_index),
            Gamma(block(agg_bas, Rb, Rb), t_bas, ccy, CurveTypes.USD_GBP_BASIS),
        ], cross_gammas=[cross])
        gamma.beyond_reference = ("cross_gamma(foreign OIS, basis): the reference raises in this contraction "
                                  "(engine.py:1936-1939); computed on the foreign curve's path-A nodes")
    return AnalyticsResult(value=value, risk=delta, gamma=gamma)


# ======================================================================================
# OIS discounted on an XCCY curve (cross-currency collateral)
# ======================================================================================
def ois_collateral_terms(swap, value_dt, fx):
    """Engine._compute_ois_xccy_collateral (engine.py:217-345): fixed coupons discounted on the XCCY nodes,
    floating coupons projected on the OIS engine grid and discounted on the XCCY nodes, everything divided by
    the curve's spot_fx.  All times use the FIXED leg's day count, as the reference does (:263)."""
    fl, ft = swap._fixed_leg, swap._float_leg
    dc = fl._dc_type
    t = lambda d: times_from_dates(d, value_dt, dc)  # noqa: E731
    out = []
    sg = _sign(fl)
    for pay, al in zip(fl._payment_dts, fl._year_fracs):
        tp = t(pay)
        if tp > 0.0:
            out.append((sg * fl._cpn * al * fl._notional / fx, None, None, tp))
    if fl._principal != 0.0 and t(fl._payment_dts[-1]) > 0.0:
        out.append((sg * fl._principal / fx, None, None, t(fl._payment_dts[-1])))
    sg, N = _sign(ft), ft._notional
    for i, pay in enumerate(ft._payment_dts):
        tp, al = t(pay), ft._year_fracs[i]
        if not tp >= 0.0:
            continue
        if al > 0:
            out.append((sg * N / fx, t(ft._start_accrued_dts[i]), t(ft._end_accrued_dts[i]), tp))
            out.append((-sg * N / fx, None, None, tp))
        if ft._spread != 0.0:
            out.append((sg * ft._spread * al * N / fx, None, None, tp))
    return out


def compute_ois_xccy_collateral(derivative, model, request_list, collateral_ccy, device=0) -> AnalyticsResult:
    reqs = set(request_list)
    if RequestTypes.GAMMA in reqs:
        raise NotImplementedError("GAMMA not yet supported for OIS with cross-currency collateral. "
                                  "Only VALUE and DELTA are currently implemented.")   # engine.py:491-495
    curves = model.curves
    ois = getattr(curves, derivative._floating_index.name)
    name = f"{derivative._currency.name}_{collateral_ccy.name}_XCCY"
    try:
        xc = getattr(curves, name)
    except AttributeError:
        raise LibError(f"XCCY curve {name} not found in model. Required for cross-currency collateral valuation.")
    vd = model.value_dt
    xs = XccySession.get(ois, xc, device)
    terms = ois_collateral_terms(derivative, vd, xc._spot_fx)
    flat = _flatten_stacked([terms], ois, xc)
    mask = _native.REQ_VALUE | (_native.REQ_DELTA if RequestTypes.DELTA in reqs else 0)
    xs.ctx_for.portfolio_upload(flat)
    agg_o = xs.ctx_for.portfolio_value_host(mask)
    value = Valuation(float(agg_o[0]), collateral_ccy) if RequestTypes.VALUE in reqs else None
    delta = None
    if RequestTypes.DELTA in reqs:
        xs.ctx_basis.portfolio_upload(flat)
        agg_b = xs.ctx_basis.portfolio_value_host(mask)
        delta = Risk([
            Delta(np.array(agg_o[1:1 + xs.n_for]), to_tenor(ois.swap_times), collateral_ccy, derivative._floating_index),
            Delta(np.array(agg_b[1:1 + xs.n_basis]), to_tenor(xc.swap_times), collateral_ccy, CurveTypes.USD_GBP_BASIS),
        ])
    cashflows = None
    if RequestTypes.CASHFLOWS in reqs:          # the reference's placeholder: an empty table (engine.py:497-501)
        from .cashflows import Cashflows
        cashflows = Cashflows([], derivative._currency)
    return AnalyticsResult(value=value, risk=delta, gamma=None, cashflows=cashflows)


# ======================================================================================
# Dual-curve floating-rate notes (VALUE only, like the reference)
# ======================================================================================
class _EngineGrid:
    """The path-B grid of an OIS curve in the shape _flatten_stacked expects of its second (discounting) grid."""

    def __init__(self, curve, d):
        self._times = curve.path_b_plan().node_time
        self._dfs = d
        self._interp_type = curve._interp_type


_dual_sessions = {}


def value_frn_dual_curve(frn, model, device=0) -> AnalyticsResult:
    """Engine._compute_frn with an index curve that is not the currency's discount curve (engine.py:700-866):
    coupons N a_i ((DF_idx(s)/DF_idx(e) - 1)/a_i + margin) DF_disc(p) plus the face value at maturity, as product
    terms on the stacked grid [index engine grid ; discount engine grid].

    Reference quirk kept: the engine caches curve tables under tuple(swap_times) alone (engine.py:2362-2376) and
    fetches the discount curve first, so an index curve quoting the SAME pillar dates resolves to the discount
    curve's tables and the note is valued single-curve."""
    from .credit import BOND_CURVE
    from .position import CurveSession, value_positions
    try:
        disc = getattr(model.curves, BOND_CURVE[frn._currency].name)
        idx = getattr(model.curves, frn._floating_index.name)
    except AttributeError as ex:
        raise LibError(str(ex))
    if tuple(idx.swap_times) == tuple(disc.swap_times):
        return value_positions([frn], disc, [RequestTypes.VALUE], device)
    key = (device, id(idx), id(disc))
    sess = _dual_sessions.get(key)
    if sess is None:
        if len(_dual_sessions) >= 8:
            _dual_sessions.pop(next(iter(_dual_sessions)))
        d_i, _, _ = CurveSession.get(idx, device).ctx.curve_read(jac=False, hess=False)
        d_d, _, _ = CurveSession.get(disc, device).ctx.curve_read(jac=False, hess=False)
        ctx = _native.Context(device)
        ctx.curve_set_tables(np.concatenate([d_i, d_d]))
        sess = _dual_sessions[key] = (ctx, _EngineGrid(disc, d_d), idx, disc)     # curves kept: ids stay unique
    ctx, grid = sess[0], sess[1]
    vd = model.value_dt
    flat = _flatten_stacked([_dual_frn_terms(frn, vd)], idx, grid)
    ctx.portfolio_upload(flat)
    agg = ctx.portfolio_value_host(_native.REQ_VALUE)
    return AnalyticsResult(value=Valuation(float(agg[0]), frn._currency), risk=None, gamma=None)


def _dual_frn_terms(frn, value_dt):
    """[(amount, t_start|None, t_end|None, t_pay)] of a dual-curve note: the cashflow rules of flatten.frn_components
    (engine.py:763-866), but every forward keeps its own product term - DF(s) - DF(e) only collapses when projection
    and discounting share a curve."""
    dc = frn._dc_type
    t = lambda d: float(times_from_dates(d, value_dt, dc))  # noqa: E731
    N, m = float(frn._face_value), float(frn._quoted_margin)
    out = []
    for i, al in enumerate(frn._year_fracs):
        tp = t(frn._payment_dts[i])
        if not tp >= 0.0:
            continue
        if i == 0 and frn._first_fixing_rate is not None:
            out.append(((float(frn._first_fixing_rate) + m) * al * N, None, None, tp))
            continue
        if al > 0:
            out.append((N, t(frn._start_accrued_dts[i]), t(frn._end_accrued_dts[i]), tp))
            out.append((-N, None, None, tp))
        if m != 0.0:
            out.append((m * al * N, None, None, tp))
    tm = t(frn._maturity_dt)
    if tm > 0.0:
        out.append((N, None, None, tm))
    return out
