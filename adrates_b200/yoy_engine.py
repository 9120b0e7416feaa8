"""Year-on-year inflation swaps on the CUDA path: VALUE, and DELTA / GAMMA to the discount curve's par rates and to
the inflation curve's ZCIS breakeven rates.

Replaces Engine._compute_yoy_iis (cavour/market/position/engine.py:986-1350):

    PV        = fixed leg (engine._price_fixed_leg_jax) + sign * sum_{t_p > 0} N a_i (I(e_i)/I(s_i) - 1 + spread) DF(p_i)
    disc risk = grad/hessian w.r.t. the engine-grid DFs (inflation factors held fixed) chained with the OIS
                bootstrap tables:  g J * 1e-4,  (J^T H J + sum_k g_k C_k) * 1e-8
    infl risk = grad/hessian w.r.t. the inflation-curve node factors (DFs held fixed) chained with
                F_k = (1 + b_k)^T_k  (inflation_curve.py:246-301): diagonal Jacobian / Hessian

Both are the generic term valuation of the kernels: with the inflation factors frozen the swap is a strip of fixed
cashflows on the OIS engine grid (single-DF terms); with the discount factors frozen every coupon is the product term
amt * I(e)/I(s) on the *inflation* node grid, whose "bootstrap tables" are the closed-form derivatives of (1 + b)^T
(uploaded with cav_curve_set_tables, like the XCCY node grids).  I(t) follows the same InterpolatorAd rules as DFs
(curves.plan_queries).  The reference has no cross gamma between the two curves (engine.py:1318-1319) and neither has
this.  CPI fixings / lags / seasonality play no role on this route: the engine reads the curve's factors directly
(engine.py:1119-1127).
"""
from __future__ import annotations

import numpy as np

from . import _native
from .curves import plan_queries
from .dates import times_from_dates, to_tenor
from .error import LibError
from .flatten import FlatPortfolio, Flattener, _Unit
from .global_types import CurrencyTypes, CurveTypes, RequestTypes, SwapTypes
from .results import AnalyticsResult, Delta, Gamma, Risk, Valuation

INFLATION_CURVE = {      # engine.py:998-1003
    (CurrencyTypes.GBP, "UK_RPI"): "GBP_RPI_INFLATION",
    (CurrencyTypes.GBP, "UK_CPI"): "GBP_CPI_INFLATION",
    (CurrencyTypes.USD, "US_CPI_U"): "USD_CPI_INFLATION",
    (CurrencyTypes.EUR, "EUR_HICP"): "EUR_HICP_INFLATION",
}
DISCOUNT_CURVE = {       # engine.py:1006-1010
    CurrencyTypes.GBP: "GBP_OIS_SONIA",
    CurrencyTypes.USD: "USD_OIS_SOFR",
    CurrencyTypes.EUR: "EUR_OIS_ESTR",
}


def _sign(leg) -> float:
    return +1.0 if leg._leg_type == SwapTypes.RECEIVE else -1.0


def inflation_tables(curve):
    """Node factors F, dF/db [G, R] and d2F/db2 [G, R, R] of F_k = (1 + b_k)^T_k; node 0 is the constant 1."""
    T = np.asarray(curve.swap_times, dtype=np.float64)
    b = np.array([z._fixed_rate for z in curve._used_swaps], dtype=np.float64)
    R = T.shape[0]
    F = np.concatenate([[1.0], (1.0 + b) ** T])
    J = np.zeros((R + 1, R))
    C = np.zeros((R + 1, R, R))
    k = np.arange(R)
    J[k + 1, k] = T * (1.0 + b) ** (T - 1.0)
    C[k + 1, k, k] = T * (T - 1.0) * (1.0 + b) ** (T - 2.0)
    return F, J, C


class InflationSession:
    """Device tables of one inflation curve's node grid."""
    _cache = {}

    @classmethod
    def get(cls, curve, device=0):
        key = (device, id(curve))
        sess = cls._cache.get(key)
        if sess is None:
            if len(cls._cache) >= 8:
                cls._cache.pop(next(iter(cls._cache))).ctx.close()
            sess = cls._cache[key] = InflationSession(curve, device)
        return sess

    def __init__(self, curve, device):
        if len(curve.swap_times) > 32:
            raise LibError("inflation curves with more than 32 pillars are not supported")
        self.curve = curve                       # keeps id(curve) unique while cached
        self.F, J, C = inflation_tables(curve)
        self.times = np.concatenate([[0.0], np.asarray(curve.swap_times, dtype=np.float64)])
        self.ctx = _native.Context(device)
        self.ctx.curve_set_tables(self.F, J, C)
        self.n_rates = J.shape[1]

    def value(self, flat, mask):
        self.ctx.portfolio_upload(flat)
        return self.ctx.portfolio_value_host(mask).copy()


def _interp(t, node_time, node_val, interp_type):
    a, b, wa, wb = plan_queries(t, node_time, interp_type)
    L = np.log(node_val)
    return np.exp(wa * L[a] + wb * L[b])


def yoy_arrays(swap, value_dt):
    """Leg arrays in the fixed leg's day count (engine.py:1082-1105)."""
    dc = swap._fixed_leg._dc_type
    t = lambda dts: np.array([times_from_dates(d, value_dt, dc) for d in dts], dtype=np.float64)  # noqa: E731
    fl, yl = swap._fixed_leg, swap._inflation_leg
    return dict(f_tp=t(fl._payment_dts), f_amt=_sign(fl) * fl._cpn * fl._notional * np.array(fl._year_fracs),
                f_principal=_sign(fl) * fl._principal,
                y_tp=t(yl._payment_dts), y_ts=t(yl._yoy_start_dts), y_te=t(yl._yoy_end_dts),
                y_scale=_sign(yl) * yl._notional * np.array(yl._year_fracs), y_spread=yl._spread)


def compute_yoy(derivatives, model, request_list, device=0) -> AnalyticsResult:
    """One or many YoY inflation swaps of one (currency, index): summed AnalyticsResult (Portfolio semantics)."""
    from .position import CurveSession
    reqs = set(request_list)
    want_cf = RequestTypes.CASHFLOWS in reqs
    if want_cf and len(derivatives) != 1:
        raise NotImplementedError("CASHFLOWS reports are per position (Position.compute), not per portfolio")
    d0 = derivatives[0]
    currency = d0._inflation_index._currency
    index_name = d0._inflation_index._index_type.name
    for d in derivatives[1:]:
        if d._inflation_index._currency != currency or d._inflation_index._index_type.name != index_name:
            raise LibError("YoY portfolio positions must share one currency and inflation index")
    if currency not in DISCOUNT_CURVE:
        raise LibError(f"No default OIS curve for currency {currency}")
    disc = getattr(model.curves, DISCOUNT_CURVE[currency], None) if _has(model, DISCOUNT_CURVE[currency]) else None
    if disc is None:
        raise LibError(f"Discount curve {DISCOUNT_CURVE[currency]} not found in model")
    key = (currency, index_name)
    if key not in INFLATION_CURVE:
        raise LibError(f"No inflation curve mapping for {currency.name} {index_name}. "
                       f"Add to model.curves as {currency.name}_{index_name}_INFLATION")
    infl = getattr(model.curves, INFLATION_CURVE[key], None) if _has(model, INFLATION_CURVE[key]) else None
    if infl is None:
        raise LibError(f"Inflation curve {INFLATION_CURVE[key]} not found in model")
    vd = model.value_dt
    want_v, want_d, want_g = (r in reqs for r in (RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA))
    mask = (_native.REQ_VALUE if want_v else 0) | (_native.REQ_DELTA if want_d else 0) | (_native.REQ_GAMMA if want_g else 0)
    cashflows = None
    if want_cf:          # engine.py:1355-1406: the non-AD valuation (its errors included) and the fixed leg's rows
        from .cashflows import yoy_cashflows
        cashflows = yoy_cashflows(d0, disc, infl, device)
    if mask == 0:
        return AnalyticsResult(cashflows=cashflows)
    dsess = CurveSession.get(disc, device)
    isess = InflationSession.get(infl, device)
    # ---- discount side: fixed cashflows on the OIS engine grid (inflation factors frozen)
    legs = [yoy_arrays(sw, vd) for sw in derivatives]
    dsess.ctx.portfolio_upload(discount_flat(legs, disc, isess.times, isess.F, infl._interp_type))
    agg_d = dsess.ctx.portfolio_value_host(mask | _native.REQ_VALUE).copy()
    value = Valuation(float(agg_d[0]), currency) if want_v else None
    delta = gamma = None
    if want_d or want_g:
        # ---- inflation side: coupons as product terms amt * I(e)/I(s) on the inflation nodes (DFs frozen)
        d_nodes, _, _ = dsess.ctx.curve_read(jac=False, hess=False)
        plan = disc.path_b_plan()
        agg_i = isess.value(inflation_flat(legs, plan.node_time, d_nodes, disc._interp_type, isess.times,
                                           infl._interp_type), mask | _native.REQ_VALUE)
        Rd, Ri = len(disc.swap_rates), isess.n_rates
        disc_type = {CurrencyTypes.GBP: CurveTypes.GBP_OIS_SONIA, CurrencyTypes.USD: CurveTypes.USD_OIS_SOFR,
                     CurrencyTypes.EUR: CurveTypes.EUR_OIS_ESTR}.get(currency, CurveTypes.GBP_OIS_SONIA)
        infl_type = CurveTypes[INFLATION_CURVE[key]]
        t_d, t_i = to_tenor(disc.swap_times), to_tenor(infl.swap_times)
        if want_d:
            delta = Risk([Delta(np.array(agg_d[1:1 + Rd]), t_d, currency, disc_type),
                          Delta(np.array(agg_i[1:1 + Ri]), t_i, currency, infl_type)])
        if want_g:
            gamma = Risk([Gamma(np.array(agg_d[33:].reshape(32, 32)[:Rd, :Rd]), t_d, currency, disc_type),
                          Gamma(np.array(agg_i[33:].reshape(32, 32)[:Ri, :Ri]), t_i, currency, infl_type)])
    return AnalyticsResult(value=value, risk=delta, gamma=gamma, cashflows=cashflows)


def _has(model, name) -> bool:
    return name in model._curves_dict


def discount_flat(legs, disc, infl_time, infl_F, infl_interp) -> FlatPortfolio:
    """One private unit per swap on the OIS engine grid: fixed coupons (t > 0, engine.py:2430) and the YoY coupons
    with their projected rates as plain amounts."""
    fl = Flattener(disc)
    for a in legs:
        live = a["f_tp"] > 0.0
        times = [((float(t), 1.0),) for t in a["f_tp"][live]]
        amts = [float(x) for x in a["f_amt"][live]]
        if a["f_principal"] != 0.0 and a["f_tp"].shape[0] and a["f_tp"][-1] > 0.0:
            times.append(((float(a["f_tp"][-1]), 1.0),))
            amts.append(float(a["f_principal"]))
        live = a["y_tp"] > 0.0
        ratio = _interp(a["y_te"], infl_time, infl_F, infl_interp) / _interp(a["y_ts"], infl_time, infl_F, infl_interp)
        pay = a["y_scale"] * (ratio - 1.0 + a["y_spread"])
        times += [((float(t), 1.0),) for t in a["y_tp"][live]]
        amts += [float(x) for x in pay[live]]
        if not times:                        # matured swap: a zero unit keeps the layout valid
            times, amts = [((0.0, 1.0),)], [0.0]
        fl.add_components([(("YD", id(a)), _Unit(times, amts), 1.0)])
    return fl.finalize(dedup=False)


def inflation_flat(legs, disc_time, disc_d, disc_interp, infl_time, infl_interp) -> FlatPortfolio:
    """One private unit per swap on the inflation node grid; term = sign N a_i DF(p_i) * I(e_i) / I(s_i), laid out as
    6 (node, weight) pairs: [end bracket (+), start bracket (-), unused]."""
    amt, ts, te, offsets = [], [], [], [0]
    for a in legs:
        live = a["y_tp"] > 0.0
        df = _interp(a["y_tp"][live], disc_time, disc_d, disc_interp) / _interp(np.zeros(1), disc_time, disc_d, disc_interp)[0]
        amt.append(a["y_scale"][live] * df)
        ts.append(a["y_ts"][live])
        te.append(a["y_te"][live])
        if not live.any():                    # zero term keeps the unit non-empty
            amt.append(np.zeros(1))
            ts.append(np.zeros(1))
            te.append(np.zeros(1))
        offsets.append(offsets[-1] + max(int(live.sum()), 1))
    amt, ts, te = np.concatenate(amt), np.concatenate(ts), np.concatenate(te)
    n = amt.shape[0]
    weight = np.zeros((n, 6))
    node = np.zeros((n, 6), dtype=np.int32)
    a_, b_, wa, wb = plan_queries(te, infl_time, infl_interp)
    weight[:, 0], weight[:, 1], node[:, 0], node[:, 1] = wa, wb, a_, b_
    a_, b_, wa, wb = plan_queries(ts, infl_time, infl_interp)
    weight[:, 2], weight[:, 3], node[:, 2], node[:, 3] = -wa, -wb, a_, b_
    node[weight == 0.0] = 0
    N = len(legs)
    return FlatPortfolio(N, n, np.array(offsets, dtype=np.int64), 6, amt, weight.reshape(-1), node.reshape(-1),
                         N, 1, np.ones(N), N, np.arange(N + 1, dtype=np.int64), np.arange(N, dtype=np.int32), None,
                         np.ones(N))

