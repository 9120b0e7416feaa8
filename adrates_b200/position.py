"""Position / Engine / Portfolio: the reference-facing entry points of the CUDA path.

    Position(derivative, model).compute(request_list, collateral_type=None) -> AnalyticsResult
        cavour/market/position/position.py:25-80, Engine.compute engine.py:89-124,
        _compute_ois_natural engine.py:153-215
    Portfolio(positions).compute(request_list) -> AnalyticsResult (sums)
        cavour/market/portfolio/portfolio.py:8-67

Unlike the reference, where every Position owns a fresh Engine and re-derives the curve
Jacobian/Hessian (position.py:55, engine.py:2362-2412), curve tables are built once per
(curve quotes, plan) on the device and shared; Portfolio.compute flattens all positions and
values them in one batched launch instead of a Python loop.
"""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

from . import _native
from .curves import OISCurve
from .dates import to_tenor
from .error import LibError
from .flatten import Flattener
from .global_types import InstrumentTypes, RequestTypes
from .results import AnalyticsResult, Delta, Gamma, Valuation


TILE_MIN_UNITS = 64     # below this the warp-per-unit kernel is as fast and the tile plan is pure overhead


def request_mask(request_list, allow_cashflows: bool = False) -> int:
    reqs = set(request_list)
    unknown = [r for r in reqs if not isinstance(r, RequestTypes)]
    if unknown:
        raise LibError(f"Unknown request types: {unknown}")
    mask = 0
    if RequestTypes.VALUE in reqs:
        mask |= _native.REQ_VALUE
    if RequestTypes.DELTA in reqs:
        mask |= _native.REQ_DELTA
    if RequestTypes.GAMMA in reqs:
        mask |= _native.REQ_GAMMA
    if RequestTypes.CASHFLOWS in reqs and not allow_cashflows:
        raise NotImplementedError("CASHFLOWS reports are offered for single-curve OIS positions (Position.compute)")
    return mask


class CurveSession:
    """Device-resident tables of one curve (replaces Engine._cached_curve)."""

    _cache = {}

    @classmethod
    def get(cls, curve: OISCurve, device: int = 0) -> "CurveSession":
        key = (device, tuple(curve.swap_rates), tuple(curve.swap_times), tuple(map(tuple, curve.year_fracs)),
               curve._interp_type)
        sess = cls._cache.pop(key, None)
        if sess is None:
            if len(cls._cache) >= 16:
                # least recently used goes; its context is NOT closed here - a caller may still hold the session (the
                # XCCY / YoY engines keep two at a time); Context.__del__ frees the device memory with the last reference
                cls._cache.pop(next(iter(cls._cache)))
            sess = CurveSession(curve, device)
        cls._cache[key] = sess          # (re)inserted at the end: dict order is the LRU order
        return sess

    def __init__(self, curve: OISCurve, device: int):
        self.ctx = _native.Context(device)
        self.curve = curve
        self.ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)


class Engine:
    def __init__(self, model, device: int = 0):
        self.model = model
        self.device = device

    def _curve_for(self, derivative) -> OISCurve:
        try:
            return getattr(self.model.curves, derivative._floating_index.name)
        except AttributeError:
            raise LibError(f"No curve {derivative._floating_index.name} in the model")

    def compute(self, derivative, request_list, collateral_type=None) -> AnalyticsResult:
        dtype = getattr(derivative, "derivative_type", None)
        if dtype == InstrumentTypes.XCCY_SWAP:
            from .xccy_engine import compute_xccy
            return compute_xccy([derivative], self.model, request_list, self.device)
        if dtype == InstrumentTypes.YOY_INFLATION_SWAP:      # Engine._compute_yoy_iis (engine.py:120-122, 986-1408)
            from .yoy_engine import compute_yoy
            return compute_yoy([derivative], self.model, request_list, self.device)
        want_cf = RequestTypes.CASHFLOWS in set(request_list) and dtype in (InstrumentTypes.BOND, InstrumentTypes.FRN)
        rest = [r for r in request_list if r != RequestTypes.CASHFLOWS] if want_cf else request_list
        if want_cf:
            request_mask(rest)                  # unknown request types are rejected as without CASHFLOWS

        def with_cashflows(res, cf):
            return AnalyticsResult(value=res.value, risk=res.risk, gamma=res.gamma, cashflows=cf)

        if dtype == InstrumentTypes.BOND:      # Engine._compute_bond (engine.py:505-640): OIS curve of the currency
            curve = self._curve_for(derivative)
            res = value_positions([derivative], curve, rest, self.device) if (rest or not want_cf) else AnalyticsResult()
            if not want_cf:
                return res
            from .cashflows import bond_cashflows
            return with_cashflows(res, bond_cashflows(derivative, curve, self.device))
        if dtype == InstrumentTypes.FRN:       # Engine._compute_frn (engine.py:700-925)
            from .credit import BOND_CURVE
            if derivative._currency not in BOND_CURVE:
                raise LibError(f"No default OIS curve for currency {derivative._currency}")
            index_curve = self._curve_for(derivative)
            try:
                discount_curve = getattr(self.model.curves, BOND_CURVE[derivative._currency].name)
            except AttributeError:
                raise LibError(f"No curve {BOND_CURVE[derivative._currency].name} in the model")
            if BOND_CURVE[derivative._currency] != derivative._floating_index:
                # the reference values dual-curve FRNs but has no Greeks for them (engine.py:921-924)
                if set(rest) - {RequestTypes.VALUE}:
                    raise LibError("Dual-curve FRN delta/gamma not yet implemented. "
                                   "Use same curve for discounting and projection.")
                from .xccy_engine import value_frn_dual_curve
                res = value_frn_dual_curve(derivative, self.model, self.device) if (rest or not want_cf) else AnalyticsResult()
            else:
                res = value_positions([derivative], index_curve, rest, self.device) if (rest or not want_cf) else AnalyticsResult()
            if not want_cf:
                return res
            from .cashflows import frn_cashflows
            return with_cashflows(res, frn_cashflows(derivative, discount_curve, index_curve, self.device))
        if dtype != InstrumentTypes.OIS_SWAP:
            raise LibError(f"{dtype} not yet implemented")
        if collateral_type is not None:
            from .global_types import collateral_to_currency
            collateral_ccy = collateral_to_currency(collateral_type)
            if collateral_ccy != derivative._currency:
                from .xccy_engine import compute_ois_xccy_collateral
                return compute_ois_xccy_collateral(derivative, self.model, request_list, collateral_ccy, self.device)
        curve = self._curve_for(derivative)
        if RequestTypes.CASHFLOWS not in set(request_list):
            return value_positions([derivative], curve, request_list, self.device)
        # CASHFLOWS (engine.py:190-213): the non-AD leg valuation on the path-A nodes, one batched DF call on the device
        from .cashflows import ois_cashflows
        rest = [r for r in request_list if r != RequestTypes.CASHFLOWS]
        request_mask(rest)                                    # unknown request types are rejected as without CASHFLOWS
        res = value_positions([derivative], curve, rest, self.device) if rest else AnalyticsResult()
        return AnalyticsResult(value=res.value, risk=res.risk, gamma=res.gamma,
                               cashflows=ois_cashflows(derivative, curve, self.device))


ARRAY_ROUTE_MIN = 512   # object books at least this large are flattened as arrays (batch.OISBook)


def _ois_conventions(d):
    """Convention tuple of a vanilla OIS whose legs batch.OISBook reproduces exactly, else None."""
    if getattr(d, "derivative_type", None) != InstrumentTypes.OIS_SWAP:
        return None
    fl, ft = d._fixed_leg, d._float_leg
    same = ("_effective_dt", "_termination_dt", "_payment_lag", "_cal_type", "_bd_type", "_dg_type", "_notional")
    if any(getattr(fl, a) != getattr(ft, a) for a in same) or fl._end_of_month or ft._end_of_month:
        return None
    if fl._principal != 0.0 or ft._principal != 0.0 or ft._notional_array or ft._notional_exchange:
        return None
    if fl._leg_type == ft._leg_type:
        return None
    return (fl._freq_type, fl._dc_type, ft._freq_type, ft._dc_type, fl._payment_lag, fl._cal_type, fl._bd_type, fl._dg_type)


def _value_as_arrays(derivatives, curve: OISCurve, mask: int, sess) -> np.ndarray | None:
    """Large books of vanilla OIS: the legs' dates are re-rolled as arrays (batch.py, identical rules) instead of
    walking ~100 Python objects per trade; one valuation per convention set, totals added.  None = not applicable."""
    from .batch import OISBook
    from .global_types import SwapTypes
    groups = {}
    for d in derivatives:
        conv = _ois_conventions(d)
        if conv is None:
            return None
        groups.setdefault(conv, []).append(d)
    total = np.zeros(_native.NOUT)
    for conv, ds in groups.items():
        n = len(ds)
        book = OISBook(curve,
                       np.fromiter((d._fixed_leg._effective_dt._n for d in ds), dtype=np.int64, count=n),
                       np.fromiter((d._fixed_leg._termination_dt._n for d in ds), dtype=np.int64, count=n),
                       np.fromiter((1.0 if d._fixed_leg._leg_type == SwapTypes.RECEIVE else -1.0 for d in ds), dtype=np.float64, count=n),
                       np.fromiter((d._fixed_leg._cpn for d in ds), dtype=np.float64, count=n),
                       np.fromiter((d._fixed_leg._notional for d in ds), dtype=np.float64, count=n),
                       np.fromiter((d._float_leg._spread for d in ds), dtype=np.float64, count=n),
                       conv[0], conv[1], conv[2], conv[3], conv[4], conv[5], conv[6], conv[7])
        flat = book.flatten(dedup=True, tiles=bool(mask & _native.REQ_GAMMA))
        sess.ctx.portfolio_upload(flat)
        total += sess.ctx.portfolio_value_host(mask)
    return total


def value_positions(derivatives, curve: OISCurve, request_list, device: int = 0, dedup=None) -> AnalyticsResult:
    """Flatten -> upload -> one batched valuation; returns the summed AnalyticsResult."""
    mask = request_mask(request_list)
    sess = CurveSession.get(curve, device)
    if dedup is None and len(derivatives) >= ARRAY_ROUTE_MIN:
        agg = _value_as_arrays(derivatives, curve, mask, sess)
        if agg is not None:
            return _result_from_totals(agg, mask, curve, derivatives[0])
    fl = Flattener(curve)
    for d in derivatives:
        fl.add_trade(d)
    flat = fl.finalize(dedup=(len(derivatives) > 1) if dedup is None else dedup)
    if (mask & _native.REQ_GAMMA) and flat.n_units >= TILE_MIN_UNITS:
        plan = curve.path_b_plan()          # books of any size get the tensor-core Greeks kernel (tiles.py)
        flat.with_tiles(plan.n_nodes, plan)
    sess.ctx.portfolio_upload(flat)
    agg = sess.ctx.portfolio_value_host(mask)
    return _result_from_totals(agg, mask, curve, derivatives[0])


def _result_from_totals(agg, mask, curve: OISCurve, derivative) -> AnalyticsResult:
    R = len(curve.swap_rates)
    tenors = to_tenor(curve.swap_times)
    ccy, idx = derivative._currency, derivative._floating_index
    value = Valuation(float(agg[0]), ccy) if mask & _native.REQ_VALUE else None
    delta = Delta(np.array(agg[1:1 + R]), tenors, ccy, idx) if mask & _native.REQ_DELTA else None
    gamma = Gamma(np.array(agg[33:].reshape(32, 32)[:R, :R]), tenors, ccy, idx) if mask & _native.REQ_GAMMA else None
    return AnalyticsResult(value=value, risk=delta, gamma=gamma)


class Position:
    def __init__(self, derivative, model):
        self.derivative = derivative
        self.model = model
        self._engine = Engine(model)

    def compute(self, request_list, collateral_type=None) -> AnalyticsResult:
        return self._engine.compute(self.derivative, request_list, collateral_type)


class Portfolio:
    """Aggregates positions; compute() is ONE batched device valuation per curve."""

    def __init__(self, positions: Iterable[Position] | None = None) -> None:
        self._positions: List[Position] = list(positions or [])

    def add_position(self, position: Position) -> None:
        self._positions.append(position)

    def positions(self) -> List[Position]:
        return list(self._positions)

    def compute(self, request_list: Iterable[RequestTypes]) -> AnalyticsResult:
        request_list = list(request_list)
        if not self._positions:
            return AnalyticsResult()
        kinds = {getattr(p.derivative, "derivative_type", None) for p in self._positions}
        if kinds == {InstrumentTypes.XCCY_SWAP}:
            from .xccy_engine import compute_xccy
            models = {id(p.model) for p in self._positions}
            if len(models) != 1:
                raise LibError("XCCY portfolio positions must share one Model")
            return compute_xccy([p.derivative for p in self._positions], self._positions[0].model, request_list)
        if kinds == {InstrumentTypes.YOY_INFLATION_SWAP}:
            from .yoy_engine import compute_yoy
            if len({id(p.model) for p in self._positions}) != 1:
                raise LibError("YoY inflation swap positions must share one Model")
            return compute_yoy([p.derivative for p in self._positions], self._positions[0].model, request_list)
        if InstrumentTypes.YOY_INFLATION_SWAP in kinds:
            raise LibError("mixed OIS / YoY inflation portfolios are not supported (the reference cannot add Risk to Delta)")
        if InstrumentTypes.XCCY_SWAP in kinds:
            raise LibError("mixed OIS / XCCY portfolios are not supported (the reference cannot add Risk to Delta)")
        buckets = {}
        singles = []          # positions the batched single-curve valuation must not take: valued like Position.compute
        for pos in self._positions:
            d = pos.derivative
            if getattr(d, "derivative_type", None) == InstrumentTypes.FRN:
                # Engine._compute_frn (engine.py:700-925): the note is discounted on the OIS curve of its currency; when
                # its index curve is another one it is a dual-curve valuation (VALUE only, Greeks raise), never the
                # single-curve unit the flattener would build on the index curve
                from .credit import BOND_CURVE
                if d._currency not in BOND_CURVE:
                    raise LibError(f"No default OIS curve for currency {d._currency}")
                if BOND_CURVE[d._currency] != d._floating_index:
                    singles.append(pos)
                    continue
            curve = pos._engine._curve_for(d)
            buckets.setdefault(id(curve), (curve, []))[1].append(d)
        total = None
        results = [value_positions(derivs, curve, request_list) for curve, derivs in buckets.values()]
        results += [pos.compute(request_list) for pos in singles]
        for res in results:
            if total is None:
                total = res
            else:   # same semantics as the reference's running sums (portfolio.py:48-65)
                total = AnalyticsResult(
                    value=None if res.value is None else total.value + res.value,
                    risk=None if res.risk is None else total.risk + res.risk,
                    gamma=None if res.gamma is None else total.gamma + res.gamma)
        return total

    def scenario_values(self, curve_name: str, shocks, pnl: bool = False, device: int = 0):
        """Values of every position under a batch of shocks of `curve_name` in one device pass: what the loop
        `[p.derivative.position(model.scenario(curve_name, s)).compute([VALUE]).value.amount for s in shocks for p in
        positions]` returns, as a torch CUDA tensor [S, n_positions] (pnl=True: minus the unshocked values).  `shocks` as
        in `Model.scenario_rates`.  All positions must share one Model and be valued on `curve_name` (OIS, bonds and
        single-curve FRNs; the reference has no batched counterpart, models.py:507-557 + position.py:62-80)."""
        from .scenarios import scenario_values_flat
        import torch
        if not self._positions:
            return torch.empty(len(shocks), 0, dtype=torch.float64, device=torch.device("cuda", device))
        if len({id(p.model) for p in self._positions}) != 1:
            raise LibError("scenario_values: positions must share one Model")
        model = self._positions[0].model
        curve = model.curves[curve_name]
        from .credit import BOND_CURVE
        for p in self._positions:
            d = p.derivative
            kind = getattr(d, "derivative_type", None)
            ok = kind in (InstrumentTypes.OIS_SWAP, InstrumentTypes.BOND) or (
                kind == InstrumentTypes.FRN and BOND_CURVE.get(d._currency) == d._floating_index)
            if not ok or p._engine._curve_for(d) is not curve:
                raise LibError(f"scenario_values: every position must be an OIS / bond / single-curve FRN valued on {curve_name}")
        fl = Flattener(curve)
        for p in self._positions:
            fl.add_trade(p.derivative)
        flat = fl.finalize(dedup=len(self._positions) > 1)
        return scenario_values_flat(curve, flat, model.scenario_rates(curve_name, shocks), device, pnl)
