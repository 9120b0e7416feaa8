"""CASHFLOWS request of a single-curve OIS (reference engine.py:190-213, `_extract_leg_cashflows` :34-86), of bonds
(engine.py:648-696), of floating-rate notes (engine.py:930-983) and of year-on-year inflation swaps (engine.py:1355-1406), and
its result containers `CashflowItem` / `Cashflows` (results.py:945-1120).

The reference answers the request by valuing both legs on the NON-AD path - `SwapFixedLeg.value` / `SwapFloatLeg.value`
(swap_fixed_leg.py:200-245, swap_float_leg.py:190-352): one `DiscountCurve.df(date, day_count)` look-up on the path-A
nodes per payment / accrual date - and reading the per-payment lists the legs keep afterwards.  Here every date of both
legs becomes one batch of year fractions whose path-A discount factors come from ONE device call (`cav_curve_df`,
`k_curve_df`: Interpolator._uinterpolate per query); the rows are then assembled on the host.  Nothing is written into the
leg objects.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List

import numpy as np

from .dates import Date, DayCount, times_from_dates
from .error import LibError
from .global_types import SwapTypes
from .results import Valuation


@dataclass(frozen=True)
class CashflowItem:
    payment_date: Date
    notional: float
    payment_fraction: float        # amount / notional as the leg holds it (unsigned)
    accrual_period: float
    amount: float                  # signed: pay legs negative
    discount_factor: float
    discounted_amount: float
    leg_type: str                  # "Fixed_Pay" | "Fixed_Rec" | "Float_Pay" | "Float_Rec"

    def to_dict(self) -> Dict[str, Any]:
        return {"payment_date": str(self.payment_date), "notional": float(self.notional),
                "payment_fraction": float(self.payment_fraction), "accrual_period": float(self.accrual_period),
                "amount": float(self.amount), "discount_factor": float(self.discount_factor),
                "discounted_amount": float(self.discounted_amount), "leg_type": self.leg_type}


class Cashflows:
    def __init__(self, cashflows: List[CashflowItem], currency):
        self.cashflows = cashflows
        self.currency = currency

    def validate(self) -> bool:
        if not isinstance(self.cashflows, list):
            raise ValueError("cashflows must be a list")
        if not all(isinstance(cf, CashflowItem) for cf in self.cashflows):
            raise ValueError("All items must be CashflowItem instances")
        return True

    def to_dict(self) -> Dict[str, Any]:
        return {"currency": self.currency.name, "cashflows": [cf.to_dict() for cf in self.cashflows],
                "total_amount": float(self.total_amount), "total_pv": float(self.total_pv), "count": len(self.cashflows)}

    @property
    def df(self):
        import pandas as pd
        if not self.cashflows:
            return pd.DataFrame()
        frame = pd.DataFrame([cf.to_dict() for cf in self.cashflows])
        frame.set_index("payment_date", inplace=True)
        return frame

    @property
    def total_amount(self) -> float:
        return sum(cf.amount for cf in self.cashflows)

    @property
    def total_pv(self) -> float:
        return sum(cf.discounted_amount for cf in self.cashflows)

    def _only(self, word: str) -> "Cashflows":
        return Cashflows([cf for cf in self.cashflows if word in cf.leg_type], self.currency)

    def fixed(self): return self._only("Fixed")
    def floating(self): return self._only("Float")
    def pay(self): return self._only("Pay")
    def receive(self): return self._only("Rec")
    def notional_exchange(self): return self._only("Notional")

    def sum(self) -> Valuation:
        return Valuation(amount=self.total_pv, currency=self.currency)

    def __len__(self) -> int:
        return len(self.cashflows)

    def __repr__(self) -> str:
        return f"Cashflows(count={len(self.cashflows)}, total_pv={self.total_pv:,.2f} {self.currency.name})"


def _times(dts, value_dt: Date, dc_type) -> np.ndarray:
    return np.atleast_1d(np.asarray(times_from_dates(list(dts), value_dt, dc_type), dtype=np.float64))


def ois_cashflows(derivative, curve, device: int = 0) -> Cashflows:
    """Cashflow table of a single-curve OIS on `curve` (discounting and projection), rows in the reference's order: fixed
    leg then floating leg, payments in schedule order; payments on or before the value date carry zeros."""
    from .position import CurveSession
    fixed, flt = derivative._fixed_leg, derivative._float_leg
    if flt._notional_exchange or flt._notional_array:
        raise LibError("CASHFLOWS: notional exchanges / amortising notionals are outside the accelerated path")
    vd = curve._value_dt
    live_x = np.array([d > vd for d in fixed._payment_dts])
    live_f = np.array([d > vd for d in flt._payment_dts])
    # one batch of path-A queries: [value date (fixed dc), fixed pays, value date (float dc), float pays, starts, ends]
    t = np.concatenate([_times([vd], vd, fixed._dc_type), _times(fixed._payment_dts, vd, fixed._dc_type),
                        _times([vd], vd, flt._dc_type), _times(flt._payment_dts, vd, flt._dc_type),
                        _times(flt._start_accrued_dts, vd, flt._dc_type), _times(flt._end_accrued_dts, vd, flt._dc_type)])
    nx, nf = len(fixed._payment_dts), len(flt._payment_dts)
    live = np.concatenate([[True], live_x, [True], live_f, live_f, live_f])
    if np.any(t[live] < 0.0):
        raise LibError("Interpolate times must all be >= 0")
    q = np.where(live, t, 0.0)                      # dead rows are not looked up by the reference; keep the batch dense
    sess = CurveSession.get(curve, device)
    df = sess.ctx.curve_df(curve._interp_type.value, curve._times, curve._dfs, q)
    df0_x, df_x = df[0], df[1:1 + nx]
    df0_f, df_p, df_s, df_e = df[1 + nx], df[2 + nx:2 + nx + nf], df[2 + nx + nf:2 + nx + 2 * nf], df[2 + nx + 2 * nf:]

    items: List[CashflowItem] = []
    sign_x = -1.0 if fixed._leg_type == SwapTypes.PAY else 1.0
    name_x = "Fixed_Pay" if fixed._leg_type == SwapTypes.PAY else "Fixed_Rec"
    name_f = "Float_Rec" if fixed._leg_type == SwapTypes.PAY else "Float_Pay"   # the reference names it off the FIXED leg
    sign_f = -1.0 if "Pay" in name_f else 1.0
    notl = float(fixed._notional)
    for i in range(nx):
        amt = float(fixed._payments[i])
        dfp = float(df_x[i] / df0_x) if live_x[i] else 0.0
        pv = amt * dfp if live_x[i] else 0.0
        if i == nx - 1 and live_x[i]:
            pv += fixed._principal * dfp * fixed._notional
        items.append(CashflowItem(fixed._payment_dts[i], notl, amt / notl if notl != 0 else 0.0, float(fixed._year_fracs[i]),
                                  sign_x * amt, dfp, sign_x * pv, name_x))
    idx_dc = DayCount(curve._dc_type)
    notl = float(flt._notional)
    for i in range(nf):
        if live_f[i]:
            alpha_idx = idx_dc.year_frac(flt._start_accrued_dts[i], flt._end_accrued_dts[i])[0]
            fwd = (df_s[i] / df_e[i] - 1.0) / alpha_idx
            amt = float((fwd + flt._spread) * flt._year_fracs[i] * flt._notional)
            dfp = float(df_p[i] / df0_f)
            pv = amt * dfp
            if i == nf - 1:
                pv += flt._principal * dfp * flt._notional
        else:
            amt = dfp = pv = 0.0
        items.append(CashflowItem(flt._payment_dts[i], notl, amt / notl if notl != 0 else 0.0, float(flt._year_fracs[i]),
                                  sign_f * amt, dfp, sign_f * pv, name_f))
    return Cashflows(items, derivative._currency)


def _curve_dfs(curve, times, device: int) -> np.ndarray:
    """Path-A discount factors of `curve` at year fractions `times`, one device call (cav_curve_df)."""
    from .position import CurveSession
    t = np.asarray(times, dtype=np.float64)
    if np.any(t < 0.0):
        raise LibError("Interpolate times must all be >= 0")
    return CurveSession.get(curve, device).ctx.curve_df(curve._interp_type.value, curve._times, curve._dfs, t)


def bond_cashflows(bond, curve, device: int = 0) -> Cashflows:
    """CASHFLOWS of a bond (reference engine.py:648-696 over Bond.value, bond.py:264-365): a "Coupon" row per non-zero coupon and
    a "Principal" row per non-zero principal repayment, discount factors `curve.df(payment date)` in the curve API's default day
    count (ACT/ACT ISDA), relative to the value date; payments on or before it carry zero DF / PV."""
    from .dates import DayCountTypes
    vd = curve._value_dt
    dts = bond._payment_dts
    live = np.array([d > vd for d in dts])
    t = np.concatenate([_times([vd], vd, DayCountTypes.ACT_ACT_ISDA), _times(dts, vd, DayCountTypes.ACT_ACT_ISDA)])
    df = _curve_dfs(curve, np.where(np.concatenate([[True], live]), t, 0.0), device)
    df0, dfp = df[0], df[1:]
    items: List[CashflowItem] = []
    for i, dt in enumerate(dts):
        rel = float(dfp[i] / df0) if live[i] else 0.0
        cpn = float(bond._coupon_payments[i])
        prin = float(bond._principal_payments[i])
        if abs(cpn) > 1e-10:
            notl = bond._principal_schedule[i]
            items.append(CashflowItem(dt, notl, cpn / notl if notl != 0 else 0.0, float(bond._year_fracs[i]), cpn, rel,
                                      cpn * rel if live[i] else 0.0, "Coupon"))
        if abs(prin) > 1e-10:
            pv = prin * rel if (live[i] and prin > 0) else 0.0
            items.append(CashflowItem(dt, prin, 1.0, 0.0, prin, rel, pv, "Principal"))
    return Cashflows(items, bond._currency)


def frn_cashflows(frn, discount_curve, index_curve, device: int = 0) -> Cashflows:
    """CASHFLOWS of a floating-rate note (reference engine.py:930-983 over FRN.value, frn.py:218-330): a "Floating_Coupon" row per
    live coupon - forward rate off the index curve (or the first fixing), plus margin, capped / floored - and a "Principal" row
    at the last payment date; all year fractions in the note's day count, forward accruals in the index curve's."""
    vd = discount_curve._value_dt
    dts = frn._payment_dts
    n = len(dts)
    live = np.array([d > vd for d in dts])
    # which live periods need index-curve look-ups (all but a first live period that carries a fixing)
    first_live = int(np.argmax(live)) if live.any() else -1
    fixed_first = frn._first_fixing_rate is not None and first_live >= 0
    need = live.copy()
    if fixed_first:
        need[first_live] = False
    ts = _times(frn._start_accrued_dts, vd, frn._dc_type)
    te = _times(frn._end_accrued_dts, vd, frn._dc_type)
    idx = _curve_dfs(index_curve, np.concatenate([np.where(need, ts, 0.0), np.where(need, te, 0.0)]), device)
    tp = np.concatenate([_times([vd], vd, frn._dc_type), _times(dts, vd, frn._dc_type)])
    dis = _curve_dfs(discount_curve, np.where(np.concatenate([[True], live]), tp, 0.0), device)
    df0, dfp = dis[0], dis[1:]
    idx_dc = DayCount(index_curve._dc_type)
    items: List[CashflowItem] = []
    face = float(frn._face_value)
    for i, dt in enumerate(dts):
        rate = amount = rel = 0.0
        if live[i]:
            if fixed_first and i == first_live:
                fwd = frn._first_fixing_rate
            else:
                fwd = (idx[i] / idx[n + i] - 1.0) / idx_dc.year_frac(frn._start_accrued_dts[i], frn._end_accrued_dts[i])[0]
            rate = fwd + frn._quoted_margin
            if frn._cap_rate is not None:
                rate = min(rate, frn._cap_rate)
            if frn._floor_rate is not None:
                rate = max(rate, frn._floor_rate)
            amount = float(rate * frn._year_fracs[i] * frn._face_value)
            rel = float(dfp[i] / df0)
        if abs(amount) > 1e-10:
            items.append(CashflowItem(dt, face, float(rate), float(frn._year_fracs[i]), amount, rel, amount * rel, "Floating_Coupon"))
        if i == n - 1:
            items.append(CashflowItem(dt, face, 1.0, 0.0, face, rel, face * rel, "Principal"))
    return Cashflows(items, frn._currency)


def yoy_cashflows(swap, discount_curve, inflation_curve, device: int = 0) -> Cashflows:
    """CASHFLOWS of a year-on-year inflation swap as the reference answers it (engine.py:1355-1406): the non-AD
    `YoYInflationSwap.value` first (yoy_inflation_swap.py:224-260: fixed leg on the path-A nodes, then the inflation leg, whose
    CPI look-ups raise for reference dates before the value date without a fixing - every sub-annual or seasoned swap,
    swap_yoy_inflation_leg.py:267-366), then the rows of the FIXED leg only: the reference reads the inflation leg's rows off
    `_payment_pvs`, an attribute that leg never has (it keeps `_pvs`), so its loop at engine.py:1371-1402 never runs.  The
    rows pinned by the unmodified reference are tests/golden/ref_cashflows_yoy.json."""
    fixed, leg = swap._fixed_leg, swap._inflation_leg
    vd = discount_curve._value_dt
    dts = fixed._payment_dts
    live = np.array([d > vd for d in dts])
    t = np.concatenate([_times([vd], vd, fixed._dc_type), _times(dts, vd, fixed._dc_type)])
    df = _curve_dfs(discount_curve, np.where(np.concatenate([[True], live]), t, 0.0), device)
    # the inflation leg's non-AD valuation: nothing of it reaches the rows, its errors (and the lists it keeps on the leg) do
    leg.value(vd, discount_curve, inflation_curve)
    index = leg._inflation_index
    sign = -1.0 if fixed._leg_type == SwapTypes.PAY else 1.0
    name = "Fixed_Pay" if fixed._leg_type == SwapTypes.PAY else "Fixed_Rec"
    notl = float(fixed._notional)
    items: List[CashflowItem] = []
    for i, dt in enumerate(dts):
        amt = float(fixed._payments[i])
        rel = float(df[1 + i] / df[0]) if live[i] else 0.0
        pv = amt * rel if live[i] else 0.0
        if i == len(dts) - 1 and live[i]:
            pv += fixed._principal * rel * fixed._notional
        items.append(CashflowItem(dt, notl, amt / notl if notl != 0 else 0.0, float(fixed._year_fracs[i]), sign * amt, rel,
                                  sign * pv, name))
    return Cashflows(items, index._currency)
