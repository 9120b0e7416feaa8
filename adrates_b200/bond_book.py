"""Array-based books of fixed-coupon bullet bonds on the CUDA valuation path.

`Engine._compute_bond` (cavour/market/position/engine.py:505-640) prices a bond with the engine's fixed-leg pricer on
the OIS curve of its currency: coupons `year_frac x coupon x face` on the payment dates plus the face value on the last
payment date, investor side, times in the bond's day count, flows on or before the value date worth nothing
(bond.py:162-245 for the schedule).  Per bond that is

    PV = face * coupon * A_s + face * R_s,   A_s = sum_{t_i > 0} alpha_i DF(t_i),   R_s = DF(t_last),

where s is the bond's schedule class (issue date, maturity date): bonds of one class share the annuity unit A_s and the
redemption unit R_s, exactly like the OIS of one schedule class share their leg units (batch.OISBook).  The schedules,
day counts and payment lags are rolled as arrays by the same rules as the object layer (`credit.Bond`), and the flat
book runs through the same kernels (VALUE / DELTA / GAMMA, scenarios).  Amortising and zero-coupon bonds stay on the
object route (`Portfolio([...]).compute`).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .batch import I64, OISBook, add_tenor, adjust, leg_schedules, serials, year_frac
from .curves import OISCurve
from .dates import BusDayAdjustTypes, CalendarTypes, DateGenRuleTypes, DayCountTypes, FrequencyTypes
from .error import LibError
from .flatten import FlatPortfolio


@dataclass
class BondBook:
    """Fixed-coupon bullet bonds on one OIS curve, one array entry per bond; conventions are per book."""
    curve: OISCurve
    issue: np.ndarray                # int64 serials
    maturity: np.ndarray             # int64 serials (unadjusted)
    coupon: np.ndarray
    face: np.ndarray
    freq_type: FrequencyTypes = FrequencyTypes.SEMI_ANNUAL
    dc_type: DayCountTypes = DayCountTypes.ACT_365F
    payment_lag: int = 0
    cal_type: CalendarTypes = CalendarTypes.WEEKEND
    bd_type: BusDayAdjustTypes = BusDayAdjustTypes.FOLLOWING
    dg_type: DateGenRuleTypes = DateGenRuleTypes.BACKWARD
    end_of_month: bool = False

    @property
    def n_trades(self) -> int:
        return int(self.issue.shape[0])

    @classmethod
    def from_arrays(cls, curve: OISCurve, issue, maturity=None, tenor_years=None, tenor_months=None, coupon=None,
                    face_value=100.0, **conventions) -> "BondBook":
        """issue / maturity: list[Date] or serials; or `tenor_years` / `tenor_months` like Date.add_tenor."""
        iss = serials(issue)
        n = iss.shape[0]
        if maturity is not None:
            mat = serials(maturity)
        elif tenor_years is not None:
            mat = add_tenor(iss, tenor_years, "Y")
        elif tenor_months is not None:
            mat = add_tenor(iss, tenor_months, "M")
        else:
            raise LibError("BondBook needs maturity dates or tenors")
        if mat.shape[0] != n:
            raise LibError("issue and maturity arrays differ in length")
        if np.any(iss >= mat):
            raise LibError("Issue date must be before maturity date")
        if coupon is None:
            raise LibError("coupon is required")
        vec = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))  # noqa: E731
        book = cls(curve, iss, mat, vec(coupon), vec(face_value), **conventions)
        if book.freq_type == FrequencyTypes.ZERO or np.any(book.coupon == 0.0):
            raise LibError("zero-coupon bonds are valued on the object route (credit.Bond)")
        return book

    def schedule_classes(self):
        span = self.maturity - self.issue
        if np.any(span <= 0) or np.any(span >= (1 << 22)):
            raise LibError("Issue date must be before maturity date")
        uniq, cls_of = np.unique((self.issue << 22) | span, return_inverse=True)
        iss = uniq >> 22
        return iss, iss + (uniq & ((1 << 22) - 1)), cls_of.reshape(-1).astype(I64)

    def flatten(self, dedup: bool = True, max_group: int = 256, tiles: bool = True) -> FlatPortfolio:
        """dedup=True: one annuity and one redemption unit per schedule class, bonds carry the weights
        (face x coupon, face).  dedup=False: one private unit per bond."""
        curve = self.curve
        vd1 = np.array([curve._value_dt._n], dtype=I64)
        iss, mat, cls_of = self.schedule_classes()
        S = iss.shape[0]
        leg = leg_schedules(iss, mat, self.freq_type, self.dc_type, self.payment_lag, self.cal_type, self.bd_type,
                            self.dg_type, self.end_of_month)
        # bond.py:221-236: the first accrual starts on the issue date itself, whatever the schedule's first date is
        # (the FORWARD rule adjusts it, the BACKWARD rule does not)
        alpha = leg.alpha.copy()
        first = leg.offsets[:-1]
        alpha[first] = year_frac(iss, leg.end[first], self.dc_type)
        owner = np.repeat(np.arange(S, dtype=I64), np.diff(leg.offsets))
        t = year_frac(np.broadcast_to(vd1, leg.pay.shape), leg.pay, self.dc_type)
        live = t > 0.0                                   # engine.py:2430: strictly after the value date
        A = (owner[live], t[live], alpha[live])
        last = leg.offsets[1:] - 1                       # the face value travels with the last payment
        r_live = t[last] > 0.0
        R = (np.arange(S, dtype=I64)[r_live], t[last][r_live], np.ones(int(r_live.sum())))
        wA = self.face * self.coupon
        wR = self.face
        if dedup:
            flat = self._flatten_shared(S, cls_of, A, R, None, wA, wR, None, max_group)
        else:
            flat = self._flatten_private(S, cls_of, A, R, None, wA, wR, None)
        if tiles:
            plan = curve.path_b_plan()
            flat.with_tiles(plan.n_nodes, plan)
        return flat

    # the flat-book builders and the device calls are those of the OIS books (they only use `curve`, `n_trades`, `flatten`)
    _plan = OISBook._plan
    _flatten_shared = OISBook._flatten_shared
    _flatten_private = OISBook._flatten_private
    def upload(self, ctx, tiles: bool = True, dedup: bool = True, device_flatten: bool = True) -> str:
        """Bond books are flattened on the host (the device flattener covers vanilla OIS)."""
        ctx.portfolio_upload(self.flatten(dedup=dedup, tiles=tiles))
        return "host"

    _value = OISBook._value
    compute = OISBook.compute
    scenario_values = OISBook.scenario_values
