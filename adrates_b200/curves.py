"""Discount curves and the host-side planning that precedes the CUDA kernels.

Two different node sets exist per curve, exactly as in the reference:

* path A - `OISCurve._times/_dfs`, bootstrapped by `OISCurve._build_curve_ad`
  (cavour/trades/rates/ois_curve.py:156-212): non-quoted coupon dates use log-linearly
  interpolated par rates.  Serves `curve.df`, `curve.df_ad`, the non-AD `leg.value`.
* path B - the dense engine grid of `Engine.build_curve_ad`
  (cavour/market/position/engine.py:2246-2360): every coupon date of every calibration
  swap is a node (duplicates kept) carrying its parent swap's rate.  VALUE/DELTA/GAMMA
  use it.  Here only the *plan* (node times, accruals, parent swap, previous-annuity
  node) is computed on the host; the recursion and its first/second-order tangents run
  on the GPU (csrc/cav_bootstrap.cu).

`plan_queries` turns cashflow times into (node, node, weight, weight) brackets with the
reference's interpolation rules (cavour/market/curves/interpolator_ad.py:186-249):
1e-10 grid snap to the first nearest node, +1e-12 bracket shift, searchsorted(side=
'right') duplicate semantics, end clamping.  All quirks live here so the kernels stay
generic (ln DF = wa*L[a] + wb*L[b]).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .dates import Date, DayCount, DayCountTypes, FrequencyTypes, annual_frequency, times_from_dates
from .argcheck import check_argument_types
from .error import LibError
from .global_types import InterpTypes

SNAP_TOL = 1e-10       # interpolator_ad.py:221
BRACKET_EPS = 1e-12    # interpolator_ad.py:224
TIME_FLOOR = 1e-15     # interpolator_ad.py:229


# ======================================================================================
# path B plan
# ======================================================================================
@dataclass
class PathBPlan:
    """Static (rate-independent) description of the engine's bootstrap grid."""
    node_time: np.ndarray   # [G] f64, sorted, duplicates kept; node_time[0] = 0
    node_acc: np.ndarray    # [G] f64 accrual of the coupon ending at the node
    node_swap: np.ndarray   # [G] i32 parent calibration swap (root uses swap 0)
    node_prev: np.ndarray   # [G] i32 node whose annuity precedes this coupon, -1 = none
    n_rates: int
    first_dup: np.ndarray   # [G] i32 first index sharing this node's time
    last_dup: np.ndarray    # [G] i32 last index sharing this node's time

    @property
    def n_nodes(self) -> int:
        return int(self.node_time.shape[0])


def plan_path_b(swap_times, year_fracs) -> PathBPlan:
    """Expand all coupon dates of all calibration swaps into the engine grid
    (engine.py:2283-2328): maturity = running sum of accruals, dependency keys are
    round(maturity, 2), stable sort by exact maturity, a key resolves to the FIRST node
    carrying it."""
    rows = [(0.0, 0.0, 0.0, None, 0)]  # (maturity, key, acc, prev_key, swap)
    for s, fracs in enumerate(year_fracs):
        run = 0.0
        for j, frac in enumerate(fracs):
            before = run
            run += frac
            rows.append((run, round(run, 2), frac, round(before, 2) if j > 0 else None, s))
    order = sorted(range(len(rows)), key=lambda i: rows[i][0])  # stable
    rows = [rows[i] for i in order]
    first_with_key = {}
    for idx, r in enumerate(rows):
        first_with_key.setdefault(r[1], idx)
    G = len(rows)
    t = np.array([r[0] for r in rows], dtype=np.float64)
    acc = np.array([r[2] for r in rows], dtype=np.float64)
    swap = np.array([r[4] for r in rows], dtype=np.int32)
    prev = np.array([-1 if r[3] is None else first_with_key.get(r[3], -1) for r in rows], dtype=np.int32)
    first = np.searchsorted(t, t, side="left").astype(np.int32)
    last = (np.searchsorted(t, t, side="right") - 1).astype(np.int32)
    if np.any(prev >= np.arange(G)):
        raise LibError("bootstrap plan is not causal (a coupon depends on a later node)")
    return PathBPlan(t, acc, swap, prev, len(year_fracs), first, last)


# ======================================================================================
# query planning: time -> (a, b, wa, wb) with ln DF(t) = wa*L[a] + wb*L[b]
# ======================================================================================
def plan_queries(t, node_time: np.ndarray, interp_type: InterpTypes):
    """Vectorised bracket planner.  Returns (a, b, wa, wb); a == b with wb == 0 for a
    grid snap or an end clamp."""
    if interp_type == InterpTypes.LINEAR_ZERO_RATES:
        lzr = True
    elif interp_type == InterpTypes.FLAT_FWD_RATES:
        lzr = False
    else:
        raise LibError("Invalid interpolation scheme.")  # interpolator_ad.py:236-237
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    x = node_time
    G = x.shape[0]
    # nearest node, first index on ties (jnp.argmin)
    r = np.searchsorted(x, t, side="left")
    lo = np.clip(r - 1, 0, G - 1)
    hi = np.clip(r, 0, G - 1)
    d_lo = np.abs(t - x[lo])
    d_hi = np.abs(t - x[hi])
    near = np.where(d_lo <= d_hi, lo, hi)
    near = np.searchsorted(x, x[near], side="left")      # first duplicate of that time
    snap = np.minimum(d_lo, d_hi) < SNAP_TOL
    # bracket on the shifted time (jnp.interp semantics)
    ts = t + BRACKET_EPS
    b = np.clip(np.searchsorted(x, ts, side="right"), 1, G - 1)
    a = b - 1
    dx = x[b] - x[a]
    flat = np.abs(dx) <= np.spacing(np.finfo(np.float64).eps)
    w = np.where(flat, 0.0, (ts - x[a]) / np.where(flat, 1.0, dx))
    above = ts > x[-1]
    below = ts < x[0]
    if lzr:
        xa = np.maximum(x[a], TIME_FLOOR)
        xb = np.maximum(x[b], TIME_FLOOR)
        wa = t * (1.0 - w) / xa
        wb = t * w / xb
        wa_hi = t / np.maximum(x[-1], TIME_FLOOR)
        wa_lo = t / np.maximum(x[0], TIME_FLOOR)
    else:
        wa = 1.0 - w
        wb = w
        wa_hi = np.ones_like(t)
        wa_lo = np.ones_like(t)
    a = np.where(above, G - 1, np.where(below, 0, a))
    b = np.where(above, G - 1, np.where(below, 0, b))
    wa = np.where(above, wa_hi, np.where(below, wa_lo, wa))
    wb = np.where(above | below, 0.0, wb)
    # snap overrides everything
    a = np.where(snap, near, a)
    b = np.where(snap, near, b)
    wa = np.where(snap, 1.0, wa)
    wb = np.where(snap, 0.0, wb)
    return a.astype(np.int32), b.astype(np.int32), wa, wb


# ======================================================================================
# curve objects (path A + API surface)
# ======================================================================================
_NODE_SCHEMES = (InterpTypes.LINEAR_ZERO_RATES, InterpTypes.FLAT_FWD_RATES, InterpTypes.LINEAR_FWD_RATES)


class DiscountCurve:
    """Base curve: `df(date, day_count)` and `df_ad(t)` on the path-A nodes
    (cavour/market/curves/discount_curve.py:300-436)."""

    _value_dt: Date
    _times: np.ndarray
    _dfs: np.ndarray
    _interp_type: InterpTypes

    def __init__(self, value_dt: Date, df_dts: list, df_values: np.ndarray, interp_type: InterpTypes = InterpTypes.FLAT_FWD_RATES):
        """Curve from year offsets and discount factors (discount_curve.py:40-93): offsets become dates with
        `add_years`, times are ACT/365 from the value date, (0, 1) is prepended unless the first date is the
        value date."""
        check_argument_types(self.__init__, locals())
        if len(df_dts) < 1:
            raise LibError("Times has zero length")
        if len(df_dts) != len(df_values):
            raise LibError("Times and Values are not the same")
        times, dfs = [0.0], [1.0]
        dts = value_dt.add_years(list(df_dts))
        start = 0
        if dts[0] == value_dt:
            dfs[0] = float(df_values[0])
            start = 1
        for i in range(start, len(dts)):
            times.append((dts[i] - value_dt) / 365.0)
            dfs.append(float(df_values[i]))
        self._times = np.array(times)
        if np.any(np.diff(self._times) <= 0.0):
            raise LibError("Times are not sorted in increasing order")
        self._value_dt = value_dt
        self._dfs = np.array(dfs)
        self._interp_type = interp_type
        self._df_dts = df_dts
        self._freq_type = FrequencyTypes.CONTINUOUS
        self._dc_type = DayCountTypes.ACT_ACT_ISDA
        from .interpolator import Interpolator
        self._interpolator = Interpolator(self._interp_type)
        self._interpolator.fit(self._times, self._dfs)

    def df(self, dt, day_count=DayCountTypes.ACT_ACT_ISDA):
        times = times_from_dates(dt, self._value_dt, day_count)
        if isinstance(times, np.ndarray):
            if np.any(times < 0.0):
                raise LibError("Interpolate times must all be >= 0")
            return np.array([self._node_df(float(u)) for u in times])
        if times < 0.0:
            raise LibError("Interpolate times must all be >= 0")
        if self._interp_type not in _NODE_SCHEMES and abs(times) >= 1e-12:
            return np.array([self._node_df(float(times))])     # the spline schemes answer a single date with a 1-vector
        return self._node_df(float(times))

    def _node_df(self, t: float) -> float:
        """Scalar look-up on the path-A nodes in the curve's own scheme (discount_curve.py:417-436 over
        interpolator.py:69-170 / the fitted splines): exact hit on node 0, first segment flat in zero rate for
        LINEAR_ZERO_RATES, true flat-forward extrapolation for FLAT_FWD_RATES."""
        from .interpolator import Interpolator, node_df
        if self._interp_type in _NODE_SCHEMES:
            return float(node_df(t, self._times, self._dfs, self._interp_type.value))
        fitted = getattr(self, "_interpolator", None)
        if fitted is None or fitted._times is not self._times or fitted._dfs is not self._dfs:   # fitted once per node set
            fitted = self._interpolator = Interpolator(self._interp_type)
            fitted.fit(self._times, self._dfs)
        return float(np.asarray(fitted.interpolate(float(t))).reshape(-1)[0])

    # -- rate views of the curve (discount_curve.py:96-296, 438-600): host arithmetic on df() --------------------------
    def value_dt(self) -> Date:
        return self._value_dt

    def _df(self, t):
        """DF at year offset(s) t in the curve's own scheme (discount_curve.py:417-436)."""
        if isinstance(t, np.ndarray):
            if np.any(t < 0.0):
                raise LibError("Interpolate times must all be >= 0")
            return np.array([self._node_df(float(u)) for u in t.ravel()])
        if t < 0.0:
            raise LibError("Interpolate times must all be >= 0")
        return self._node_df(float(t))

    def _zero_to_df(self, value_dt, rates, times, freq_type: FrequencyTypes, dc_type: DayCountTypes = None):
        t = np.maximum(np.array([times]) if isinstance(times, float) else times, 1e-12)
        if freq_type == FrequencyTypes.CONTINUOUS:
            return np.exp(-rates * t)
        if freq_type == FrequencyTypes.SIMPLE:
            return 1.0 / (1.0 + rates * t)
        if freq_type in (FrequencyTypes.ANNUAL, FrequencyTypes.SEMI_ANNUAL, FrequencyTypes.QUARTERLY, FrequencyTypes.MONTHLY):
            f = annual_frequency(freq_type)
            return 1.0 / np.power(1.0 + rates / f, f * t)
        raise LibError("Unknown Frequency type")

    def _df_to_zero(self, dfs, maturity_dts, freq_type: FrequencyTypes, dc_type: DayCountTypes) -> np.ndarray:
        dates = [maturity_dts] if isinstance(maturity_dts, Date) else maturity_dts
        values = [dfs] if isinstance(dfs, float) else dfs
        if len(dates) != len(values):
            raise LibError("Date list and df list do not have same length")
        f = annual_frequency(freq_type)
        times = times_from_dates(dates, self._value_dt, dc_type)
        out = []
        for df, t in zip(values, times):
            t = max(t, 1e-12)
            if freq_type == FrequencyTypes.CONTINUOUS:
                out.append(-np.log(df) / t)
            elif freq_type == FrequencyTypes.SIMPLE:
                out.append((1.0 / df - 1.0) / t)
            else:
                out.append((np.power(df, -1.0 / (t * f)) - 1.0) * f)
        return np.array(out)

    def zero_rate(self, dts, freq_type: FrequencyTypes = FrequencyTypes.CONTINUOUS, dc_type: DayCountTypes = DayCountTypes.ACT_360):
        """Zero rate(s) to the date(s) in the given compounding and day count (discount_curve.py:180-205)."""
        if not isinstance(freq_type, FrequencyTypes):
            raise LibError("Invalid Frequency type.")
        if not isinstance(dc_type, DayCountTypes):
            raise LibError("Invalid Day Count type.")
        zeros = self._df_to_zero(self.df(dts), dts, freq_type, dc_type)
        return zeros[0] if isinstance(dts, Date) else np.array(zeros)

    def cc_rate(self, dts, dc_type: DayCountTypes = DayCountTypes.SIMPLE):
        return self.zero_rate(dts, FrequencyTypes.CONTINUOUS, dc_type)

    def swap_rate(self, effective_dt: Date, maturity_dt, freq_type=FrequencyTypes.ANNUAL,
                  dc_type: DayCountTypes = DayCountTypes.THIRTY_E_360) -> np.ndarray:
        """Par rate(s) of fixed-for-floating swaps off this one curve, (DF(start) - DF(end)) / annuity on an unadjusted-calendar
        schedule whose first date is forced to the effective date; always an array (discount_curve.py:217-296)."""
        from .dates import Schedule
        if effective_dt < self._value_dt:
            raise LibError("Swap starts before the curve valuation date.")
        if not isinstance(freq_type, FrequencyTypes):
            raise LibError("Invalid Frequency type.")
        if freq_type == FrequencyTypes.SIMPLE:
            raise LibError("Cannot calculate par rate with simple yield freq.")
        if freq_type == FrequencyTypes.CONTINUOUS:
            raise LibError("Cannot calculate par rate with continuous freq.")
        dc = DayCount(dc_type)
        rates = []
        for mat in ([maturity_dt] if isinstance(maturity_dt, Date) else maturity_dt):
            if mat <= effective_dt:
                raise LibError("Maturity date is before the swap start date.")
            flow = list(Schedule(effective_dt, mat, freq_type).generate())
            flow[0] = effective_dt
            annuity, df = 0.0, 1.0
            for prev, nxt in zip(flow[:-1], flow[1:]):
                df = self.df(nxt)
                annuity += dc.year_frac(prev, nxt)[0] * df
            rates.append(0.0 if abs(annuity) < 1e-12 else (self.df(effective_dt) - df) / annuity)
        return np.array(rates)

    def survival_prob(self, dt: Date):
        return self.df(dt)

    def fwd(self, dts):
        """Continuously compounded one-day forward rate at the date(s) (discount_curve.py:446-468)."""
        nxt = [dts.add_days(1)] if isinstance(dts, Date) else [d.add_days(1) for d in dts]
        f = np.log(self.df(dts) / self.df(nxt)) / (1.0 / 365.0)
        return f[0] if isinstance(dts, Date) else np.array(f)

    def _fwd(self, times):
        h = 1e-6
        times = np.maximum(times, h)
        return np.log(self._df(times - h) / self._df(times + h)) / (2.0 * h)

    def bump(self, bump_size: float) -> "DiscountCurve":
        """A new curve with every zero rate moved up by bump_size (discount_curve.py:488-507).  As in the reference the node
        times are handed to the constructor as year offsets, so its `add_years` / 365-day round trip applies."""
        values = self._dfs * np.exp(-bump_size * self._times)
        return DiscountCurve(self._value_dt, self._times.tolist(), values, self._interp_type)

    def fwd_rate(self, start_dt, date_or_tenor, dc_type: DayCountTypes = DayCountTypes.ACT_360):
        """Simply compounded forward rate(s) from the start date(s) to a date / tenor (discount_curve.py:511-560)."""
        if isinstance(start_dt, Date):
            starts = [start_dt]
        elif isinstance(start_dt, list):
            starts = start_dt
        else:
            raise LibError("Start date and end date must be same types.")
        dc = DayCount(dc_type)
        out = []
        for i, d1 in enumerate(starts):
            d2 = d1.add_tenor(date_or_tenor) if isinstance(date_or_tenor, str) else \
                (date_or_tenor if isinstance(date_or_tenor, Date) else date_or_tenor[i])
            out.append((self.df(d1) / self.df(d2) - 1.0) / dc.year_frac(d1, d2)[0])
        return out[0] if isinstance(start_dt, Date) else np.array(out)

    def df_ad(self, dt, day_count=DayCountTypes.ACT_ACT_ISDA):
        """DF at time(s) `dt` in years (the reference's name for the argument): linear interpolation of piecewise forward rates
        on the path-A nodes, independent of `_interp_type` (discount_curve.py:317-415).  Evaluated by the CUDA library
        (cav_df_ad); there is no host fallback."""
        from . import _native
        tt = np.atleast_1d(np.asarray(dt, dtype=np.float64))
        out = _native.lib().df_ad(self._times, self._dfs, tt)
        return out[0] if np.ndim(dt) == 0 else out


class OISCurve(DiscountCurve):
    """Curve implied by OIS par rates (ois_curve.py:86-212).  Construction collects
    `swap_rates / swap_times / year_fracs` (the engine inputs) and bootstraps path A."""

    def __init__(self, value_dt: Date, ois_swaps: list, interp_type: InterpTypes = InterpTypes.FLAT_FWD_RATES,
                 check_refit: bool = False):
        check_argument_types(self.__init__, locals())
        if not isinstance(value_dt, Date):
            raise LibError("value_dt must be a Date")
        if not ois_swaps:
            raise LibError("OISCurve needs at least one calibration swap")
        self._value_dt = value_dt
        self._used_swaps = ois_swaps
        self._interp_type = interp_type
        self._check_refit = check_refit
        self._interpolator = None
        self._dc_type = ois_swaps[0]._float_leg._dc_type
        days = DayCount(self._dc_type).days_in_year()
        self.swap_rates = [s._fixed_coupon for s in ois_swaps]
        self.swap_times = [(s._adjusted_fixed_dts[-1] - value_dt) / days for s in ois_swaps]
        self.year_fracs = [list(s._fixed_leg._year_fracs) for s in ois_swaps]
        self._plan_b = None
        self._bootstrap_path_a()

    # -- path A -----------------------------------------------------------------------
    def _bootstrap_path_a(self):
        """d = (1 - r*annuity_prev)/(1 + r*acc); coupon dates that are not quoted
        maturities are filled in on demand with exp(interp(ln r)) par rates; annuities are
        memoised by round(t, 2) (ois_curve.py:156-212)."""
        times, dfs = self._path_a_scan([float(r) for r in self.swap_rates])
        self._times = np.array(times, dtype=np.float64)
        self._dfs = np.array(dfs, dtype=np.float64)

    def _path_a_scan(self, rates):
        """The path-A recursion on generic numbers (floats, or dual2.D2 for the node Jacobian)."""
        from .dual2 import exp, interp, log
        st = np.array(self.swap_times, dtype=np.float64)
        log_r = [log(r) for r in rates]
        annuity = {}
        times, dfs = [0.0], [1.0]

        def node(i, t_mat, rate, drop):
            fracs = self.year_fracs[i]
            if len(fracs) == 1:
                acc = fracs[0]
                d = 1.0 / (rate * acc + 1.0)
                a = d * acc
            else:
                acc = fracs[-1 - drop]
                t_prev = sum(fracs[:-1 - drop])
                key = round(t_prev, 2)
                if key not in annuity:
                    r_prev = exp(interp(t_prev, st, log_r))
                    annuity[key] = node(i, t_prev, r_prev, drop + 1)
                d = (1.0 - rate * annuity[key]) / (rate * acc + 1.0)
                a = annuity[key] + d * acc
            times.append(t_mat)
            dfs.append(d)
            annuity[round(t_mat, 2)] = a
            return a

        for i in range(len(self._used_swaps)):
            node(i, self.swap_times[i], rates[i], 0)
        return times, dfs

    def path_a_jacobian(self) -> np.ndarray:
        """d(path-A node DFs `_dfs`) / d(par rates), [nodes, R]: the map from this curve's quotes to the node DFs other
        curves are built on (XccyCurve reads the foreign curve through them), by first-order forward mode through the
        path-A recursion.  The reference has no such table - it chains `_mixed_hess_foreign_basis` (path-A nodes) with the
        engine-grid Jacobian and fails on the shape mismatch (engine.py:1936-1939)."""
        if getattr(self, "_jac_path_a", None) is None:
            from .dual2 import D2
            R = len(self.swap_rates)
            _, dfs = self._path_a_scan([D2.var(float(r), k, R) for k, r in enumerate(self.swap_rates)])
            self._jac_path_a = np.vstack([d.g if isinstance(d, D2) else np.zeros(R) for d in dfs])
        return self._jac_path_a

    # -- path B plan (cached; depends on dates only) --------------------------------------
    def path_b_plan(self) -> PathBPlan:
        if self._plan_b is None:
            self._plan_b = plan_path_b(self.swap_times, self.year_fracs)
        return self._plan_b

    def _check_refits(self, swap_tol: float = 1e-10):
        """Every calibration swap must reprice to ~0 on path A (ois_curve.py:344-358)."""
        for swap in self._used_swaps:
            v = swap.value(self._value_dt, self, self) / swap._fixed_leg._notional
            if abs(v) > swap_tol:
                raise LibError("Swap not repriced.")
