"""Host-side discount-factor interpolation on curve nodes: the reference's `interpolate` / `_uinterpolate` /
`Interpolator` (cavour/market/curves/interpolator.py:35-170, 197-560).

This is the NON-AD look-up behind `DiscountCurve.df` and the path-A bootstrap - host code in the reference too, a few nodes per
call.  Batched queries of the valuation path do not come through here: they run on the device (`cav_curve_df` / `k_curve_df` for
path-A nodes, the planner + `exp` of the tile kernels for the engine grid).

Three schemes work on the nodes directly (FLAT_FWD_RATES, LINEAR_FWD_RATES, LINEAR_ZERO_RATES); the spline schemes fit a SciPy
interpolant of zero rates or log discount factors once (`Interpolator.fit`).
"""
from __future__ import annotations

import math

import numpy as np

from .error import LibError
from .global_types import InterpTypes

G_SMALL = 1e-12           # the reference's g_small (utils/global_vars.py:4)
_FWD_SMALL = 1e-10        # regulariser of the first LINEAR_FWD_RATES segment (interpolator.py:75, 147-149)


def node_df(t: float, times: np.ndarray, dfs: np.ndarray, method: int) -> float:
    """One discount factor from node arrays (times, dfs) - interpolator.py:69-170.

    The bracket is the first node at or after t, found from the front (curves bootstrapped by the reference carry node times
    that differ by an ulp); past the last node the last segment extrapolates.  Index arithmetic is kept as the reference has
    it, including what it does for t before a first node that is not zero (the segment wraps to the last node)."""
    x, d = times, dfs
    n = len(x)
    if t == x[0]:
        return d[0]
    ge = x >= t
    i = int(np.argmax(ge)) if ge.any() else n - 1
    if t > x[i]:
        i = n
    if method == InterpTypes.LINEAR_ZERO_RATES.value:
        if i == 1:                                   # first segment: flat in the first node's zero rate
            z1 = z2 = -math.log(d[1]) / x[1]
            lo, hi = 0, 1
        elif i < n:
            z1, z2 = -math.log(d[i - 1]) / x[i - 1], -math.log(d[i]) / x[i]
            lo, hi = i - 1, i
        else:                                        # beyond the grid: flat in the last zero rate
            z1 = z2 = -math.log(d[n - 1]) / x[n - 1]
            lo, hi = n - 2, n - 1
        z = ((x[hi] - t) * z1 + (t - x[lo]) * z2) / (x[hi] - x[lo])
        return math.exp(-z * t)
    if method == InterpTypes.FLAT_FWD_RATES.value:
        lo, hi = (i - 1, i) if i < n else (n - 2, n - 1)
        y1, y2 = -math.log(d[lo]), -math.log(d[hi])
        return math.exp(-((x[hi] - t) * y1 + (t - x[lo]) * y2) / (x[hi] - x[lo]))
    if method == InterpTypes.LINEAR_FWD_RATES.value:
        if i == 1:
            return math.exp(-t * (-math.log(d[1] + _FWD_SMALL)) / (x[1] + _FWD_SMALL))
        last = -math.log(d[i - 1] / d[i - 2]) / (x[i - 1] - x[i - 2])       # forward rate of the segment before the bracket
        if i < n:
            here = -math.log(d[i] / d[i - 1]) / (x[i] - x[i - 1])
            fwd = ((x[i] - t) * last + (t - x[i - 1]) * here) / (x[i] - x[i - 1])
        else:
            fwd = last
        return d[i - 1] * math.exp(-fwd * (t - x[i - 1]))
    raise LibError("Invalid interpolation scheme.")


_uinterpolate = node_df


def _as_nodes(a):
    return np.asarray(a, dtype=np.float64)


def _vinterpolate(xValues, xvector, dfs, method):
    x, d = _as_nodes(xvector), _as_nodes(dfs)
    return np.array([node_df(float(u), x, d, method) for u in np.asarray(xValues, dtype=np.float64).ravel()])


def interpolate(t, times, dfs, method: int):
    """`interpolate(t, times, dfs, InterpTypes.X.value)` (interpolator.py:35-65): a float gives a float, an array an array."""
    if isinstance(t, (float, np.floating)):
        if t < 0.0:
            raise LibError("Interpolate times must all be >= 0")
        return float(node_df(float(t), _as_nodes(times), _as_nodes(dfs), method))
    if isinstance(t, np.ndarray):
        if np.any(t < 0.0):
            raise LibError("Interpolate times must all be >= 0")
        return _vinterpolate(t, times, dfs, method)
    raise LibError("Unknown input type" + str(type(t)))


_ZERO_RATE_SPLINES = (InterpTypes.PCHIP_ZERO_RATES, InterpTypes.FINCUBIC_ZERO_RATES, InterpTypes.NATCUBIC_ZERO_RATES)
_LOG_DF_SPLINES = (InterpTypes.PCHIP_LOG_DISCOUNT, InterpTypes.NATCUBIC_LOG_DISCOUNT)


class Interpolator:
    """`Interpolator(interp_type).fit(times, dfs)` then `.interpolate(t)` (interpolator.py:197-560)."""

    def __init__(self, interpolator_type: InterpTypes):
        self._interp_type = interpolator_type
        self._interp_fn = None
        self._times = None
        self._dfs = None
        self._refit_curve = False

    def fit(self, times, dfs):
        self._times, self._dfs = times, dfs
        if len(times) == 1:
            return
        kind = self._interp_type
        if kind not in _ZERO_RATE_SPLINES and kind not in _LOG_DF_SPLINES:
            return                                   # node schemes need no fit
        from scipy.interpolate import CubicSpline, PchipInterpolator
        x, d = _as_nodes(times), _as_nodes(dfs)
        if kind in _LOG_DF_SPLINES:
            y = np.log(d)
        else:
            y = -np.log(d) / (x + G_SMALL)
            if x[0] == 0.0:
                y[0] = y[1]                          # the zero rate at t = 0 is the first node's
        if kind in (InterpTypes.PCHIP_ZERO_RATES, InterpTypes.PCHIP_LOG_DISCOUNT):
            self._interp_fn = PchipInterpolator(x, y)
        elif kind == InterpTypes.FINCUBIC_ZERO_RATES:    # zero curvature on the left, zero slope on the right
            self._interp_fn = CubicSpline(x, y, bc_type=((2, 0.0), (1, 0.0)))
        else:
            self._interp_fn = CubicSpline(x, y, bc_type="natural")

    def _uinterpolate(self, t, times, dfs, method):
        return node_df(float(t), _as_nodes(times), _as_nodes(dfs), method)

    def _vinterpolate(self, xValues, xvector, dfs, method):
        out = _vinterpolate(xValues, xvector, dfs, method)
        return out.item() if out.size == 1 else out

    def simple_interpolate(self, t, times, dfs, method: int):
        if isinstance(t, (float, np.floating)):
            if t < 0.0:
                raise LibError("Interpolate times must all be >= 0")
            return self._uinterpolate(t, times, dfs, method)
        if isinstance(t, np.ndarray):
            if np.any(t < 0.0):
                raise LibError("Interpolate times must all be >= 0")
            return self._vinterpolate(t, times, dfs, method)
        raise LibError("Unknown input type" + str(type(t)))

    def interpolate(self, t):
        """A float t gives a float for the node schemes and a one-element array for the spline schemes, as the reference does."""
        if self._dfs is None:
            raise LibError("Dfs have not been set.")
        if isinstance(t, (float, np.floating)):
            if t < 0.0:
                raise LibError("Interpolate times must all be >= 0")
            if abs(t) < G_SMALL:
                return 1.0
            tvec = np.array([t], dtype=np.float64)
        elif isinstance(t, np.ndarray):
            if np.any(t < 0.0):
                raise LibError("Interpolate times must all be >= 0")
            tvec = t
        else:
            raise LibError("t is not a recognized type")
        if self._interp_type in _LOG_DF_SPLINES:
            return np.exp(self._interp_fn(tvec))
        if self._interp_type in _ZERO_RATE_SPLINES:
            return np.exp(-tvec * self._interp_fn(tvec))
        return self._vinterpolate(tvec, self._times, self._dfs, self._interp_type.value)
