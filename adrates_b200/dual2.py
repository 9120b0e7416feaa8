"""Second-order forward-mode numbers (value, gradient, Hessian) over numpy.

The reference obtains the second-order tables of its cross-currency curve by nesting JAX transforms through the
bootstrap scan (`jacfwd(jacrev(...))`, cavour/trades/rates/xccy_curve.py:594-690).  The scan is a short chain of smooth
closed-form operations (+, -, *, /, exp, log), so the same derivatives follow exactly from propagating
    f = (v, g, H)     g = grad f,  H = hess f   w.r.t. a fixed list of n input variables
through it: (fg)'' = f g'' + g f'' + f' g'^T + g' f'^T, (1/f)'' = -f''/f^2 + 2 f' f'^T / f^3, exp / log likewise.
Used on the host for curve construction only (once per curve, a few hundred operations on n <= ~100 variables).
"""
from __future__ import annotations

import numpy as np


class D2:
    __slots__ = ("v", "g", "h")

    def __init__(self, v, g, h):
        self.v, self.g, self.h = float(v), g, h

    # ---- construction
    @staticmethod
    def const(v, n):
        return D2(v, np.zeros(n), np.zeros((n, n)))

    @staticmethod
    def var(v, i, n):
        g = np.zeros(n)
        g[i] = 1.0
        return D2(v, g, np.zeros((n, n)))

    def _lift(self, o):
        return o if isinstance(o, D2) else D2(o, np.zeros_like(self.g), np.zeros_like(self.h))

    # ---- arithmetic
    def __neg__(self):
        return D2(-self.v, -self.g, -self.h)

    def __add__(self, o):
        if isinstance(o, D2):
            return D2(self.v + o.v, self.g + o.g, self.h + o.h)
        return D2(self.v + o, self.g, self.h)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, D2):
            return D2(self.v - o.v, self.g - o.g, self.h - o.h)
        return D2(self.v - o, self.g, self.h)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, D2):
            cross = np.outer(self.g, o.g)
            return D2(self.v * o.v, self.v * o.g + o.v * self.g, self.v * o.h + o.v * self.h + cross + cross.T)
        return D2(self.v * o, self.g * o, self.h * o)

    __rmul__ = __mul__

    def inv(self):
        r = 1.0 / self.v
        return D2(r, -self.g * r * r, -self.h * r * r + 2.0 * np.outer(self.g, self.g) * r * r * r)

    def __truediv__(self, o):
        if isinstance(o, D2):
            return self * o.inv()
        return self * (1.0 / o)

    def __rtruediv__(self, o):
        return self.inv() * o

    def exp(self):
        e = float(np.exp(self.v))
        return D2(e, e * self.g, e * (self.h + np.outer(self.g, self.g)))

    def log(self):
        r = 1.0 / self.v
        return D2(float(np.log(self.v)), self.g * r, self.h * r - np.outer(self.g, self.g) * r * r)


def exp(x):
    return x.exp() if isinstance(x, D2) else float(np.exp(x))


def log(x):
    return x.log() if isinstance(x, D2) else float(np.log(x))


def value(x) -> float:
    return x.v if isinstance(x, D2) else float(x)


def interp(t: float, xp: np.ndarray, yp: list):
    """np.interp(t, xp, yp) for a scalar t and ordinates that may be D2 (piecewise linear, clamped at both ends)."""
    n = len(xp)
    if t <= xp[0]:
        return yp[0]
    if t >= xp[-1]:
        return yp[-1]
    i = int(np.searchsorted(xp, t, side="right")) - 1
    i = min(max(i, 0), n - 2)
    dx = xp[i + 1] - xp[i]
    if dx == 0.0:
        return yp[i]
    w = (t - xp[i]) / dx
    return yp[i] * (1.0 - w) + yp[i + 1] * w
