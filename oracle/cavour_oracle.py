"""CPU oracle for the Cavour/ADRates valuation-and-Greeks hot path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  Nothing under
adrates_b200/ imports it; the product path fails loudly without its CUDA library.

What it restates (all citations into /root/reference):
  plan_path_b / bootstrap_dfs     Engine.build_curve_ad                engine.py:2246-2360
  bootstrap_tables                Engine._cached_curve (jacrev/hessian of the scan, done
                                  here as exact forward-mode tangents)  engine.py:2362-2412
  simple_interpolate              InterpolatorAd.simple_interpolate    interpolator_ad.py:186-249
  fixed_leg_pv / float_leg_pv     Engine._price_fixed_leg_jax          engine.py:2414-2448
                                  Engine._float_leg_jax                engine.py:2639-2728
  leg_analytics                   _fixed_leg_analytics / _float_leg_analytics: grad and dense
                                  G x G Hessian w.r.t. node DFs, `g @ J * 1e-4`,
                                  `(J.T H J + sum_k g_k C_k) * 1e-8`   engine.py:2498-2576, 2808-2934
  ois_analytics                   Engine._compute_ois_natural          engine.py:153-215
  path_a_bootstrap / df_ad        OISCurve._build_curve_ad             ois_curve.py:156-212
                                  DiscountCurve._linear_forward_interp discount_curve.py:385-415

The differentiation the reference delegates to JAX (a third-party dependency that is not
vendored and not pinned: requirements.txt does not list it) is restated as exact
second-order forward-mode arithmetic on sparse node sets (class Dual2), since AD of these
closed-form expressions is exact up to round-off.  jnp.interp / searchsorted(side='right')
duplicate-node semantics follow jax/_src/numpy/lax_numpy.py::_interp.

PINNED: tests/test_oracle_golden.py checks this module against (1) the reference's executed
notebook values (notebooks/intro.ipynb cells 24-44) and (2) outputs of the unmodified
reference engine run in the build container for 27 trades on 6 curves, including its AD
Jacobian/Hessian tables (tests/golden/*, generator tests/golden/gen/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

LINEAR_ZERO_RATES = 4   # InterpTypes values, global_types.py:76-84
FLAT_FWD_RATES = 1


# --------------------------------------------------------------------------------------
# path B: plan + recursion + tangents
# --------------------------------------------------------------------------------------
def plan_path_b(swap_times, year_fracs):
    """engine.py:2283-2328.  Returns dict(times, acc, swap, prev)."""
    points = [dict(maturity=0.0, key=0.0, acc=0.0, prev_key=None, swap=0)]
    for i, fracs in enumerate(year_fracs):
        cumsum = 0.0
        for j, frac in enumerate(fracs):
            prev_cum = cumsum
            cumsum += frac
            points.append(dict(maturity=cumsum, key=round(cumsum, 2), acc=frac,
                               prev_key=round(prev_cum, 2) if j > 0 else None, swap=i))
    pts = sorted(points, key=lambda p: p["maturity"])
    lookup = {}
    for idx, p in enumerate(pts):
        if p["key"] not in lookup:
            lookup[p["key"]] = idx
    prev = [-1 if p["prev_key"] is None else lookup.get(p["prev_key"], -1) for p in pts]
    return dict(times=np.array([p["maturity"] for p in pts]), acc=np.array([p["acc"] for p in pts]),
                swap=np.array([p["swap"] for p in pts], dtype=np.int64), prev=np.array(prev, dtype=np.int64))


def bootstrap_dfs(rates, plan):
    """The lax.scan body, engine.py:2337-2349."""
    G = len(plan["times"])
    pv01 = np.zeros(G)
    dfs = np.zeros(G)
    for i in range(G):
        r = rates[plan["swap"][i]]
        acc = plan["acc"][i]
        p = plan["prev"][i]
        prev_pv01 = 0.0 if p < 0 else pv01[p]
        dfs[i] = 1.0 / (1.0 + r * acc) if p < 0 else (1.0 - r * prev_pv01) / (1.0 + r * acc)
        pv01[i] = prev_pv01 + acc * dfs[i]
    return dfs


def bootstrap_tables(rates, plan):
    """dfs[G], jac[G,R] = d dfs/d rates, hess[G,R,R] (what jacrev/hessian return at
    engine.py:2388-2389), by propagating first/second-order tangents through the recursion."""
    rates = np.asarray(rates, dtype=np.float64)
    G, R = len(plan["times"]), len(rates)
    d = np.zeros(G)
    P = np.zeros(G)
    dP = np.zeros((G, R))
    d2P = np.zeros((G, R, R))
    J = np.zeros((G, R))
    C = np.zeros((G, R, R))
    for i in range(G):
        s = int(plan["swap"][i])
        r = rates[s]
        a = plan["acc"][i]
        p = int(plan["prev"][i])
        e = np.zeros(R)
        e[s] = 1.0
        if p < 0:
            Pp, dPp, d2Pp = 0.0, np.zeros(R), np.zeros((R, R))
        else:
            Pp, dPp, d2Pp = P[p], dP[p], d2P[p]
        u = 1.0 - r * Pp
        v = 1.0 + r * a
        du = -(e * Pp + r * dPp)
        dv = a * e
        d2u = -(np.outer(e, dPp) + np.outer(dPp, e) + r * d2Pp)
        d[i] = u / v
        J[i] = (du - d[i] * dv) / v
        C[i] = (d2u - np.outer(J[i], dv) - np.outer(dv, J[i])) / v
        P[i] = Pp + a * d[i]
        dP[i] = dPp + a * J[i]
        d2P[i] = d2Pp + a * C[i]
    return d, J, C


# --------------------------------------------------------------------------------------
# interpolation with sparse second-order duals w.r.t. node DFs
# --------------------------------------------------------------------------------------
class Dual2:
    """value + sparse gradient {node: g} + sparse Hessian {(n, m): h} (both orders kept)."""
    __slots__ = ("v", "g", "h")

    def __init__(self, v, g=None, h=None):
        self.v, self.g, self.h = v, g or {}, h or {}

    @staticmethod
    def _lift(o):
        return o if isinstance(o, Dual2) else Dual2(float(o))

    def __add__(self, o):
        o = Dual2._lift(o)
        g = dict(self.g)
        for k, x in o.g.items():
            g[k] = g.get(k, 0.0) + x
        h = dict(self.h)
        for k, x in o.h.items():
            h[k] = h.get(k, 0.0) + x
        return Dual2(self.v + o.v, g, h)

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, o):
        return self + (-Dual2._lift(o))

    def __rsub__(self, o):
        return Dual2._lift(o) + (-self)

    def __mul__(self, o):
        if not isinstance(o, Dual2):
            c = float(o)
            return Dual2(self.v * c, {k: x * c for k, x in self.g.items()}, {k: x * c for k, x in self.h.items()})
        g = {}
        for k, x in self.g.items():
            g[k] = g.get(k, 0.0) + x * o.v
        for k, x in o.g.items():
            g[k] = g.get(k, 0.0) + x * self.v
        h = {}
        for k, x in self.h.items():
            h[k] = h.get(k, 0.0) + x * o.v
        for k, x in o.h.items():
            h[k] = h.get(k, 0.0) + x * self.v
        for n, gn in self.g.items():
            for m, gm in o.g.items():
                h[(n, m)] = h.get((n, m), 0.0) + gn * gm
                h[(m, n)] = h.get((m, n), 0.0) + gn * gm
        return Dual2(self.v * o.v, g, h)

    __rmul__ = __mul__

    def recip(self):
        f = 1.0 / self.v
        f1 = -f * f
        f2 = 2.0 * f * f * f
        g = {k: f1 * x for k, x in self.g.items()}
        h = {k: f1 * x for k, x in self.h.items()}
        for n, gn in self.g.items():
            for m, gm in self.g.items():
                h[(n, m)] = h.get((n, m), 0.0) + f2 * gn * gm
        return Dual2(f, g, h)

    def __truediv__(self, o):
        if not isinstance(o, Dual2):
            return self * (1.0 / float(o))
        return self * o.recip()

    def __rtruediv__(self, o):
        return Dual2._lift(o) * self.recip()


def _interp_bracket(tt_adj, x):
    """jnp.interp's cell choice: i = clip(searchsorted(x, t, 'right'), 1, n-1)."""
    i = int(np.clip(np.searchsorted(x, tt_adj, side="right"), 1, len(x) - 1))
    return i - 1, i


def simple_interpolate(tt, x, d, method, dual=True):
    """One query of InterpolatorAd.simple_interpolate (interpolator_ad.py:210-243) as a
    function of the node DFs `d`.  Returns Dual2 (or float when dual=False)."""
    distances = np.abs(tt - x)
    grid_idx = int(np.argmin(distances))           # first index on ties
    at_grid = distances[grid_idx] < 1e-10
    if at_grid:                                    # lax.select picks d[grid_idx]; no flow elsewhere
        v = float(d[grid_idx])
        return Dual2(v, {grid_idx: 1.0}) if dual else v
    tt_adj = tt + 1e-12
    a, b = _interp_bracket(tt_adj, x)
    if tt_adj > x[-1]:
        nodes = [(len(x) - 1, 1.0)]                # jnp.interp clamps to fp[-1]
    elif tt_adj < x[0]:
        nodes = [(0, 1.0)]
    else:
        dx = x[b] - x[a]
        if abs(dx) <= np.spacing(np.finfo(np.float64).eps):
            nodes = [(a, 1.0)]
        else:
            w = (tt_adj - x[a]) / dx
            nodes = [(a, 1.0 - w), (b, w)]
    # interpolated quantity y = sum_k c_k * f(d_k); DF = exp(-y * scale)
    if method == LINEAR_ZERO_RATES:
        # r_k = -ln d_k / max(x_k, 1e-15);  DF = exp(-interp(r) * tt)
        coef = [(k, c * tt / max(x[k], 1e-15)) for k, c in nodes]
    elif method == FLAT_FWD_RATES:
        coef = [(k, c) for k, c in nodes]
    else:
        raise ValueError("Invalid interpolation scheme.")
    # ln DF = sum_k w_k ln d_k
    ln_df = sum(w * math.log(d[k]) for k, w in coef)
    v = math.exp(ln_df)
    if not dual:
        return v
    g, h = {}, {}
    for k, w in coef:
        g[k] = g.get(k, 0.0) + v * w / d[k]
    for k, w in coef:
        for m, wm in coef:
            h[(k, m)] = h.get((k, m), 0.0) + v * w * wm / (d[k] * d[m])
        h[(k, k)] = h.get((k, k), 0.0) - v * w / (d[k] * d[k])
    return Dual2(v, g, h)


# --------------------------------------------------------------------------------------
# leg pricers as functions of the node DFs
# --------------------------------------------------------------------------------------
def fixed_leg_pv(x, d, method, payment_times, payments, principal, leg_sign, value_time, dual=True):
    """engine.py:2425-2448 (mask is strict: payment_times > value_time)."""
    df_val = simple_interpolate(value_time, x, d, method, dual)
    total = Dual2(0.0) if dual else 0.0
    last_rel = None
    for t, pay in zip(payment_times, payments):
        df_rel = simple_interpolate(t, x, d, method, dual) / df_val
        last_rel = df_rel
        if t > value_time:
            total = total + df_rel * pay
    if len(payment_times) and payment_times[-1] > value_time:
        total = total + last_rel * principal
    return total * leg_sign


def float_leg_pv(x, d, method, payment_times, start_times, end_times, pay_alphas, spreads, notionals, principal,
                 leg_sign, value_time, first_fixing_rate=0.0, override_first=False, dual=True):
    """engine.py:2662-2728 on a single curve (mask is >=)."""
    df_val = simple_interpolate(value_time, x, d, method, dual)
    total = Dual2(0.0) if dual else 0.0
    last_rel = None
    for i in range(len(payment_times)):
        df_s = simple_interpolate(start_times[i], x, d, method, dual)
        df_e = simple_interpolate(end_times[i], x, d, method, dual)
        if pay_alphas[i] > 0:
            fwd = (df_s / df_e - 1.0) / pay_alphas[i]
        else:
            fwd = Dual2(0.0) if dual else 0.0
        if override_first and i == 0:
            fwd = Dual2(first_fixing_rate) if dual else first_fixing_rate
        cf = (fwd + spreads[i]) * (pay_alphas[i] * notionals[i])
        df_rel = simple_interpolate(payment_times[i], x, d, method, dual) / df_val
        last_rel = df_rel
        if payment_times[i] >= value_time:
            total = total + cf * df_rel
    if len(payment_times) and payment_times[-1] >= value_time:
        total = total + last_rel * principal
    return total * leg_sign


def leg_analytics(pv: Dual2, jac, hess_curve):
    """value, delta[R], gamma[R,R] with the reference's dense chain rule
    (engine.py:2551-2568): sensitivities = grad_dfs @ jac, *1e-4;
    gammas = (jac.T @ hess_dfs @ jac + sum_k grad_dfs[k] * hess_curve[k]) * 1e-8."""
    G = jac.shape[0]
    grad_dfs = np.zeros(G)
    for k, v in pv.g.items():
        grad_dfs[k] = v
    hess_dfs = np.zeros((G, G))
    for (n, m), v in pv.h.items():
        hess_dfs[n, m] = v
    delta = (grad_dfs @ jac) * 1e-4
    term1 = jac.T @ hess_dfs @ jac
    term2 = np.sum(grad_dfs[:, None, None] * hess_curve, axis=0)
    return pv.v, delta, (term1 + term2) * 1e-8


def ois_analytics(curve_tables, method, fixed, floating):
    """Engine._compute_ois_natural (engine.py:153-189): fixed + floating leg analytics.
    curve_tables = (times, dfs, jac, hess).  `fixed` / `floating` are dicts of the leg
    arrays the engine extracts (engine.py:2519-2527, 2858-2877)."""
    x, d, jac, hess = curve_tables
    pv_fix = fixed_leg_pv(x, d, method, fixed["payment_times"], fixed["payments"], fixed.get("principal", 0.0),
                          fixed["leg_sign"], fixed.get("value_time", 0.0))
    pv_flt = float_leg_pv(x, d, method, floating["payment_times"], floating["start_times"], floating["end_times"],
                          floating["pay_alphas"], floating["spreads"], floating["notionals"],
                          floating.get("principal", 0.0), floating["leg_sign"], floating.get("value_time", 0.0))
    v1, d1, g1 = leg_analytics(pv_fix, jac, hess)
    v2, d2, g2 = leg_analytics(pv_flt, jac, hess)
    return v1 + v2, d1 + d2, g1 + g2


def ois_value_only(x, d, method, fixed, floating):
    """VALUE without derivatives (used by scenario-revaluation checks)."""
    a = fixed_leg_pv(x, d, method, fixed["payment_times"], fixed["payments"], fixed.get("principal", 0.0),
                     fixed["leg_sign"], fixed.get("value_time", 0.0), dual=False)
    b = float_leg_pv(x, d, method, floating["payment_times"], floating["start_times"], floating["end_times"],
                     floating["pay_alphas"], floating["spreads"], floating["notionals"],
                     floating.get("principal", 0.0), floating["leg_sign"], floating.get("value_time", 0.0), dual=False)
    return a + b


# --------------------------------------------------------------------------------------
# path A + df_ad
# --------------------------------------------------------------------------------------
def path_a_bootstrap(swap_rates, swap_times, year_fracs):
    """OISCurve._build_curve_ad (ois_curve.py:156-212)."""
    st = np.array(swap_times)
    lr = np.log(np.array(swap_rates))
    pv01_dict = {}
    times, dfs = [0.0], [1.0]

    def calc(i, target=None, step=0):
        if target is None:
            t_mat, rate = swap_times[i], swap_rates[i]
        else:
            t_mat, rate = target, float(np.exp(np.interp(target, st, lr)))
        fr = year_fracs[i]
        if len(fr) == 1:
            acc = fr[0]
            df_mat = 1.0 / (acc * rate + 1.0)
            pv01 = acc * df_mat
        else:
            acc = fr[-1 - step]
            last_payment = sum(fr[:-1 - step])
            if round(last_payment, 2) not in pv01_dict:
                step += 1
                pv01_dict[round(last_payment, 2)] = calc(i, last_payment, step)
            df_mat = (1.0 - rate * pv01_dict[round(last_payment, 2)]) / (acc * rate + 1)
            pv01 = pv01_dict[round(last_payment, 2)] + acc * df_mat
        times.append(t_mat)
        dfs.append(df_mat)
        pv01_dict[round(t_mat, 2)] = pv01
        return pv01

    for i in range(len(swap_rates)):
        calc(i)
    return np.array(times), np.array(dfs)


def df_ad(t, times, dfs):
    """DiscountCurve._linear_forward_interp (discount_curve.py:401-415)."""
    t = np.asarray(t, dtype=np.float64)
    fwd = -np.log(dfs[1:] / dfs[:-1]) / (times[1:] - times[:-1])
    f = np.interp(t, times[:-1], fwd)
    i0 = np.searchsorted(times, t, side="right") - 1
    return dfs[i0] * np.exp(-f * (t - times[i0]))
