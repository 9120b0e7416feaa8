"""XccyCurve: foreign-in-domestic discount curve implied by cross-currency basis swaps.

Mirror of cavour/trades/rates/xccy_curve.py for the part the valuation path consumes:
`_times`, `_dfs`, `swap_times`, `basis_spreads`, `_spot_fx`, `_interp_type`, `_dc_type`, `df()` and
`_jac_basis` = d(xccy DFs)/d(pillar basis spreads) (xccy_curve.py:594).  The payment-point plan is host
work (dates, legs); the recursion itself with its first- and second-order tangents w.r.t. the pillar spreads also runs on
the device, batched over shocked spread sets (`device_tables` -> cav_xccy_curve_scan / k_xccy_scan); the tables reach the
valuation kernels through cav_curve_set_tables.

Bootstrap (xccy_curve.py:707-935 plan, :954-1206 recursion), per foreign-leg payment point in
(time, swap) order:
    cashflow      = N*(DF_s/DF_e - 1) [+N on the last payment]  or -+N for a notional exchange,
                    + basis_swap * accrual * N          (DF_s, DF_e log-linear on the foreign OIS nodes)
    DF_inter(t)   = DF_prev * DF_ois(t)/DF_ois(t_prev) * exp(-basis_swap * (t - t_prev))
    at a maturity : DF(t) solves  PV_dom + spot * (-(known PV) - cashflow * DF(t)) = 0
The reference differentiates this scan with jacrev; here the first-order tangents w.r.t. the pillar
spreads are propagated alongside the values (exact forward mode).  The second-order tables the GAMMA
request needs (xccy_curve.py:594-690) are produced on first use by running the same scan on second-order
forward-mode numbers (dual2.D2):
    _hess_basis                 d2(xccy DFs) / d(pillar spreads)^2                    [nodes, nb, nb]
    _jac_foreign_curve_dfs      d(xccy DFs) / d(foreign curve node DFs)               [nodes, nf]
    _mixed_hess_foreign_basis   d2(xccy DFs) / d(pillar spreads) d(foreign node DFs)  [nodes, nb, nf]
For the last two the reference re-derives the payment-time foreign DFs from the foreign curve's node DFs by
log-linear interpolation inside the differentiated function (xccy_curve.py:640-660); so does `_scan`.

`use_ad`: the reference has two builders.  `_build_curve_ad` (use_ad=True: what `Model.build_xccy_curve` passes and the engine
consumes) is the scan above, and this class reproduces its nodes and `_jac_basis` to 7e-16 on random pillar sets, leg
frequencies and spots (tools/reftests/differential_fuzz.py).  `_build_curve` (use_ad=False, the constructor's own default) is
a host loop that projects the foreign forwards through `foreign_curve.df` in that curve's interpolation scheme and discounts
through temporary curves (xccy_curve.py:230-526); on flat-forward OIS curves it gives the same nodes, on LINEAR_ZERO_RATES
ones its discount factors differ from the reference's own AD builder by up to 2e-5.  This class builds the AD nodes for
both settings - KNOWN DIFFERENCE for direct `XccyCurve(..., use_ad=False)` construction over non-flat-forward OIS curves.
"""
from __future__ import annotations

import numpy as np

from .curves import DiscountCurve
from .dates import Date, DayCountTypes, times_from_dates
from .argcheck import check_argument_types
from .error import LibError
from .global_types import InterpTypes


class XccyCurve(DiscountCurve):
    def __init__(self, value_dt: Date, basis_swaps: list, domestic_curve: DiscountCurve, foreign_curve: DiscountCurve,
                 spot_fx: float,
                 interp_type: InterpTypes = InterpTypes.FLAT_FWD_RATES, check_refit: bool = False,
                 use_ad: bool = False):
        check_argument_types(self.__init__, locals())
        if not basis_swaps:
            raise LibError("XccyCurve needs at least one basis swap")
        self._value_dt = value_dt
        self._used_swaps = sorted(basis_swaps, key=lambda s: s._maturity_dt._n)
        self._domestic_curve = domestic_curve
        self._foreign_curve = foreign_curve
        self._spot_fx = spot_fx
        self._interp_type = interp_type
        self._check_refit = check_refit
        self._use_ad = use_ad
        self._dc_type = DayCountTypes.ACT_365F
        self.basis_spreads = [s._foreign_spread for s in self._used_swaps]
        self.swap_times = [(s._maturity_dt - value_dt) / 365.0 for s in self._used_swaps]
        self._bootstrap()

    # ---------------------------------------------------------------------------------
    def _points(self):
        """Foreign-leg payment points of all calibration swaps (xccy_curve.py:720-806).  The
        reference reads them off legs that `value()` has mutated (an effective-date exchange row
        with zero accrual is inserted, swap_float_leg.py:306-318); here the row is built directly."""
        vd, fc = self._value_dt, self._foreign_curve
        pts = []
        for k, swap in enumerate(self._used_swaps):
            leg = swap._foreign_leg
            pv_dom = swap._domestic_leg.value(vd, self._domestic_curve, self._domestic_curve)
            rows = []
            if leg._notional_exchange and leg._effective_dt >= vd:
                rows.append((leg._effective_dt, 0.0, leg._effective_dt, leg._effective_dt))
            rows += list(zip(leg._payment_dts, leg._year_fracs, leg._start_accrued_dts, leg._end_accrued_dts))
            for pay, alpha, start, end in rows:
                if not pay >= vd:
                    continue
                exch = abs(alpha) < 1e-10
                pts.append(dict(
                    time=(pay - vd) / 365.0, swap=k, is_mat=(pay == swap._maturity_dt), at_val=(pay == vd),
                    alpha=alpha, notional=leg._notional, exch=exch,
                    last=(pay == swap._maturity_dt) and leg._notional_exchange,
                    spread_sens=0.0 if exch else alpha * leg._notional,
                    t_start=times_from_dates(start, vd, fc._dc_type), t_end=times_from_dates(end, vd, fc._dc_type),
                    df_ois=fc.df(pay, fc._dc_type), pv_dom=pv_dom))
        pts.sort(key=lambda p: (p["time"], p["swap"]))
        return pts

    def _bootstrap(self):
        pts = self._points()
        nb = len(self._used_swaps)
        fx, fd = np.asarray(self._foreign_curve._times, dtype=np.float64), np.asarray(self._foreign_curve._dfs, dtype=np.float64)
        log_fd = np.log(fd)
        n = len(pts)
        df = np.zeros(n)
        ddf = np.zeros((n, nb))                   # d df / d pillar spreads
        pv_c = np.zeros(n)
        dpv_c = np.zeros((n, nb))
        prev = -1                                  # previous XCCY node (any swap), in time order
        node_idx, seen = [], set()
        for i, p in enumerate(pts):
            k = p["swap"]
            basis = self.basis_spreads[k]
            e_k = np.zeros(nb)
            e_k[k] = 1.0
            if p["exch"]:
                base = p["notional"] if p["last"] else -p["notional"]
            else:
                df_s = np.exp(np.interp(p["t_start"], fx, log_fd))
                df_e = np.exp(np.interp(p["t_end"], fx, log_fd))
                fwd = (df_s / df_e - 1.0) / max(p["alpha"], 1e-10) if p["alpha"] > 1e-10 else 0.0
                base = fwd * p["alpha"] * p["notional"] + (p["notional"] if p["last"] else 0.0)
            cash = base + basis * p["spread_sens"]
            dcash = p["spread_sens"] * e_k
            if prev < 0:
                d_int = p["df_ois"] * np.exp(-basis * p["time"])
                dd_int = -p["time"] * d_int * e_k
            else:
                q = pts[prev]
                grow = (p["df_ois"] / q["df_ois"]) * np.exp(-basis * (p["time"] - q["time"]))
                d_int = df[prev] * grow
                dd_int = ddf[prev] * grow - (p["time"] - q["time"]) * d_int * e_k
            if p["at_val"]:
                pv_c[i], dpv_c[i] = cash, dcash
            elif not p["is_mat"]:
                pv_c[i] = cash * d_int
                dpv_c[i] = dcash * d_int + cash * dd_int
            if p["is_mat"]:
                same = [j for j in range(i) if pts[j]["swap"] == k]
                known = pv_c[same].sum() + pv_c[i]
                dknown = dpv_c[same].sum(axis=0) + dpv_c[i]
                # par: PV_dom + spot * (-(known) - cash * DF) = 0  (foreign legs pay)
                num = -(p["pv_dom"] + self._spot_fx * (-known))
                den = self._spot_fx * (-cash)
                dnum = self._spot_fx * dknown
                dden = -self._spot_fx * dcash
                if abs(den) > 1e-12:
                    df[i] = num / den
                    ddf[i] = (dnum - df[i] * dden) / den
                else:
                    df[i], ddf[i] = d_int, dd_int
            else:
                df[i], ddf[i] = d_int, dd_int
            if not p["at_val"]:
                prev = i
                key = round(p["time"], 4)
                if key not in seen:
                    seen.add(key)
                    node_idx.append(i)
        self._times = np.concatenate([[0.0], [pts[i]["time"] for i in node_idx]])
        self._dfs = np.concatenate([[1.0], df[node_idx]])
        self._repr_dfs = self._dfs
        self._jac_basis = np.vstack([np.zeros((1, nb)), ddf[node_idx]])
        self._pts, self._node_idx = pts, node_idx
        self._second = None                       # second-order tables, built on first use
        if self._check_refit:
            self._check_refits(1e-10)               # the reference's SWAP_TOL (xccy_curve.py:77)

    # ---------------------------------------------------------------------------------
    def scan_plan(self):
        """Per-point arrays of the bootstrap recursion for the device scan (cav_xccy_curve_scan): everything that does not
        depend on the pillar spreads - times, swap index, flags (1 exchange | 4 on the valuation date | 8 maturity), spread
        sensitivities, the cashflow at zero spread (forward rate from the foreign curve, xccy_curve.py:636-639), foreign DFs at
        the payment dates and the domestic-leg PVs."""
        pts = self._pts
        fx, fd = np.asarray(self._foreign_curve._times, dtype=np.float64), np.asarray(self._foreign_curve._dfs, dtype=np.float64)
        log_fd = np.log(fd)
        n = len(pts)
        base = np.zeros(n)
        flags = np.zeros(n, dtype=np.int32)
        for i, p in enumerate(pts):
            if p["exch"]:
                base[i] = p["notional"] if p["last"] else -p["notional"]
            else:
                df_s = np.exp(np.interp(p["t_start"], fx, log_fd))
                df_e = np.exp(np.interp(p["t_end"], fx, log_fd))
                fwd = (df_s / df_e - 1.0) / max(p["alpha"], 1e-10) if p["alpha"] > 1e-10 else 0.0
                base[i] = fwd * p["alpha"] * p["notional"] + (p["notional"] if p["last"] else 0.0)
            flags[i] = (1 if p["exch"] else 0) | (4 if p["at_val"] else 0) | (8 if p["is_mat"] else 0)
        return dict(time=np.array([p["time"] for p in pts]), swap=np.array([p["swap"] for p in pts], dtype=np.int32), flags=flags,
                    sens=np.array([p["spread_sens"] for p in pts]), base=base, df_ois=np.array([p["df_ois"] for p in pts]),
                    pv_dom=np.array([p["pv_dom"] for p in pts]))

    def device_tables(self, ctx=None, spreads=None, order: int = 2):
        """The curve's node tables from the device bootstrap (k_xccy_scan): (dfs [S, nodes], jac_basis [S, nodes, nb] | None,
        hess_basis [S, nodes, nb, nb] | None) for S spread sets (default: the curve's own spreads, S = 1), rows in the order of
        `_times` (row 0 = the (0, 1.0) node).  Shocked basis curves re-bootstrap in one launch: spreads[S][nb]."""
        from . import _native
        own = ctx is None
        if own:
            ctx = _native.Context(0)
        try:
            sp = np.atleast_2d(np.asarray(self.basis_spreads if spreads is None else spreads, dtype=np.float64))
            pl = self.scan_plan()
            df, jac, hess = ctx.xccy_curve_scan(pl["time"], pl["swap"], pl["flags"], pl["sens"], pl["base"], pl["df_ois"], pl["pv_dom"],
                                                self._spot_fx, sp, order=order)
        finally:
            if own:
                ctx.close()
        idx = np.asarray(self._node_idx, dtype=np.int64)
        S, nb = sp.shape
        dfs = np.concatenate([np.ones((S, 1)), df[:, idx]], axis=1)
        jn = None if jac is None else np.concatenate([np.zeros((S, 1, nb)), jac[:, idx]], axis=1)
        hn = None if hess is None else np.concatenate([np.zeros((S, 1, nb, nb)), hess[:, idx]], axis=1)
        return dfs, jn, hn

    def _scan(self, basis, df_ois):
        """The bootstrap recursion of `_bootstrap` on generic numbers (floats or dual2.D2): basis[k] = spread of pillar k,
        df_ois[i] = foreign OIS discount factor at payment point i.  Returns the XCCY discount factor of every point."""
        from .dual2 import exp
        pts, spot = self._pts, self._spot_fx
        fx, fd = np.asarray(self._foreign_curve._times, dtype=np.float64), np.asarray(self._foreign_curve._dfs, dtype=np.float64)
        log_fd = np.log(fd)
        df, pv_c = [None] * len(pts), [0.0] * len(pts)
        prev = -1
        for i, p in enumerate(pts):
            k = p["swap"]
            b = basis[k]
            if p["exch"]:
                base = p["notional"] if p["last"] else -p["notional"]
            else:       # forward rates are payment-level constants of the differentiated function (xccy_curve.py:636-639)
                df_s = np.exp(np.interp(p["t_start"], fx, log_fd))
                df_e = np.exp(np.interp(p["t_end"], fx, log_fd))
                fwd = (df_s / df_e - 1.0) / max(p["alpha"], 1e-10) if p["alpha"] > 1e-10 else 0.0
                base = fwd * p["alpha"] * p["notional"] + (p["notional"] if p["last"] else 0.0)
            cash = b * p["spread_sens"] + base
            if prev < 0:
                d_int = df_ois[i] * exp(-(b * p["time"]))
            else:
                q = pts[prev]
                d_int = df[prev] * (df_ois[i] / df_ois[prev]) * exp(-(b * (p["time"] - q["time"])))
            if p["at_val"]:
                pv_c[i] = cash
            elif not p["is_mat"]:
                pv_c[i] = cash * d_int
            if p["is_mat"]:
                known = pv_c[i]
                for j in range(i):
                    if pts[j]["swap"] == k:
                        known = known + pv_c[j]
                num = -(-(known * spot) + p["pv_dom"])
                den = -(cash * spot)
                from .dual2 import value
                df[i] = num / den if abs(value(den)) > 1e-12 else d_int
            else:
                df[i] = d_int
            if not p["at_val"]:
                prev = i
        return df

    def _second_order(self):
        """(_hess_basis, _jac_foreign_curve_dfs, _mixed_hess_foreign_basis); row 0 is the (0, 1.0) node."""
        if self._second is None:
            from .dual2 import D2, exp, interp, log
            pts, idx = self._pts, self._node_idx
            nb = len(self._used_swaps)
            # (a) w.r.t. the pillar spreads, payment-time foreign DFs held at their stored values (xccy_curve.py:578-606)
            basis = [D2.var(s, k, nb) for k, s in enumerate(self.basis_spreads)]
            dfa = self._scan(basis, [p["df_ois"] for p in pts])
            jac = np.vstack([np.zeros((1, nb))] + [dfa[i].g[None, :] for i in idx])
            hess = np.concatenate([np.zeros((1, nb, nb))] + [dfa[i].h[None, :, :] for i in idx])
            if not np.allclose(jac, self._jac_basis, rtol=1e-10, atol=1e-14):
                raise LibError("XccyCurve: second-order scan disagrees with the first-order tangents")
            # (b) w.r.t. (pillar spreads, foreign curve node DFs): payment-time DFs by log-linear interpolation of the node DFs
            fx = np.asarray(self._foreign_curve._times, dtype=np.float64)
            fd = np.asarray(self._foreign_curve._dfs, dtype=np.float64)
            nf = fd.shape[0]
            n = nb + nf
            basis = [D2.var(s, k, n) for k, s in enumerate(self.basis_spreads)]
            log_nodes = [log(D2.var(d, nb + j, n)) for j, d in enumerate(fd)]
            dfb = self._scan(basis, [exp(interp(p["time"], fx, log_nodes)) for p in pts])
            jac_f = np.vstack([np.zeros((1, nf))] + [dfb[i].g[None, nb:] for i in idx])
            mixed = np.concatenate([np.zeros((1, nb, nf))] + [dfb[i].h[None, :nb, nb:] for i in idx])
            self._second = (hess, jac_f, mixed)
        return self._second

    @property
    def _hess_basis(self):
        return self._second_order()[0]

    @property
    def _jac_foreign_curve_dfs(self):
        return self._second_order()[1]

    @property
    def _mixed_hess_foreign_basis(self):
        return self._second_order()[2]

    # ---------------------------------------------------------------------------------
    def df(self, dt, day_count=None):
        """Always ACT/365F, whatever day count is passed (xccy_curve.py:1210-1234)."""
        return DiscountCurve.df(self, dt, DayCountTypes.ACT_365F)

    def _check_refits(self, swap_tol: float):
        """Reprices the calibration swaps through the non-AD `XccyBasisSwap.value` (xccy_curve.py:1238-1272).  Passes when the
        OIS curves interpolate flat-forward; with another scheme the bootstrap's log-linear forward projection and the legs'
        curve look-ups differ by ~1e-5 of the notional and the check raises - in the reference as here."""
        for swap in self._used_swaps:
            v = swap.value(value_dt=self._value_dt, domestic_discount_curve=self._domestic_curve,
                           foreign_discount_curve=self._foreign_curve, xccy_discount_curve=self, spot_fx=self._spot_fx)
            v_normalized = v / swap._domestic_notional
            if abs(v_normalized) > swap_tol:
                raise LibError(f"XCCY swap with maturity {swap._maturity_dt} not repriced. "
                               f"Difference is {abs(v_normalized)}")

    def par_residuals(self) -> list:
        """(domestic PV + spot x foreign PV) / domestic notional of every calibration swap on the bootstrapped curve - the
        condition the bootstrap solves (xccy_curve.py:465-474).  Zero to rounding when the OIS curves interpolate
        flat-forward (the bootstrap projects forwards log-linearly).  A diagnostic beyond the reference's _check_refits."""
        out = []
        for swap in self._used_swaps:
            pv_dom = swap._domestic_leg.value(self._value_dt, self._domestic_curve, self._domestic_curve)
            pv_for = swap._foreign_leg.value(self._value_dt, self, self._foreign_curve)
            out.append((pv_dom + self._spot_fx * pv_for) / swap._domestic_notional)
        return out
