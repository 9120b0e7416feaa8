"""ctypes binding of the C ABI in include/adrates_b200.h.

The product path has no CPU fallback: if the shared library is missing or no CUDA device
is usable, calls raise LibError instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .error import LibError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadrates_b200.so")

REQ_VALUE, REQ_DELTA, REQ_GAMMA, REQ_ALLREDUCE = 1, 2, 4, 8
COMM_HANDLE_BYTES = 96
NOUT = 1057

_P = C.c_void_p
_SIGS = {
    "cav_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "cav_destroy": (None, [_P]),
    "cav_last_error": (C.c_char_p, [_P]),
    "cav_version": (C.c_int, []),
    "cav_sync": (C.c_int, [_P]),
    "cav_timer_start": (C.c_int, [_P]),
    "cav_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "cav_launch_count": (C.c_int64, [_P]),
    "cav_set_stream": (C.c_int, [_P, _P]),
    "cav_profile": (C.c_int, [_P, C.c_int]),
    "cav_set_async_upload": (C.c_int, [_P, C.c_int]),
    "cav_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "cav_curve_build": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "cav_curve_read": (C.c_int, [_P, _P, _P, _P]),
    "cav_curve_set_tables": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int]),
    "cav_curve_rebuild_dev": (C.c_int, [_P, _P]),
    "cav_df_ad": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int64, _P]),
    "cav_xccy_curve_scan": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_double, _P, C.c_int, C.c_int, _P, _P, _P]),
    "cav_portfolio_upload": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int, _P, _P, _P, C.c_int64, C.c_int, _P,
                                       C.c_int64, _P, _P, _P, _P]),
    "cav_portfolio_set_tiles": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_int, _P, _P,
                                          _P]),
    "cav_portfolio_value": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "cav_portfolio_value_host": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "cav_portfolio_delta_gemm": (C.c_int, [_P, _P, _P, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "cav_scenarios": (C.c_int, [_P, _P, C.c_int, _P]),
    "cav_scenarios_info": (C.c_int, [_P, _P]),
    "cav_curve_df": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, C.c_int64, _P]),
    "cav_cashflow_pv": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_double, C.c_int64, _P, _P, _P, _P, _P]),
    "cav_cashflow_pv_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_double, C.c_int64, _P, _P, _P, _P, _P]),
    "cav_book_set_holidays": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "cav_book_from_arrays": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_uint32]),
    "cav_book_info": (C.c_int, [_P, _P]),
    "cav_comm_local_handle": (C.c_int, [_P, _P]),
    "cav_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "cav_comm_status": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cav_book_read": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cav_book_read_tiles": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}
E_INVALID, E_CUDA, E_STATE, E_UNSUPPORTED = -1, -2, -3, -4
TENOR_YEARS, TENOR_MONTHS = 0, 1
BOOK_TILES, BOOK_DATES_I32, BOOK_SIGN_I8 = 1, 2, 4      # flags of cav_book_from_arrays


class BookConv(C.Structure):
    """cav_book_conv (include/adrates_b200.h)."""
    _fields_ = [("value_dt", C.c_int64), ("fixed_freq_months", C.c_int32), ("float_freq_months", C.c_int32),
                ("fixed_dc", C.c_int32), ("float_dc", C.c_int32), ("cal_type", C.c_int32), ("bd_type", C.c_int32),
                ("dg_type", C.c_int32), ("end_of_month", C.c_int32), ("payment_lag", C.c_int32)]

EXPORTS = tuple(_SIGS)

_dll = None


def load_dll():
    """dlopen the library and type its entry points (no CUDA call is made)."""
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise LibError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the valuation path)")
        dll = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(dll, name)
            fn.restype, fn.argtypes = res, args
        _dll = dll
    return _dll


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)   # raw address (e.g. torch tensor .data_ptr())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One device context (cav_ctx)."""

    def __init__(self, device: int = 0):
        self._dll = load_dll()
        h = _P()
        rc = self._dll.cav_create(C.byref(h), int(device))
        if rc != 0 or not h:
            raise LibError(f"cav_create(device={device}) failed with code {rc}: no usable CUDA device "
                           "(the valuation path has no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._dll.cav_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LibError(f"adrates_b200 native error {rc}: {self._dll.cav_last_error(self._h).decode()}", rc)

    # ---- misc
    def sync(self):
        self._ck(self._dll.cav_sync(self._h))

    def timer_start(self):
        self._ck(self._dll.cav_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self._dll.cav_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def set_stream(self, cuda_stream: int):
        self._ck(self._dll.cav_set_stream(self._h, int(cuda_stream)))

    def set_async_upload(self, enable: bool):
        """Pipeline the per-trade arrays of portfolio_upload behind the units kernel (see the header for the
        lifetime contract of the host buffers)."""
        self._ck(self._dll.cav_set_async_upload(self._h, 1 if enable else 0))

    def profile(self, enable: bool):
        self._ck(self._dll.cav_profile(self._h, 1 if enable else 0))

    def last_kernel_ms(self):
        ms = (C.c_float * 3)()
        self._ck(self._dll.cav_last_kernel_ms(self._h, ms))
        return [float(x) for x in ms]

    def launch_count(self) -> int:
        return int(self._dll.cav_launch_count(self._h))

    # ---- curve
    def curve_build(self, interp_method: int, swap_rates, plan, order: int = 2):
        r = _f64(swap_rates)
        t, a = _f64(plan.node_time), _f64(plan.node_acc)
        s = np.ascontiguousarray(plan.node_swap, dtype=np.int32)
        p = np.ascontiguousarray(plan.node_prev, dtype=np.int32)
        self._ck(self._dll.cav_curve_build(self._h, int(interp_method), _ptr(r), r.shape[0], _ptr(t), _ptr(a),
                                           _ptr(s), _ptr(p), t.shape[0], int(order)))
        self._G, self._R = t.shape[0], r.shape[0]

    def curve_rebuild_dev(self, swap_rates_dev):
        self._ck(self._dll.cav_curve_rebuild_dev(self._h, _ptr(swap_rates_dev)))

    def curve_set_tables(self, dfs, jac=None, hess=None):
        d = _f64(dfs)
        J = None if jac is None else _f64(jac)
        H = None if hess is None else _f64(hess)
        R = 1 if J is None else J.shape[1]
        self._ck(self._dll.cav_curve_set_tables(self._h, _ptr(d), _ptr(J), _ptr(H), d.shape[0], R))
        self._G, self._R = d.shape[0], R

    def curve_read(self, jac=True, hess=True):
        G, R = self._G, self._R
        d = np.empty(G)
        J = np.empty((G, R)) if jac else None
        H = np.empty((G, R, R)) if hess else None
        self._ck(self._dll.cav_curve_read(self._h, _ptr(d), _ptr(J), _ptr(H)))
        return d, J, H

    def df_ad(self, node_time, node_df, t):
        x, d, tt = _f64(node_time), _f64(node_df), _f64(t)
        out = np.empty_like(tt)
        self._ck(self._dll.cav_df_ad(self._h, _ptr(x), _ptr(d), x.shape[0], _ptr(tt), tt.shape[0], _ptr(out)))
        return out

    def xccy_curve_scan(self, pt_time, pt_swap, pt_flags, pt_sens, pt_base, pt_df_ois, pt_pv_dom, spot_fx: float, spreads, order: int = 1):
        """XccyCurve bootstrap on the device for one or more spread sets: (dfs [S, n], jac [S, n, nb] | None, hess [S, n, nb, nb] | None)."""
        tt, ss, bb, oo, pp = _f64(pt_time), _f64(pt_sens), _f64(pt_base), _f64(pt_df_ois), _f64(pt_pv_dom)
        sw = np.ascontiguousarray(pt_swap, dtype=np.int32)
        fl = np.ascontiguousarray(pt_flags, dtype=np.int32)
        sp = np.atleast_2d(_f64(spreads))
        n, (S, nb) = tt.shape[0], sp.shape
        for arr in (ss, bb, oo, pp, sw, fl):
            if arr.shape[0] != n:
                raise LibError("xccy_curve_scan: per-point arrays differ in length")
        df = np.empty((S, n))
        jac = np.empty((S, n, nb)) if order >= 1 else None
        hess = np.empty((S, n, nb, nb)) if order >= 2 else None
        self._ck(self._dll.cav_xccy_curve_scan(self._h, n, nb, _ptr(tt), _ptr(sw), _ptr(fl), _ptr(ss), _ptr(bb), _ptr(oo), _ptr(pp),
                                               float(spot_fx), _ptr(sp), S, int(order), _ptr(df),
                                               _ptr(jac) if jac is not None else None, _ptr(hess) if hess is not None else None))
        return df, jac, hess

    def curve_df(self, interp_method: int, node_time, node_df, t):
        """DiscountCurve.df / Interpolator._uinterpolate on the path-A nodes."""
        x, d, tt = _f64(node_time), _f64(node_df), _f64(t)
        out = np.empty_like(tt)
        self._ck(self._dll.cav_curve_df(self._h, int(interp_method), _ptr(x), _ptr(d), x.shape[0], _ptr(tt),
                                        tt.shape[0], _ptr(out)))
        return out

    def cashflow_pv(self, interp_method: int, node_time, node_df, t_value: float, offsets, t, amt):
        """PV per trade of explicit cashflows on the path-A nodes; returns (pv[n_trades], total)."""
        x, d, tt, aa = _f64(node_time), _f64(node_df), _f64(t), _f64(amt)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        n = off.shape[0] - 1
        pv, tot = np.zeros(n), C.c_double()
        self._ck(self._dll.cav_cashflow_pv(self._h, int(interp_method), _ptr(x), _ptr(d), x.shape[0], float(t_value), n,
                                           _ptr(off), _ptr(tt), _ptr(aa), _ptr(pv), C.byref(tot)))
        return pv, float(tot.value)

    def cashflow_pv_dev(self, interp_method: int, node_time, node_df, t_value: float, n_trades: int, offsets_dev, t_dev, amt_dev,
                        pv_dev, total_dev=None):
        """cashflow_pv on device-resident cashflow arrays (raw device addresses / tensors' data_ptr()); asynchronous."""
        x, d = _f64(node_time), _f64(node_df)
        self._ck(self._dll.cav_cashflow_pv_dev(self._h, int(interp_method), _ptr(x), _ptr(d), x.shape[0], float(t_value), int(n_trades),
                                               _ptr(offsets_dev), _ptr(t_dev), _ptr(amt_dev), _ptr(pv_dev), _ptr(total_dev)))

    # ---- portfolio
    @staticmethod
    def _checked_flat(flat):
        """The arrays of a FlatPortfolio with the dtypes / lengths the C ABI reads (raw pointers cross it): arrays of the
        right dtype and layout pass through untouched, anything else (an int64 node array, a strided view) is converted;
        wrong lengths raise."""
        def arr(a, dtype, n, name):
            if a is None:
                return None
            b = np.ascontiguousarray(a, dtype=dtype)
            if b.ndim != 1 or b.shape[0] != n:
                raise LibError(f"FlatPortfolio.{name}: expected {n} entries of {np.dtype(dtype).name}, got shape {np.shape(a)}")
            return b
        U, T, N, G, P, K = flat.n_units, flat.n_terms, flat.n_trades, flat.n_groups, flat.n_pairs, flat.n_comp
        return dict(unit_offsets=arr(flat.unit_offsets, np.int64, U + 1, "unit_offsets"), amt=arr(flat.amt, np.float64, T, "amt"),
                    weight=arr(flat.weight, np.float64, T * P, "weight"), node=arr(flat.node, np.int32, T * P, "node"),
                    comp_weight=arr(flat.comp_weight, np.float64, N * K, "comp_weight"),
                    group_offsets=arr(flat.group_offsets, np.int64, G + 1, "group_offsets"),
                    group_units=arr(flat.group_units, np.int32, G * K, "group_units"),
                    out_index=arr(flat.out_index, np.int64, N, "out_index"),
                    unit_weight=arr(flat.unit_weight, np.float64, U, "unit_weight"))

    def portfolio_upload(self, flat):
        """flat: adrates_b200.flatten.FlatPortfolio (numpy or pinned torch-backed arrays)."""
        a = self._checked_flat(flat)
        self._ck(self._dll.cav_portfolio_upload(
            self._h, flat.n_units, flat.n_terms, _ptr(a["unit_offsets"]), flat.n_pairs, _ptr(a["amt"]),
            _ptr(a["weight"]), _ptr(a["node"]), flat.n_trades, flat.n_comp, _ptr(a["comp_weight"]), flat.n_groups,
            _ptr(a["group_offsets"]), _ptr(a["group_units"]), _ptr(a["out_index"]), _ptr(a["unit_weight"])))
        self._n_trades = flat.n_trades
        self._flat_in_flight = (flat, a)     # async upload: the host arrays must outlive the copies
        tp = getattr(flat, "tile_plan", None)
        if tp is not None:
            self.portfolio_set_tiles(tp)

    def portfolio_set_tiles(self, tp):
        """tp: adrates_b200.tiles.TilePlan (all units covered)."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        mask = None if getattr(tp, "tile_mask", None) is None else np.ascontiguousarray(tp.tile_mask, dtype=np.uint32)
        arrs = [i32(tp.tile_units), i32(tp.tile_kstart), i32(tp.tile_kcount), i32(tp.k_row), i32(tp.k_pos),
                i32(tp.k_coef), i32(tp.pairs)]
        opt = lambda name: None if getattr(tp, name, None) is None else i32(getattr(tp, name))  # noqa: E731
        pos2, coef2, perm = opt("k_pos2"), opt("k_coef2"), opt("perm")
        self._tiles_in_flight = (arrs, pos2, coef2, perm, mask)     # async upload: outlive the copies
        self._ck(self._dll.cav_portfolio_set_tiles(self._h, tp.n_tiles, tp.tile_size, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                                   arrs[3].shape[0], _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]),
                                                   _ptr(pos2), _ptr(coef2), arrs[6].shape[0] // 2, _ptr(arrs[6]),
                                                   _ptr(mask), _ptr(perm)))

    # ---- multi-GPU totals (one process per GPU)
    def comm_local_handle(self) -> bytes:
        buf = C.create_string_buffer(COMM_HANDLE_BYTES)
        self._ck(self._dll.cav_comm_local_handle(self._h, buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, handles):
        """handles: the comm_local_handle() blobs of all ranks in rank order."""
        blob = b"".join(handles)
        if len(blob) != world * COMM_HANDLE_BYTES:
            raise LibError(f"comm_init: expected {world} handles of {COMM_HANDLE_BYTES} bytes")
        self._ck(self._dll.cav_comm_init(self._h, int(rank), int(world), blob))
        self._comm_world = world

    def comm_status(self):
        r, w, lost = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._dll.cav_comm_status(self._h, C.byref(r), C.byref(w), C.byref(lost)))
        return int(r.value), int(w.value), bool(lost.value)

    # ---- device-side flattening of array books
    def book_set_holidays(self, table=None):
        """cav_book_set_holidays: the non-business-day bitmap (adrates_b200.holidays table) the device flattener walks in
        Calendar.adjust for cal_type >= 3; None clears it.  The same table is uploaded once per context."""
        if table is getattr(self, "_hol_table", None):
            return
        if table is None:
            self._ck(self._dll.cav_book_set_holidays(self._h, None, 0, 0))
        else:
            from . import holidays
            words = table.words()
            self._ck(self._dll.cav_book_set_holidays(self._h, _ptr(words), holidays.BASE, holidays.N_DAYS))
        self._hol_table = table

    def book_from_arrays(self, conv: "BookConv", effective, termination=None, tenor=None, tenor_unit: int = TENOR_YEARS,
                         fixed_sign=None, coupon=None, notional=None, spread=None, tiles: bool = True):
        """cav_book_from_arrays: per-trade arrays -> flat book + tile plan in HBM (no flat arrays cross PCIe)."""
        def arr(a, dtype):
            return None if a is None else np.ascontiguousarray(a, dtype=dtype)
        # narrow inputs travel as they are: int32 day serials (both date arrays) and int8 sides (flags of the C call)
        is_dt = lambda a, dt: isinstance(a, np.ndarray) and a.dtype == dt      # noqa: E731
        d32 = is_dt(effective, np.int32) and (termination is None or is_dt(termination, np.int32))
        s8 = is_dt(fixed_sign, np.int8)
        date_t = np.int32 if d32 else np.int64
        eff = arr(effective, date_t)
        n = eff.shape[0]
        term, ten = arr(termination, date_t), arr(tenor, np.int32)
        sg = arr(fixed_sign, np.int8 if s8 else np.float64)
        cp, no, sp = (arr(a, np.float64) for a in (coupon, notional, spread))
        for name, a in (("termination", term), ("tenor", ten), ("fixed_sign", sg), ("coupon", cp), ("notional", no), ("spread", sp)):
            if a is not None and a.shape != (n,):
                raise LibError(f"book_from_arrays: {name} must have one entry per trade")
        self._ck(self._dll.cav_book_from_arrays(self._h, C.addressof(conv), n, _ptr(eff), _ptr(term), _ptr(ten), int(tenor_unit),
                                                _ptr(sg), _ptr(cp), _ptr(no), _ptr(sp),
                                                (BOOK_TILES if tiles else 0) | (BOOK_DATES_I32 if d32 else 0) | (BOOK_SIGN_I8 if s8 else 0)))
        self._n_trades = n

    def book_info(self) -> dict:
        out = np.zeros(10, dtype=np.int64)
        self._ck(self._dll.cav_book_info(self._h, _ptr(out)))
        keys = ("n_units", "n_terms", "n_trades", "n_groups", "n_pairs", "n_comp", "n_tiles", "n_krows", "n_pair_rows", "built")
        return {k: int(v) for k, v in zip(keys, out)}

    def book_read(self):
        """The flat arrays of the portfolio on the device as a FlatPortfolio (tests / inspection)."""
        from .flatten import FlatPortfolio
        i = self.book_info()
        U, T, N, G, P, K = (i[k] for k in ("n_units", "n_terms", "n_trades", "n_groups", "n_pairs", "n_comp"))
        off, amt, w = np.empty(U + 1, dtype=np.int64), np.empty(T), np.empty(T * P)
        node, cw = np.empty(T * P, dtype=np.int32), np.empty(N * K)
        go, gu = np.empty(G + 1, dtype=np.int64), np.empty(G * K, dtype=np.int32)
        oi, uw = np.empty(N, dtype=np.int64), np.empty(U)
        self._ck(self._dll.cav_book_read(self._h, _ptr(off), _ptr(amt), _ptr(w), _ptr(node), _ptr(cw), _ptr(go), _ptr(gu),
                                         _ptr(oi), _ptr(uw)))
        return FlatPortfolio(U, T, off, P, amt, w, node, N, K, cw, G, go, gu, oi, uw)

    def book_read_tiles(self) -> dict:
        i = self.book_info()
        nt, nk, npr = i["n_tiles"], i["n_krows"], i["n_pair_rows"]
        i32 = lambda n: np.empty(n, dtype=np.int32)  # noqa: E731
        out = dict(tile_units=i32(nt * 16), tile_kstart=i32(nt), tile_kcount=i32(nt), tile_npos=i32(nt),
                   tile_mask=np.empty(nt, dtype=np.uint32), k_row=i32(nk), k_desc=i32(nk), pairs=i32(2 * npr), perm=i32(32),
                   class_begin=i32(7))
        self._ck(self._dll.cav_book_read_tiles(self._h, *(_ptr(out[k]) for k in (
            "tile_units", "tile_kstart", "tile_kcount", "tile_npos", "tile_mask", "k_row", "k_desc", "pairs", "perm", "class_begin"))))
        out["n_tiles"] = nt
        return out

    def portfolio_value(self, mask: int, pv_dev=None, delta_dev=None, gamma_dev=None, agg_dev=None):
        self._ck(self._dll.cav_portfolio_value(self._h, mask, _ptr(pv_dev), _ptr(delta_dev), _ptr(gamma_dev),
                                               _ptr(agg_dev)))

    def portfolio_value_host(self, mask: int, pv_dev=None, delta_dev=None, gamma_dev=None, agg_host=None):
        if agg_host is None:
            agg_host = np.empty(NOUT)
        self._ck(self._dll.cav_portfolio_value_host(self._h, mask, _ptr(pv_dev), _ptr(delta_dev), _ptr(gamma_dev),
                                                    _ptr(agg_host)))
        return agg_host

    def portfolio_delta_gemm(self, pv_dev, delta_dev):
        """PV + delta with the chain rule as a DMMA GEMM; returns (gemm_ms, gemm_flops)."""
        ms, fl = C.c_float(), C.c_double()
        self._ck(self._dll.cav_portfolio_delta_gemm(self._h, _ptr(pv_dev), _ptr(delta_dev), C.byref(ms), C.byref(fl)))
        return float(ms.value), float(fl.value)

    def scenarios(self, shocked_rates, pnl_dev):
        r = _f64(shocked_rates)
        self._ck(self._dll.cav_scenarios(self._h, _ptr(r), r.shape[0], _ptr(pnl_dev)))

    def scenarios_info(self) -> dict:
        out = np.zeros(4, dtype=np.int64)
        self._ck(self._dll.cav_scenarios_info(self._h, _ptr(out)))
        return dict(zip(("queries", "chains", "chain_terms", "chain_kernel"), (int(x) for x in out)))


_default = {}


def lib(device: int = 0) -> Context:
    """Process-wide default context per device."""
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
