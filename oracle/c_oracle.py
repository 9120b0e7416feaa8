"""ctypes wrapper of oracle/liboracle.so (TEST / BASELINE INFRASTRUCTURE, see cavour_oracle.c)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(path)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ois_batch(tables, method, sched_tables, trades, want=7, dense=True, n_threads=0):
    """tables = (x[G], d[G], J[G,R], C[G,R,R]); sched_tables = dict(fo, f_pay_t, f_alpha, lo, l_start_t, l_end_t,
    l_pay_t, l_alpha); trades = dict(sched, coupon, notional, spread, fixed_sign).  Returns pv, delta, gamma."""
    x, d, J, Cc = [np.ascontiguousarray(a, dtype=np.float64) for a in tables]
    G, R = J.shape
    st = {k: np.ascontiguousarray(v, dtype=np.int64 if k in ("fo", "lo") else np.float64) for k, v in sched_tables.items()}
    sched = np.ascontiguousarray(trades["sched"], dtype=np.int32)
    tr = {k: np.ascontiguousarray(trades[k], dtype=np.float64) for k in ("coupon", "notional", "spread", "fixed_sign")}
    n = sched.shape[0]
    pv = np.zeros(n) if want & 1 else None
    dl = np.zeros((n, R)) if want & 2 else None
    gm = np.zeros((n, R, R)) if want & 4 else None
    lib().oracle_ois_batch(
        C.c_int(G), C.c_int(R), C.c_int(method), _p(x), _p(d), _p(J), _p(Cc),
        _p(st["fo"]), _p(st["f_pay_t"]), _p(st["f_alpha"]),
        _p(st["lo"]), _p(st["l_start_t"]), _p(st["l_end_t"]), _p(st["l_pay_t"]), _p(st["l_alpha"]),
        C.c_int64(n), _p(sched), _p(tr["coupon"]), _p(tr["notional"]), _p(tr["spread"]), _p(tr["fixed_sign"]),
        C.c_int(want), C.c_int(1 if dense else 0), C.c_int(n_threads), _p(pv), _p(dl), _p(gm))
    return pv, dl, gm


def ois_batch_units(tables, method, sched_tables, trades, want=7, n_threads=0, out=None):
    """The same per-trade results as ois_batch(dense=False) computed WITH the unit factorisation of the GPU path: both legs
    of every distinct schedule are valued once on unit trades (sparse chain rule), then every trade is the weighted sum of
    two unit rows (oracle_expand_units).  Returns (pv, delta, gamma, seconds_units, seconds_expand)."""
    import time
    x, d, J, Cc = [np.ascontiguousarray(a, dtype=np.float64) for a in tables]
    G, R = J.shape
    st = {k: np.ascontiguousarray(v, dtype=np.int64 if k in ("fo", "lo") else np.float64) for k, v in sched_tables.items()}
    sched = np.ascontiguousarray(trades["sched"], dtype=np.int32)
    used, inv = np.unique(sched, return_inverse=True)
    S = used.shape[0]
    ones, zeros = np.ones(S), np.zeros(S)
    used32 = np.ascontiguousarray(used, dtype=np.int32)
    t0 = time.perf_counter()
    u_pv, u_dl, u_gm = np.zeros(2 * S), np.zeros((2 * S, R)), np.zeros((2 * S, R, R))
    for leg, off, sign in ((1, 0, ones), (2, S, -ones)):       # annuity of N c = 1 receive-fixed; floating leg of N = 1 received
        lib().oracle_ois_batch_legs(
            C.c_int(G), C.c_int(R), C.c_int(method), _p(x), _p(d), _p(J), _p(Cc),
            _p(st["fo"]), _p(st["f_pay_t"]), _p(st["f_alpha"]),
            _p(st["lo"]), _p(st["l_start_t"]), _p(st["l_end_t"]), _p(st["l_pay_t"]), _p(st["l_alpha"]),
            C.c_int64(S), _p(used32), _p(ones), _p(ones), _p(zeros), _p(sign),
            C.c_int(want | 1), C.c_int(0), C.c_int(n_threads), C.c_int(leg), _p(u_pv[off:]), _p(u_dl[off:]), _p(u_gm[off:]))
    t1 = time.perf_counter()
    n = sched.shape[0]
    fs = np.ascontiguousarray(trades["fixed_sign"], dtype=np.float64)
    no = np.ascontiguousarray(trades["notional"], dtype=np.float64)
    wa = fs * no * np.ascontiguousarray(trades["coupon"], dtype=np.float64)
    wf = -fs * no
    ua = np.ascontiguousarray(inv, dtype=np.int32)
    uf = np.ascontiguousarray(inv + S, dtype=np.int32)
    if out is None:
        out = (np.empty(n) if want & 1 else None, np.empty((n, R)) if want & 2 else None, np.empty((n, R, R)) if want & 4 else None)
    pv, dl, gm = out
    lib().oracle_expand_units(C.c_int64(n), C.c_int(R), _p(ua), _p(uf), _p(wa), _p(wf), _p(u_pv), _p(u_dl), _p(u_gm),
                              C.c_int(n_threads), _p(pv), _p(dl), _p(gm))
    t2 = time.perf_counter()
    return pv, dl, gm, t1 - t0, t2 - t1
