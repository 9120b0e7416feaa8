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
