"""TEST INFRASTRUCTURE: host build of adrates_b200/csrc/cav_book_core.h (tests/native/book_core_host.cpp) through ctypes.
Only tests import this; the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "book_core_host.cpp")
HDR = os.path.join(HERE, "..", "adrates_b200", "csrc", "cav_book_core.h")
LIB = os.path.join(HERE, "native", "libbookcore.so")
_dll = None


def lib():
    global _dll
    if _dll is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
            # -ffp-contract=off: the device code must not fuse either (the rules are compared bit for bit)
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
        _dll = C.CDLL(LIB)
    return _dll


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def conv9(value_dt, fixed_step, float_step, fixed_dc, float_dc, cal, bd, dg, eom=0):
    return i64([value_dt, fixed_step, float_step, fixed_dc, float_dc, cal, bd, dg, int(eom)])


def ymd(n):
    n = i64(n)
    d, m, y = (np.empty_like(n) for _ in range(3))
    lib().bkh_ymd(_p(n), C.c_int64(n.size), _p(d), _p(m), _p(y))
    return d, m, y


def add_months(n, mm, eom=False):
    n = i64(n)
    mm = i64(np.broadcast_to(mm, n.shape))
    out = np.empty_like(n)
    lib().bkh_add_months(_p(n), _p(mm), C.c_int64(n.size), int(eom), _p(out))
    return out


def add_tenor(n, c, years=True):
    n = i64(n)
    c = i64(np.broadcast_to(c, n.shape))
    out = np.empty_like(n)
    lib().bkh_add_tenor(_p(n), _p(c), C.c_int64(n.size), int(years), _p(out))
    return out


def set_holidays(words, base, ndays):
    """non-business-day bitmap used by the following calls with cal >= 3 (None clears)"""
    if words is None:
        lib().bkh_set_holidays(None, C.c_int64(0), C.c_int64(0))
    else:
        w = np.ascontiguousarray(words, dtype=np.uint32)
        lib().bkh_set_holidays(_p(w), C.c_int64(int(base)), C.c_int64(int(ndays)))


def adjust(n, bd, cal=2):
    n = i64(n)
    out = np.empty_like(n)
    lib().bkh_adjust(_p(n), C.c_int64(n.size), int(bd), int(cal), _p(out))
    return out


def year_frac(n1, n2, dc):
    n1, n2 = i64(n1), i64(n2)
    out = np.empty(n1.shape)
    lib().bkh_year_frac(_p(n1), _p(n2), C.c_int64(n1.size), int(dc), _p(out))
    return out


def schedule(eff, term, step, cal, bd, dg, eom=False):
    """dates (int64 array) or a negative error code."""
    buf = np.empty(4096, dtype=np.int64)
    n = lib().bkh_schedule(C.c_int64(int(eff)), C.c_int64(int(term)), int(step), int(cal), int(bd), int(dg), int(eom), _p(buf), 4096)
    return n if n < 0 else buf[:n].copy()


def plan_queries(t, x, lzr):
    t = np.ascontiguousarray(t, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.empty(t.size, dtype=np.int32), np.empty(t.size, dtype=np.int32)
    wa, wb = np.empty(t.size), np.empty(t.size)
    lib().bkh_plan_queries(_p(t), C.c_int64(t.size), _p(x), int(x.size), int(lzr), _p(a), _p(b), _p(wa), _p(wb))
    return a, b, wa, wb


def flatten_classes(cv, eff, term, with_spread, x, lzr):
    """The unit arrays of S schedule classes in the device layout: (err, unit_offsets, amt, weight, node, has3, uid3, time)."""
    eff, term = i64(eff), i64(term)
    S = eff.size
    ws = None if with_spread is None else np.ascontiguousarray(with_spread, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    cnt3 = np.zeros(3 * S, dtype=np.int32)
    f = lib().bkh_flatten_classes
    err = f(_p(cv), C.c_int64(S), _p(eff), _p(term), _p(ws), _p(x), int(x.size), int(lzr), _p(cnt3), None, None, None, None, None)
    has3 = cnt3 > 0
    uid3 = np.cumsum(has3) - has3
    U = int(has3.sum())
    unit_cnt = np.zeros(U, dtype=np.int64)
    unit_cnt[uid3[has3]] = cnt3[has3]
    off = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(unit_cnt, out=off[1:])
    base3 = np.where(has3, off[:-1][np.minimum(uid3, max(U - 1, 0))] if U else 0, 0).astype(np.int64)
    T = int(off[-1])
    amt, time, weight = np.empty(T), np.empty(T), np.empty(2 * T)
    node = np.empty(2 * T, dtype=np.int32)
    if err == 0 and T:
        err = f(_p(cv), C.c_int64(S), _p(eff), _p(term), _p(ws), _p(x), int(x.size), int(lzr), _p(cnt3), _p(base3), _p(amt),
                _p(weight), _p(node), _p(time))
    return err, off, amt, weight, node, has3, uid3, time


def plan_tiles(unit_offsets, weight, node, G, support):
    off = i64(unit_offsets)
    U = off.size - 1
    T = int(off[-1])
    w = np.ascontiguousarray(weight, dtype=np.float64)
    nd = np.ascontiguousarray(node, dtype=np.int32)
    sup = np.ascontiguousarray(support, dtype=np.uint32)
    tu = np.empty(16 * max(U, 1), dtype=np.int32)
    ks, kc, npos = (np.empty(max(U, 1), dtype=np.int32) for _ in range(3))
    mask = np.empty(max(U, 1), dtype=np.uint32)
    k_row, k_desc = np.empty(5 * T + 8, dtype=np.int32), np.empty(5 * T + 8, dtype=np.int32)
    pairs, perm = np.empty(2 * G, dtype=np.int32), np.empty(32, dtype=np.int32)
    cnt = np.zeros(4, dtype=np.int64)
    lib().bkh_plan_tiles(C.c_int64(U), _p(off), _p(w), _p(nd), int(G), _p(sup), _p(tu), _p(ks), _p(kc), _p(npos), _p(mask),
                         _p(k_row), _p(k_desc), _p(pairs), _p(perm), _p(cnt))
    nt, nk, npair = int(cnt[0]), int(cnt[1]), int(cnt[2])
    return dict(n_tiles=nt, tile_units=tu[:16 * nt], tile_kstart=ks[:nt], tile_kcount=kc[:nt], tile_npos=npos[:nt],
                tile_mask=mask[:nt], k_row=k_row[:nk], k_desc=k_desc[:nk], pairs=pairs[:2 * npair], perm=perm,
                n_groups=int(cnt[3]))
