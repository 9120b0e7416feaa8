"""The two tile kernels of the Greeks stage - k_units_mma (every warp walks all phases) and k_units_mma_ws (warp-specialised:
front warps build coefficient tiles ahead of the mma warps, mbarrier hand-over) - evaluate the same arithmetic in the same
order: per-trade PV / delta / gamma rows must be bit-identical, the portfolio totals (different partial-sum orders) equal to
1e-12 of the absolute mass, and the specialised kernel must match the C oracle on a sample.  The switch CAV_UNITS_WS is read
once per process, so each variant runs in its own process (tools/ws_check.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(ws, n, oracle=False, prepass=0):
    env = dict(os.environ, CAV_UNITS_WS=str(ws), CAV_TERM_PREPASS=str(prepass))
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ws_check.py"), str(n)] + (["--oracle"] if oracle else [])
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=540)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("WSCHECK ")][-1]
    return json.loads(line[len("WSCHECK "):])


def test_warp_specialised_tile_kernel_matches_single_role_kernel_and_oracle():
    n = 40_000          # ~2 800 tiles over four size classes: every persistent CTA of the specialised kernel takes several tiles
    a = _run(0, n)
    b = _run(1, n, oracle=True)
    assert a["tiles"] == b["tiles"] > 1000
    assert (a["pv"], a["delta"], a["gamma"]) == (b["pv"], b["delta"], b["gamma"])
    agg_a, agg_b = np.array(a["agg"]), np.array(b["agg"])
    mass = b["agg_abs"]
    assert abs(agg_a[0] - agg_b[0]) <= 1e-12 * mass[0]
    assert np.max(np.abs(agg_a[1:33] - agg_b[1:33])) <= 1e-12 * mass[1]
    assert np.max(np.abs(agg_a[33:] - agg_b[33:])) <= 1e-12 * mass[2]
    assert b["oracle_err"] < 1e-10


def test_term_scalar_prepass_is_bit_identical_in_both_tile_kernels():
    """k_term_scalars (p = amt * DF of every term in one streaming pass, CAV_TERM_PREPASS=1) against the tile kernels
    computing p in place: the same expression, so PV / delta / gamma rows must not move by a bit - single-role and
    warp-specialised kernel alike - and the totals agree to 1e-12 of the absolute mass."""
    n = 40_000
    for ws in (0, 1):
        a = _run(ws, n)
        b = _run(ws, n, oracle=(ws == 1), prepass=1)
        assert (a["pv"], a["delta"], a["gamma"]) == (b["pv"], b["delta"], b["gamma"]), ws
        assert np.max(np.abs(np.array(a["agg"]) - np.array(b["agg"]))[33:]) <= 1e-12 * b["agg_abs"][2]
        if ws == 1:
            assert b["oracle_err"] < 1e-10


def test_parity_suites_pass_with_the_warp_specialised_kernel_forced():
    """Small and ragged books (padding units, single-tile classes, bonds, device-built plans) through k_units_mma_ws: the parity
    suites that pin the tiled Greeks to the reference engine's goldens, re-run in a process where CAV_UNITS_WS=1."""
    env = dict(os.environ, CAV_UNITS_WS="1", CAV_TERM_PREPASS="1")      # and the term-scalar pre-pass feeding it
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_gpu_bond_book.py"),
           os.path.join(ROOT, "tests", "test_gpu_book_device.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=560, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]
