"""XccyCurve bootstrap on the device (cav_xccy_curve_scan / k_xccy_scan): discount factors, d DF / d spread (`_jac_basis`,
xccy_curve.py:594) and d2 DF / d spread^2 (`_hess_basis`, xccy_curve.py:603-606) from exact forward-mode tangents through the
scan of xccy_curve.py:954-1206.  Pinned by the reference's own curve (DFs and AD Jacobian in tests/golden/ref_xccy.json), by the
host scan on second-order forward-mode numbers, and - for batches of shocked basis curves - by re-bootstrapping each curve."""
import copy

import numpy as np
import pytest

from adrates_b200 import _native
from tests.conftest import load_golden
from tests.util_xccy import build_xccy_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xc_setup():
    g = load_golden("ref_xccy.json")
    m = build_xccy_model(g)
    return g, m.curves.GBP_USD_BASIS


def test_device_bootstrap_matches_reference_curve_and_jacobian(xc_setup):
    g, xc = xc_setup
    dfs, jac, hess = xc.device_tables(order=2)
    assert dfs.shape == (1, len(xc._times)) and jac.shape[1:] == xc._jac_basis.shape
    assert np.max(np.abs(dfs[0] - np.array(g["xccy_dfs"]))) < 1e-14
    J = np.array(g["xccy_jac_basis"])
    assert np.max(np.abs(jac[0] - J)) < 1e-12 * np.abs(J).max()
    assert np.max(np.abs(jac[0] - xc._jac_basis)) < 1e-13 * np.abs(J).max()
    H = xc._hess_basis                         # host scan on second-order forward-mode numbers (dual2)
    assert hess[0].shape == H.shape
    assert np.max(np.abs(hess[0] - H)) < 1e-11 * np.abs(H).max()
    assert np.array_equal(hess[0], np.swapaxes(hess[0], 1, 2)) or np.max(np.abs(hess[0] - np.swapaxes(hess[0], 1, 2))) < 1e-13 * np.abs(H).max()


def test_device_hessian_is_the_finite_difference_of_the_device_jacobian(xc_setup):
    _, xc = xc_setup
    nb = len(xc.basis_spreads)
    h = 1e-6
    base = np.array(xc.basis_spreads)
    bumps = np.concatenate([base + h * np.eye(nb), base - h * np.eye(nb)])
    _, jac, _ = xc.device_tables(spreads=bumps, order=1)
    fd = (jac[:nb] - jac[nb:]) / (2 * h)                   # [k][node][b] = d/ds_k of d DF_node / d s_b
    _, _, hess = xc.device_tables(order=2)
    scale = np.abs(hess[0]).max()
    assert np.max(np.abs(np.transpose(fd, (1, 0, 2)) - hess[0])) < 1e-6 * scale


def test_shocked_basis_curves_rebuild_in_one_launch(xc_setup):
    _, xc = xc_setup
    rng = np.random.Generator(np.random.PCG64(5))
    nb = len(xc.basis_spreads)
    S = 7
    spreads = np.array(xc.basis_spreads)[None, :] + rng.normal(0.0, 5e-4, (S, nb))
    ctx = _native.Context(0)
    dfs, jac, _ = xc.device_tables(ctx=ctx, spreads=spreads, order=1)
    ctx.close()
    for s in range(S):
        c = copy.copy(xc)
        c.basis_spreads = list(spreads[s])
        c._bootstrap()
        assert np.max(np.abs(dfs[s] - c._dfs)) < 1e-14
        assert np.max(np.abs(jac[s] - c._jac_basis)) < 1e-13 * np.abs(c._jac_basis).max()


def test_scan_rejects_malformed_plans(xc_setup):
    _, xc = xc_setup
    pl = xc.scan_plan()
    ctx = _native.Context(0)
    bad = pl["swap"].copy()
    bad[0] = 99
    with pytest.raises(Exception):
        ctx.xccy_curve_scan(pl["time"], bad, pl["flags"], pl["sens"], pl["base"], pl["df_ois"], pl["pv_dom"], xc._spot_fx,
                            xc.basis_spreads, order=0)
    with pytest.raises(Exception):
        ctx.xccy_curve_scan(pl["time"][::-1].copy(), pl["swap"], pl["flags"], pl["sens"], pl["base"], pl["df_ois"], pl["pv_dom"],
                            xc._spot_fx, xc.basis_spreads, order=0)
    ctx.close()
