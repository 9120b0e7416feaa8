"""jax.ffi binding (adrates_b200/jax_binding.py + csrc/jax_ffi_adapter.cc): jax.grad / jax.hessian of the portfolio PV are
the library's ladder and gamma matrix.  Runs where `import jax` works with a CUDA backend; JAX is not installed in the image
this repository is developed in, where everything but the import-free checks skips."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_sources_are_present_and_bind_the_c_abi():
    """Import-free: the adapter names only entry points include/adrates_b200.h declares, and the Python side registers the
    symbol the adapter defines."""
    src = open(os.path.join(ROOT, "adrates_b200", "csrc", "jax_ffi_adapter.cc")).read()
    hdr = open(os.path.join(ROOT, "include", "adrates_b200.h")).read()
    for fn in ("cav_set_stream", "cav_curve_rebuild_dev", "cav_portfolio_value", "cav_last_error"):
        assert fn + "(" in src and fn + "(" in hdr, fn
    assert "XLA_FFI_DEFINE_HANDLER_SYMBOL(CavPortfolioTotals" in src
    py = open(os.path.join(ROOT, "adrates_b200", "jax_binding.py")).read()
    assert "lib.CavPortfolioTotals" in py and 'TARGET = "cav_portfolio_totals"' in py
    import adrates_b200.jax_binding as jb            # importable without jax
    from adrates_b200.build import jax_ffi_include
    try:
        import jax  # noqa: F401
    except ImportError:
        assert jax_ffi_include() is None
        from adrates_b200.error import LibError
        with pytest.raises(LibError, match="needs jax"):
            jb.portfolio_pv_function(None, 32)


@pytest.mark.gpu
def test_grad_and_hessian_compose_through_the_custom_call():
    jax = pytest.importorskip("jax")
    import jax.numpy as jnp
    jax.config.update("jax_enable_x64", True)
    from adrates_b200 import _native
    from adrates_b200.jax_binding import portfolio_pv_function
    from adrates_b200.market_data import readme_gbp_curve
    from adrates_b200.synthetic import make_array_book
    curve = readme_gbp_curve()
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    make_array_book(curve, 4000, seed=5).upload(ctx)
    agg = ctx.portfolio_value_host(7)
    R = len(curve.swap_rates)
    pv, grad_pv, totals = portfolio_pv_function(ctx, R)
    r = jnp.asarray(curve.swap_rates)
    assert abs(float(pv(r)) - agg[0]) <= 1e-12 * abs(agg[0])
    g = np.asarray(jax.grad(pv)(r))
    assert np.allclose(g * 1e-4, agg[1:1 + R], rtol=1e-12, atol=1e-9)
    H = np.asarray(jax.hessian(pv)(r))
    assert np.allclose(H * 1e-8, agg[33:].reshape(32, 32)[:R, :R], rtol=1e-12, atol=1e-12)
    # a composition the reference's users write: risk of a function of the PV, under jit
    f = jax.jit(lambda x: jnp.tanh(pv(x) * 1e-9))
    h = 1e-6
    e = np.zeros(R)
    e[20] = h
    fd = (float(f(r + e)) - float(f(r - e))) / (2 * h)
    assert abs(float(jax.grad(f)(r)[20]) - fd) <= 1e-5 * max(abs(fd), 1e-12)
