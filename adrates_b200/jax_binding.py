"""jax.ffi binding of the valuation: portfolio PV as a JAX function of the curve's par rates whose derivatives are the
library's own Greeks.

The reference composes `jax.grad` / `jax.hessian` around its leg pricers and chains the result through the bootstrap tables
(cavour/market/position/engine.py:2551-2568, 2909-2926).  Here the whole chain is one custom call:

    totals = ffi_call("cav_portfolio_totals")(rates)        # [PV, ladder(32) per bp, gamma(32x32) per bp^2], device resident
    pv(rates)        custom_jvp:  tangent = <grad_pv(rates), t>
    grad_pv(rates)   custom_jvp:  primal = 1e4 * ladder,  tangent = (1e8 * gamma) @ t

so `jax.grad(pv)`, `jax.jacfwd(jax.grad(pv))` and `jax.hessian(pv)` return the ladder and the gamma matrix (in per-unit-rate
units) without JAX ever differentiating through the kernels; third derivatives are not defined (the gamma matrix is treated as
constant).  `jax` is imported lazily: the module can be imported without it, `portfolio_pv_function` raises LibError.

Needs `jax` with a CUDA backend and the adapter library built by `adrates_b200.build.build_jax_ffi()`.  JAX is not part of the
image this repository is developed in (tests/test_jax_binding.py skips there); INTEGRATION.md section 3.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native
from .error import LibError

_HERE = os.path.dirname(os.path.abspath(__file__))
ADAPTER_PATH = os.path.join(_HERE, "libadrates_b200_jax.so")
TARGET = "cav_portfolio_totals"
_registered = False


def _register():
    global _registered
    if _registered:
        return
    try:
        import jax
    except ImportError as ex:
        raise LibError(f"adrates_b200.jax_binding needs jax ({ex})")
    if not os.path.exists(ADAPTER_PATH):
        from .build import build_jax_ffi
        if build_jax_ffi() is None:
            raise LibError("jaxlib's FFI headers (xla/ffi/api/ffi.h) were not found: cannot build the jax.ffi adapter")
    _native.load_dll()                                   # the adapter links against libadrates_b200.so
    lib = C.CDLL(ADAPTER_PATH)
    ffi = jax.ffi if hasattr(jax, "ffi") else jax.extend.ffi
    ffi.register_ffi_target(TARGET, ffi.pycapsule(lib.CavPortfolioTotals), platform="CUDA")
    _registered = True


def portfolio_pv_function(ctx: "_native.Context", n_rates: int):
    """(pv, grad_pv, totals): JAX functions of the par-rate vector f64[n_rates] for the portfolio held by `ctx` (curve built
    from its plan with order=2, portfolio uploaded or device-flattened).  pv -> scalar, grad_pv -> f64[n_rates] = dPV/d rates,
    totals -> the raw f64[1057] block.  Enable x64 (`jax.config.update("jax_enable_x64", True)`)."""
    _register()
    import jax
    import jax.numpy as jnp
    ffi = jax.ffi if hasattr(jax, "ffi") else jax.extend.ffi
    handle = np.int64(C.cast(ctx._h, C.c_void_p).value)
    mask = np.int32(_native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA)
    out_type = jax.ShapeDtypeStruct((_native.NOUT,), jnp.float64)
    R = int(n_rates)

    def totals(rates):
        return ffi.ffi_call(TARGET, out_type)(jnp.asarray(rates, dtype=jnp.float64), ctx=handle, mask=mask)

    @jax.custom_jvp
    def grad_pv(rates):
        return totals(rates)[1:1 + R] * 1e4                      # ladder is per bp: 1e-4 dPV/dr

    @grad_pv.defjvp
    def _grad_pv_jvp(primals, tangents):
        (rates,), (t,) = primals, tangents
        tot = totals(rates)
        gamma = tot[33:].reshape(32, 32)[:R, :R] * 1e8           # gamma is per bp^2: 1e-8 d2PV/dr2
        return tot[1:1 + R] * 1e4, gamma @ t

    @jax.custom_jvp
    def pv(rates):
        return totals(rates)[0]

    @pv.defjvp
    def _pv_jvp(primals, tangents):
        (rates,), (t,) = primals, tangents
        return totals(rates)[0], jnp.dot(grad_pv(rates), t)

    return pv, grad_pv, totals
