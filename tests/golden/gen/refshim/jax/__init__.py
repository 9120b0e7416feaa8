"""Torch-backed stand-in for the tiny slice of JAX the reference uses.

TEST INFRASTRUCTURE ONLY.  JAX/jaxlib are not installed in the build container, so
the reference (/root/reference, pure Python + JAX) cannot be imported as-is.  This
package lets the UNMODIFIED reference modules run here by mapping the ~40 jax/jnp
entry points they call onto torch (float64) and torch.func (grad/hessian/jacrev).
It is used by tests/golden/gen/make_golden.py to produce the committed golden
vectors; nothing in the product path or in the GPU tests imports it.

Semantics that matter for parity are written out explicitly:
  * jnp.interp   -> searchsorted(side='right'), clip to [1, n-1], linear, clamp ends
  * lax.scan     -> python loop, stacked outputs
  * x.at[i].set  -> clone + index assignment (functional update)
"""
import sys
import types
import functools
import numpy as _np
import torch as _torch

_torch.set_default_dtype(_torch.float64)


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def jit(fun=None, *, static_argnums=None, static_argnames=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _t(x):
    if isinstance(x, _torch.Tensor):
        return x
    return numpy.array(x)


def grad(f, argnums=0):
    def g(*args):
        args = list(args)
        args[argnums] = _t(args[argnums])
        return _torch.func.grad(lambda a: _scalar(f(*args[:argnums], a, *args[argnums + 1:])))(args[argnums])
    return g


def _scalar(v):
    v = _t(v)
    return v.reshape(())


def jacrev(f, argnums=0):
    def g(*args):
        args = list(args)
        args[argnums] = _t(args[argnums])
        return _torch.func.jacrev(lambda a: f(*args[:argnums], a, *args[argnums + 1:]))(args[argnums])
    return g


def jacfwd(f, argnums=0):
    def g(*args):
        args = list(args)
        args[argnums] = _t(args[argnums])
        return _torch.func.jacfwd(lambda a: f(*args[:argnums], a, *args[argnums + 1:]))(args[argnums])
    return g


def hessian(f, argnums=0):
    # jax.hessian = jacfwd(jacrev(f)), same composition here.
    def g(*args):
        args = list(args)
        args[argnums] = _t(args[argnums])
        inner = _torch.func.jacrev(lambda a: f(*args[:argnums], a, *args[argnums + 1:]))
        return _torch.func.jacfwd(inner)(args[argnums])
    return g


def linearize(f, *primals):
    raise NotImplementedError("linearize is not needed for the golden generator")


def vmap(f, in_axes=0, out_axes=0):
    def g(x, *rest):
        x = _t(x)
        outs = [f(x[i], *rest) for i in range(x.shape[0])]
        return _torch.stack([_t(o) for o in outs])
    return g


# ----------------------------------------------------------------------------------
# x.at[i].set(v) / .add(v)
class _At:
    def __init__(self, t):
        self._t = t

    def __getitem__(self, idx):
        return _AtIdx(self._t, idx)


class _AtIdx:
    def __init__(self, t, idx):
        self._t, self._idx = t, idx

    def set(self, v):
        out = self._t.clone()
        out[self._idx] = v
        return out

    def add(self, v):
        out = self._t.clone()
        out[self._idx] = out[self._idx] + v
        return out


_torch.Tensor.at = property(lambda self: _At(self))

from . import numpy  # noqa: E402
from . import lax    # noqa: E402
