"""jax.lax subset (scan as a python loop, select as where)."""
import torch as _torch
from . import numpy as jnp


def scan(f, init, xs, length=None):
    if isinstance(xs, (tuple, list)):
        n = xs[0].shape[0]
        get = lambda i: tuple(x[i] for x in xs)
    else:
        n = xs.shape[0]
        get = lambda i: xs[i]
    carry = init
    ys = []
    for i in range(n):
        carry, y = f(carry, get(i))
        ys.append(y)
    if ys and ys[0] is not None:
        if isinstance(ys[0], (tuple, list)):
            ys = tuple(_torch.stack([jnp.array(y[k]) for y in ys]) for k in range(len(ys[0])))
        else:
            ys = _torch.stack([jnp.array(y) for y in ys])
    return carry, ys


def select(pred, on_true, on_false):
    return _torch.where(jnp.array(pred), jnp.array(on_true), jnp.array(on_false))


def cond(pred, tf, ff, *ops):
    return tf(*ops) if bool(pred) else ff(*ops)


def stop_gradient(x):
    return jnp.array(x).detach()
