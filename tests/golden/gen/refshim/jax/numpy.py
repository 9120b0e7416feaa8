"""jax.numpy subset on torch float64 (see package docstring)."""
import builtins as _bi
import numpy as _np
import torch as _torch

ndarray = _torch.Tensor
int32 = _torch.int32
int64 = _torch.int64
float64 = _torch.float64
bool_ = _torch.bool
inf = float("inf")
nan = float("nan")
pi = _np.pi


def _is_t(x):
    return isinstance(x, _torch.Tensor)


def array(x, dtype=None):
    if _is_t(x):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, _np.ndarray):
        t = _torch.from_numpy(_np.ascontiguousarray(x))
        if dtype is None and t.dtype in (_torch.float32, _torch.float16):
            t = t.to(_torch.float64)
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)):
        if len(x) and _bi.any(_is_t(e) for e in x):
            t = _torch.stack([array(e) if not _is_t(e) else e for e in x])
            t = t.to(_torch.float64) if (dtype is None and not t.dtype.is_floating_point and False) else t
            return t if dtype is None else t.to(dtype)
        a = _np.array(x)
        if dtype is None and a.dtype.kind == "f":
            a = a.astype(_np.float64)
        if a.dtype == object:
            raise TypeError("ragged input to jnp.array")
        return array(a, dtype)
    if isinstance(x, bool):
        return _torch.tensor(x)
    if isinstance(x, int):
        return _torch.tensor(x, dtype=dtype or _torch.int64)
    return _torch.tensor(float(x), dtype=dtype or _torch.float64)


asarray = array


def atleast_1d(x):
    x = array(x)
    return x.reshape(1) if x.ndim == 0 else x


def squeeze(x, axis=None):
    x = array(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def where(c, a, b):
    c = array(c)
    if not _is_t(a) and not _is_t(b):
        a = array(a)
    return _torch.where(c, a, b)


def log(x):
    return _torch.log(array(x))


def exp(x):
    return _torch.exp(array(x))


def abs(x):
    return _torch.abs(array(x))


def power(a, b):
    return _torch.pow(array(a), b)


def append(a, v):
    a = array(a).reshape(-1)
    v = atleast_1d(v).reshape(-1)
    if a.numel() == 0:
        return v.clone()
    return _torch.cat([a, v.to(a.dtype)])


def sum(x, axis=None):
    x = array(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def dot(a, b):
    a, b = array(a), array(b)
    if a.ndim == 1 and b.ndim == 1:
        return _torch.dot(a, b)
    return a @ b


def concatenate(xs, axis=0):
    return _torch.cat([array(x) for x in xs], dim=axis)


def full_like(a, v):
    a = array(a)
    if _is_t(v):
        return _torch.ones_like(a) * v
    return _torch.full_like(a, v)


def zeros(shape, dtype=None):
    return _torch.zeros(shape, dtype=dtype or _torch.float64)


def ones(shape, dtype=None):
    return _torch.ones(shape, dtype=dtype or _torch.float64)


def empty(shape, dtype=None):
    return _torch.zeros(shape, dtype=dtype or _torch.float64)


def empty_like(a):
    return _torch.zeros_like(array(a))


def ndim(x):
    return array(x).ndim if not isinstance(x, (int, float)) else 0


def maximum(a, b):
    a = array(a)
    if not _is_t(b):
        b = _torch.tensor(float(b), dtype=a.dtype)
    return _torch.maximum(a, b)


def minimum(a, b):
    a = array(a)
    if not _is_t(b):
        b = _torch.tensor(float(b), dtype=a.dtype)
    return _torch.minimum(a, b)


def clip(x, lo, hi):
    return _torch.clamp(array(x), lo, hi)


def broadcast_to(x, shape):
    return array(x).broadcast_to(tuple(shape))


def arange(*a, **k):
    return _torch.arange(*a, **k)


def searchsorted(a, v, side="left"):
    a = array(a).detach()
    v = array(v).detach().to(a.dtype)
    return _torch.searchsorted(a, v, right=(side == "right"))


def interp(x, xp, fp, left=None, right=None):
    """jnp.interp semantics (jax/_src/numpy/lax_numpy.py::_interp):
    i = clip(searchsorted(xp, x, side='right'), 1, len(xp)-1); linear between
    (i-1, i); zero-width cells return fp[i-1]; outside the range clamp to the ends."""
    x = array(x)
    xp = array(xp)
    fp = array(fp)
    xs = x.to(xp.dtype)
    i = _torch.clamp(_torch.searchsorted(xp.detach(), xs.detach(), right=True), 1, xp.shape[0] - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = xs - xp[i - 1]
    eps = float(_np.spacing(_np.finfo(_np.float64).eps))
    dx0 = _torch.abs(dx) <= eps
    f = _torch.where(dx0, fp[i - 1], fp[i - 1] + (delta / _torch.where(dx0, _torch.ones_like(dx), dx)) * df)
    f = _torch.where(xs < xp[0], fp[0] if left is None else left, f)
    f = _torch.where(xs > xp[-1], fp[-1] if right is None else right, f)
    return f


def min(x, axis=None):
    x = array(x)
    return x.min() if axis is None else x.min(dim=axis).values


def max(x, axis=None):
    x = array(x)
    return x.max() if axis is None else x.max(dim=axis).values


def argmin(x):
    # first index on ties, as jnp.argmin / np.argmin
    x = array(x).detach()
    m = x.min()
    return _torch.nonzero(x == m)[0, 0]


def any(x):
    return bool(array(x).any())


def isnan(x):
    return _torch.isnan(array(x))


def isinf(x):
    return _torch.isinf(array(x))


def unique(x):
    return _torch.unique(array(x))


def diff(x):
    x = array(x)
    return x[1:] - x[:-1]


def transpose(x, axes=None):
    x = array(x)
    return x.T if axes is None else x.permute(*axes)


def einsum(expr, *ops):
    return _torch.einsum(expr, *[array(o) for o in ops])


def cumsum(x, axis=0):
    return _torch.cumsum(array(x), dim=axis)


def sqrt(x):
    return _torch.sqrt(array(x))


def stack(xs, axis=0):
    return _torch.stack([array(x) for x in xs], dim=axis)
