"""Argument type checks of the constructors, as the reference does them (cavour/utils/helpers.py:508-527, 618-636):
`check_argument_types(func, locals())` holds every annotated argument against its annotation with `isinstance` and raises
`LibError("Argument Type Error")` otherwise.  `float` admits int / float / numpy float64, a tuple annotation any of its members,
`List[...]` lists and arrays.  One deliberate difference: an argument whose declared default is None may be None (the reference
rejects its own defaults there, e.g. `InflationCurve(discount_curve: DiscountCurve = None)`).
"""
from __future__ import annotations

import inspect
import typing

import numpy as np

from .error import LibError

_NAMES = None


def _namespace() -> dict:
    """Names annotations may use, resolved once (the modules import each other, so this cannot happen at import time)."""
    global _NAMES
    if _NAMES is None:
        from . import curves, dates, global_types, inflation
        ns = {"np": np, "list": list, "dict": dict, "float": float, "int": int, "bool": bool, "str": str, "type": type, "None": None,
              "Optional": typing.Optional, "Dict": typing.Dict, "List": typing.List, "Union": typing.Union}
        for mod in (dates, global_types, curves, inflation):
            ns.update({k: v for k, v in vars(mod).items() if inspect.isclass(v)})
        _NAMES = ns
    return _NAMES


def to_usable_type(t):
    """An annotation as something `isinstance` accepts (helpers.py:508-527)."""
    origin = getattr(t, "__origin__", None)
    if origin is not None:
        if origin is list:
            return (list, np.ndarray)
        if origin is dict:
            return dict
        if origin is typing.Union:
            return tuple(to_usable_type(a) for a in t.__args__)
        return t
    if t is float:
        return (int, float, np.float64)
    if isinstance(t, tuple):
        return tuple(to_usable_type(a) for a in t)
    return t


def _flat(t):
    return tuple(x for a in t for x in _flat(a)) if isinstance(t, tuple) else (t,)


_PLANS: dict = {}


def _plan(func):
    """[(argument, allowed types, may be None)] of a function, resolved once."""
    key = getattr(func, "__func__", func)
    plan = _PLANS.get(key)
    if plan is None:
        defaults = {k: p.default for k, p in inspect.signature(func).parameters.items()}
        plan = []
        for name, ann in getattr(func, "__annotations__", {}).items():
            if name == "return":
                continue
            if isinstance(ann, str):          # annotations are strings under `from __future__ import annotations`
                try:
                    ann = eval(ann, {"__builtins__": {}}, _namespace())
                except Exception:  # noqa: BLE001  (a name this table does not know: not checked)
                    continue
            allowed = _flat(to_usable_type(ann))
            if all(inspect.isclass(a) for a in allowed):
                plan.append((name, allowed, defaults.get(name, inspect.Parameter.empty) is None))
        _PLANS[key] = plan
    return plan


def check_argument_types(func, values: dict) -> None:
    """Raises LibError("Argument Type Error") for the first annotated argument of `func` whose value in `values` (the caller's
    locals()) is not an instance of its annotation."""
    for name, allowed, none_ok in _plan(func):
        if name not in values:
            continue
        value = values[name]
        if value is None and none_ok:
            continue
        if not isinstance(value, allowed):
            raise LibError("Argument Type Error")
