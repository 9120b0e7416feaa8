#!/usr/bin/env python
"""Public names and call signatures of the reference modules on or next to the path (classes, methods, functions; parameter
NAMES and which of them have defaults - no code), read with `ast` from /root/reference.  TEST INFRASTRUCTURE, build container
only:

    python tests/golden/gen/make_api_names.py

Writes tests/golden/ref_api_names.json; tests/test_api_surface_cpu.py holds this package against it.
"""
import ast
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ref_api_names.json")
MODULES = [
    "cavour/utils/date.py", "cavour/utils/calendar.py", "cavour/utils/day_count.py", "cavour/utils/schedule.py",
    "cavour/utils/frequency.py", "cavour/utils/global_types.py", "cavour/utils/currency.py", "cavour/utils/error.py",
    "cavour/market/curves/interpolator.py", "cavour/market/curves/discount_curve.py", "cavour/market/curves/inflation_curve.py",
    "cavour/market/indices/inflation_index.py", "cavour/market/position/position.py", "cavour/market/portfolio/portfolio.py",
    "cavour/requests/results.py", "cavour/models/models.py",
    "cavour/trades/rates/ois.py", "cavour/trades/rates/ois_curve.py", "cavour/trades/rates/swap_fixed_leg.py",
    "cavour/trades/rates/swap_float_leg.py", "cavour/trades/rates/xccy_basis_swap.py", "cavour/trades/rates/xccy_curve.py",
    "cavour/trades/rates/xccy_fix_float_swap.py", "cavour/trades/rates/xccy_fix_fix_swap.py", "cavour/trades/rates/zcis.py",
    "cavour/trades/rates/swap_inflation_leg.py", "cavour/trades/rates/yoy_inflation_swap.py",
    "cavour/trades/rates/swap_yoy_inflation_leg.py", "cavour/trades/credit/bond.py", "cavour/trades/credit/frn.py",
]


def signature(fn: ast.FunctionDef):
    a = fn.args
    args = a.posonlyargs + a.args
    names = [x.arg for x in args]
    n_default = len(a.defaults)
    out = [{"name": n, "default": i >= len(names) - n_default} for i, n in enumerate(names)]
    for rec, x in zip(out, args):
        if x.annotation is not None:
            rec["ann"] = ast.unparse(x.annotation)
    out += [{"name": x.arg, "default": d is not None, "kwonly": True} for x, d in zip(a.kwonlyargs, a.kw_defaults)]
    if a.vararg:
        out.append({"name": "*" + a.vararg.arg})
    if a.kwarg:
        out.append({"name": "**" + a.kwarg.arg})
    return out


def is_property(fn):
    return any(isinstance(d, ast.Name) and d.id == "property" for d in fn.decorator_list)


def self_attrs(cls: ast.ClassDef):
    """Attributes the constructor leaves on the object: `self.x = ...` in __init__ and in the schedule generators it calls."""
    names = set()
    for sub in cls.body:
        if isinstance(sub, ast.FunctionDef) and (sub.name == "__init__" or "generate" in sub.name):
            for node in ast.walk(sub):
                targets = []
                if isinstance(node, ast.Assign):
                    targets = node.targets
                elif isinstance(node, (ast.AugAssign, ast.AnnAssign)):
                    targets = [node.target]
                for t in targets:
                    for leaf in ast.walk(t):
                        if isinstance(leaf, ast.Attribute) and isinstance(leaf.value, ast.Name) and leaf.value.id == "self":
                            names.add(leaf.attr)
    return sorted(names)


def main():
    out = {}
    for rel in MODULES:
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        mod = {"classes": {}, "functions": {}}
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                members, enum_members = {}, []
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef):
                        members[sub.name] = {"params": signature(sub), "property": is_property(sub),
                                             "checks_types": any(isinstance(n, ast.Call) and getattr(n.func, "id", "") == "check_argument_types"
                                                                 for n in ast.walk(sub))}
                    elif isinstance(sub, ast.Assign) and all(isinstance(t, ast.Name) for t in sub.targets):
                        enum_members += [t.id for t in sub.targets]
                mod["classes"][node.name] = {"methods": members, "assigned": enum_members,
                                             "bases": [ast.unparse(b) for b in node.bases], "init_attrs": self_attrs(node)}
            elif isinstance(node, ast.FunctionDef):
                mod["functions"][node.name] = {"params": signature(node)}
        out[rel] = mod
    with open(OUT, "w") as f:
        json.dump(out, f, sort_keys=True)
    print("modules", len(out), "classes", sum(len(m["classes"]) for m in out.values()))


if __name__ == "__main__":
    main()
