#!/usr/bin/env python
"""Known answers of the reference's non-AD interpolation (cavour/market/curves/interpolator.py: `interpolate`,
`Interpolator.fit / interpolate`; cavour/market/curves/discount_curve.py: `DiscountCurve.df`) for every InterpTypes member.
TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_interp.py

Writes tests/golden/ref_interp.json.
"""
import json
import os

import numpy as np

from cavour.utils.date import Date
from cavour.market.curves.interpolator import Interpolator, InterpTypes, interpolate
from cavour.market.curves.discount_curve import DiscountCurve

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
NODE_SETS = {
    "from_zero": ([0.0, 0.25, 1.0, 2.0, 5.0, 10.0, 30.0], [1.0, 0.99, 0.955, 0.91, 0.80, 0.64, 0.30]),
    "late_start": ([0.5, 1.0, 2.0, 5.0], [0.98, 0.95, 0.90, 0.80]),
    "inverted": ([0.0, 1.0, 2.0, 3.0, 7.0], [1.0, 0.94, 0.895, 0.86, 0.75]),
    "two_nodes": ([1.0, 2.0], [0.95, 0.90]),
}
QUERIES = [0.0, 1e-13, 0.1, 0.25, 0.6, 1.0, 1.5, 2.0, 3.7, 5.0, 8.0, 10.0, 29.0, 30.0, 31.0, 45.0]
NODE_SCHEMES = ("FLAT_FWD_RATES", "LINEAR_FWD_RATES", "LINEAR_ZERO_RATES")


def flt(v):
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    return [float(x) for x in a]


def main():
    out = {"queries": QUERIES, "node_sets": {k: {"times": t, "dfs": d} for k, (t, d) in NODE_SETS.items()}, "function": {},
           "class_scalar": {}, "class_array": {}, "curve_df": {}}
    for name, (t, d) in NODE_SETS.items():
        times, dfs = np.array(t), np.array(d)
        for it in InterpTypes:
            key = f"{name}/{it.name}"
            qs = [q for q in QUERIES if not (name == "two_nodes" and it.name == "LINEAR_FWD_RATES" and q > 1.0)]   # needs 3 nodes
            if it.name in NODE_SCHEMES:
                out["function"][key] = {"q": qs, "scalar": [float(interpolate(float(q), times, dfs, it.value)) for q in qs],
                                        "array": flt(interpolate(np.array(qs), times, dfs, it.value))}
                continue      # the class runs these through jax arrays (`.size` of the torch stand-in is a method): same arithmetic
            f = Interpolator(it)
            f.fit(times, dfs)
            res = [f.interpolate(float(q)) for q in qs]
            out["class_scalar"][key] = {"q": qs, "v": [flt(r)[0] for r in res],
                                        "is_array": [bool(isinstance(r, np.ndarray)) for r in res]}
            out["class_array"][key] = {"q": qs, "v": flt(f.interpolate(np.array(qs)))}
    # DiscountCurve.df over dates, every scheme (the curve prepends (0, 1) unless the first date is the value date)
    vd = Date(30, 4, 2024)
    offsets, values = [0.5, 1.0, 2.0, 5.0, 10.0], [0.975, 0.95, 0.90, 0.78, 0.60]
    dates = [vd, vd.add_tenor("1M"), vd.add_tenor("9M"), vd.add_tenor("18M"), vd.add_tenor("4Y"), vd.add_tenor("10Y"),
             vd.add_tenor("12Y")]
    out["curve_df_inputs"] = {"value_dt": [30, 4, 2024], "offsets": offsets, "values": values,
                              "dates": [[x._d, x._m, x._y] for x in dates]}
    for it in InterpTypes:
        c = DiscountCurve(vd, offsets, np.array(values), it)
        singles = [c.df(x) for x in dates]
        out["curve_df"][it.name] = {"single": [flt(s)[0] for s in singles],
                                    "single_is_array": [bool(isinstance(s, np.ndarray)) for s in singles],
                                    "list": flt(c.df(dates)), "times": flt(c._times), "dfs": flt(c._dfs)}
    with open(os.path.join(OUT, "ref_interp.json"), "w") as f:
        json.dump(out, f)
    print("wrote", sum(len(v) for v in out.values() if isinstance(v, dict)))


if __name__ == "__main__":
    main()
