#!/usr/bin/env python
"""Known answers of the result containers' export helpers (cavour/requests/results.py: to_dict / to_json / to_csv / df / matrix
of Valuation, Ladder, Delta, Gamma, CrossGamma; Risk.__repr__ / has_cross_gamma / all_cross_gammas) from the UNMODIFIED
reference.  TEST INFRASTRUCTURE, build container only:

    PYTHONPATH=tests/golden/gen/refshim:/root/reference python tests/golden/gen/make_golden_results_export.py

Writes tests/golden/ref_results_export.json.
"""
import contextlib
import io
import json
import os

import numpy as np
import jax.numpy as jnp

from cavour.requests.results import Valuation, Delta, Gamma, CrossGamma, Risk
from cavour.utils.currency import CurrencyTypes
from cavour.utils.global_types import CurveTypes

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
TENORS = ["1D", "1W", "1M", "1Y", "2Y"]
TEN2 = ["1Y", "5Y", "10Y"]


def printed(obj):
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            obj.matrix
    except Exception as ex:  # noqa: BLE001  (Gamma.matrix formats its labels as numbers: tenor strings raise)
        return "error: " + type(ex).__name__ + ": " + str(ex)
    return buf.getvalue()


def main():
    rng = np.random.default_rng(9)
    lad = np.round(rng.normal(size=5) * 100, 6)
    lad[0] = 0.0
    g = np.round(rng.normal(size=(5, 5)), 6)
    g = g + g.T
    g[0, :] = 0.0
    g[:, 0] = 0.0
    x = np.round(rng.normal(size=(5, 3)), 6)
    v = Valuation(1234.5678, CurrencyTypes.GBP)
    d = Delta(jnp.array(lad), TENORS, CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    d2 = Delta(jnp.array(lad[:3] * 2), TEN2, CurrencyTypes.GBP, CurveTypes.USD_GBP_BASIS)
    gm = Gamma(jnp.array(g), TENORS, CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    cg = CrossGamma(jnp.array(x), TENORS, TEN2, CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS, CurrencyTypes.GBP)
    gnum = Gamma(jnp.array(g), [0.0027, 0.0192, 0.0833, 1.0, 2.0], CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA)
    risk = Risk([d, d2])
    grisk = Risk([gm], cross_gammas=[cg])
    out = {
        "inputs": {"ladder": lad.tolist(), "gamma": g.tolist(), "cross": x.tolist(), "tenors": TENORS, "tenors2": TEN2},
        "valuation": {"to_dict": v.to_dict(), "to_json": v.to_json(), "to_csv": v.to_csv(), "repr": repr(v)},
        "ladder": {"to_dict": d.ladder.to_dict(), "df_csv": d.ladder.df.to_csv(), "repr": repr(d.ladder)},
        "delta": {"to_dict": d.to_dict(), "to_json": d.to_json(), "to_csv": d.to_csv(), "repr": repr(d)},
        "gamma": {"to_dict": gm.to_dict, "to_json": gm.to_json(), "to_csv": gm.to_csv(), "matrix": printed(gm), "repr": repr(gm)},
        "gamma_numeric_tenors": {"matrix": printed(gnum)},
        "cross": {"to_dict": cg.to_dict, "to_json": cg.to_json(), "to_csv": cg.to_csv(), "matrix": printed(cg), "repr": repr(cg)},
        "risk": {"repr": repr(risk), "gamma_repr": repr(grisk),
                 "has": [grisk.has_cross_gamma(CurveTypes.GBP_OIS_SONIA, CurveTypes.USD_GBP_BASIS),
                         grisk.has_cross_gamma(CurveTypes.USD_GBP_BASIS, CurveTypes.GBP_OIS_SONIA)],
                 "all_keys": [list(k) for k in grisk.all_cross_gammas]},
    }
    with open(os.path.join(OUT, "ref_results_export.json"), "w") as f:
        json.dump(out, f)
    print(out["risk"], out["gamma"]["matrix"], out["cross"]["matrix"], sep="\n")


if __name__ == "__main__":
    main()
