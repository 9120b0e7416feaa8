"""Value a private-layout book (one unit per trade) with the tile kernel chosen by CAV_UNITS_WS and print digests of the
per-trade rows and the totals; with --oracle also the worst scaled error of a sample against the C oracle.
Used by tests/test_gpu_units_ws.py (the switch is read once per process, so each variant runs in its own process)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adrates_b200 import _native
from adrates_b200.market_data import readme_model
from adrates_b200.synthetic import flatten_book, make_book, reference_leg_tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
model = readme_model()
curve = model.curves.GBP_OIS_SONIA
book = make_book(curve, n, seed=11)
flat = flatten_book(book, dedup=False)
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
ctx.portfolio_upload(flat)
pv = torch.empty(n, dtype=torch.float64, device="cuda")
dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda")
agg = ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr()).copy()
ctx.sync()
out = {"ws": os.environ.get("CAV_UNITS_WS"), "n": n, "tiles": int(flat.tile_plan.n_tiles) if flat.tile_plan is not None else 0,
       "pv": hashlib.sha256(pv.cpu().numpy().tobytes()).hexdigest(), "delta": hashlib.sha256(dl.cpu().numpy().tobytes()).hexdigest(),
       "gamma": hashlib.sha256(gm.cpu().numpy().tobytes()).hexdigest(), "agg": [float(x) for x in agg],
       "agg_abs": [float(pv.abs().sum()), float(dl.abs().sum()), float(gm.abs().sum())]}
if "--oracle" in sys.argv:
    from oracle import cavour_oracle as orc, c_oracle
    k = 256
    plan = orc.plan_path_b(curve.swap_times, curve.year_fracs)
    d, J, C = orc.bootstrap_tables(curve.swap_rates, plan)
    tr = dict(sched=book.sched[:k], coupon=book.coupon[:k], notional=book.notional[:k], spread=book.spread[:k], fixed_sign=book.fixed_sign[:k])
    pv_c, dl_c, gm_c = c_oracle.ois_batch((plan["times"], d, J, C), curve._interp_type.value, reference_leg_tables(book), tr, dense=False)
    N = book.notional[:k]
    out["oracle_err"] = max(float(np.max(np.abs(pv[:k].cpu().numpy() - pv_c) / np.maximum(np.abs(pv_c), N))),
                            float(np.max(np.abs(dl[:k].cpu().numpy() - dl_c) / np.maximum(np.abs(dl_c), (N * 1e-4)[:, None]))),
                            float(np.max(np.abs(gm[:k].cpu().numpy() - gm_c) / np.maximum(np.abs(gm_c), (N * 1e-8)[:, None, None]))))
print("WSCHECK " + json.dumps(out))
