"""Scenario revaluation timing (BASELINE config 4 slice): S shocked curves x T trades on one GPU."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book, shocked_rate_scenarios
from bench import load_curve

ap = argparse.ArgumentParser()
ap.add_argument("--scen", type=int, default=2000)
ap.add_argument("--trades", type=int, default=100000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-offset-bd", type=int, default=250, help="forward starts up to this many business days (1: ~200 units, all hot)")
ap.add_argument("--variants", default="22,23,33,22,33")
a = ap.parse_args()
cv, curve = load_curve()
book = make_array_book(curve, a.trades, seed=20240430, max_offset_bd=a.max_offset_bd)        # the book of bench.py's config-4 extra, flattened on the device
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
book.upload(ctx, tiles=False)
info = ctx.book_info()
class flat: n_units, n_terms = info["n_units"], info["n_terms"]
shocked = shocked_rate_scenarios(curve, a.scen)
pnl = torch.empty(a.scen, a.trades, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ref = None
for variant in a.variants.split(","):
  os.environ["CAV_SCEN_EXPAND"], os.environ["CAV_SCEN_UNITS"] = variant[0], variant[1]
  for r in range(a.reps):
    t0 = time.perf_counter()
    ctx.scenarios(shocked, pnl.data_ptr())
    ctx.sync()
    dt = time.perf_counter() - t0
    if r == 0:
        chk = pnl.sum().item()
        ref = chk if ref is None else ref
        assert chk == ref
    if r == a.reps - 1: print("  info", ctx.scenarios_info())
    print(f"expand/units variant {variant} rep {r}: {dt*1e3:.2f} ms  {a.scen*a.trades/dt/1e9:.2f} G revaluations/s  units={flat.n_units} terms={flat.n_terms}")
