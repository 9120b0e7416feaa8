"""The headline step for profilers: device flatten of the 1M-trade array book, three PV+delta+gamma valuations, one PV+delta
valuation, one 2 000 x 100 000 scenario call.  No CPU oracle, no extras (short replays under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book, shocked_rate_scenarios
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cv, curve = load_curve()
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
book = make_array_book(curve, n)
book.upload(ctx)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.portfolio_value(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr())
ctx.sync()
ctx.portfolio_value(3, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr())
ctx.sync()
if "--scen" in sys.argv:
    sub = make_array_book(curve, 100_000)
    sub.upload(ctx, tiles=False)
    pnl = torch.empty(2000, 100_000, dtype=torch.float64, device="cuda")
    ctx.scenarios(shocked_rate_scenarios(curve, 2000), pnl.data_ptr())
print("ok", float(agg[0].item()))
