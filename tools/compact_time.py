"""Expansion stage of the 1M-trade step with L2 flushed before each step, per-kernel times from the library's events:
request masks (7 = PV + delta + gamma, 4 = gamma only) x CAV_EXPAND_COMPACT (read on every valuation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cv, curve = load_curve()
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
make_array_book(curve, n).upload(ctx)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx.profile(True)
def run(mode, mask, reps=30):
    os.environ["CAV_EXPAND_COMPACT"] = str(mode)
    ks = np.zeros(3)
    for i in range(reps + 3):
        flush.zero_(); torch.cuda.synchronize()
        ctx.portfolio_value(mask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr()); ctx.sync()
        if i >= 3:
            ks += np.array(ctx.last_kernel_ms()[:3])
    return ks / reps
for mask in (7, 4, 7, 4):
    for mode in (0, 1):
        ks = run(mode, mask)
        print(f"mask={mask} compact={mode}: step {ks.sum():.4f} ms   units / expand / totals {np.round(ks, 4)}")
