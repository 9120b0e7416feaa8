"""Time k_expand with different request masks (dedup layout, 1M trades)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_book, flatten_book
from bench import load_curve
n = 1_000_000
cv, curve = load_curve()
flat = flatten_book(make_book(curve, n), dedup=True, max_group=int(os.environ.get("MAXG", "256")))
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
ctx.portfolio_upload(flat)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.profile(True)
for mask, name in ((7, "pv+delta+gamma"), (4, "gamma only"), (3, "pv+delta")):
    best = 1e9
    for r in range(6):
        ctx.portfolio_value(mask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr()); ctx.sync()
        best = min(best, ctx.last_kernel_ms()[1])
    print(f"groups={flat.n_groups} {name}: k_expand {best:.3f} ms = {n*8192/best/1e6:.0f} GB/s of gamma rows")
