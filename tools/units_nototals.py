"""Units stage of the private layout with and without the portfolio totals (partials accumulated in the tile kernel's epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_book, flatten_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
cv, curve = load_curve()
flat = flatten_book(make_book(curve, n), dedup=False)
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
ctx.portfolio_upload(flat)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.profile(True)
for name, a in (("with totals", agg.data_ptr()), ("no totals", None), ("with totals", agg.data_ptr()), ("no totals", None)):
    best = 1e9
    for r in range(6):
        ctx.portfolio_value(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), a); ctx.sync()
        best = min(best, ctx.last_kernel_ms()[0])
    print(f"private n={n} {name}: units {best:.3f} ms -> {1e6 / n * best:.3f} ms per 1M")
