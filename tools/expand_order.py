"""Does the scattered row order (out_index) cost store bandwidth?  k_expand with and without the permutation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_book, flatten_book
from bench import load_curve
n = 1_000_000
cv, curve = load_curve()
flat = flatten_book(make_book(curve, n), dedup=True)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
for name, oi in (("scattered rows (out_index)", flat.out_index), ("rows in group order (identity)", None)):
    flat.out_index = oi
    ctx = _native.Context(0)
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    ctx.portfolio_upload(flat)
    torch.cuda.synchronize()
    ctx.profile(True)
    best = 1e9
    for r in range(8):
        ctx.portfolio_value(4, None, None, gm.data_ptr(), agg.data_ptr()); ctx.sync()
        best = min(best, ctx.last_kernel_ms()[1])
    print(f"{name}: k_expand (gamma rows) {best:.3f} ms = {n*8192/best/1e6:.0f} GB/s")
    ctx.close()
