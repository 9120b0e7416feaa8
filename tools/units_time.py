"""Time the units stage of cav_portfolio_value on the private layout (no parity gate: for diagnostic builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_book, flatten_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
layout = sys.argv[2] if len(sys.argv) > 2 else "private"
cv, curve = load_curve()
flat = flatten_book(make_book(curve, n), dedup=(layout == "dedup"))
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
ctx.portfolio_upload(flat)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ctx.profile(True)
best = [1e9, 1e9, 1e9]
for r in range(6):
    ctx.portfolio_value(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr()); ctx.sync()
    ms = ctx.last_kernel_ms()
    best = [min(a, b) for a, b in zip(best, ms)]
print(f"{layout} n={n} units={flat.n_units}: units {best[0]:.3f} ms  expand {best[1]:.3f} ms  totals {best[2]:.3f} ms  "
      f"-> {1e6 / n * best[0]:.3f} ms per 1M units-stage")
# checksum of the results (bit-identity of kernel variants across processes)
print("checksum", float(gm[:: max(1, n // 1000)].sum().item()), float(dl.sum().item()), float(agg[0].item()))
