"""One BASELINE config-4 call (S shocked curves x 100k trades) for profilers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book, shocked_rate_scenarios
from bench import load_curve
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
cv, curve = load_curve()
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
make_array_book(curve, n).upload(ctx, tiles=False)
print("book", ctx.book_info())
shocked = shocked_rate_scenarios(curve, S)
pnl = torch.empty(S, n, dtype=torch.float64, device="cuda")
import time
for r in range(int(os.environ.get("REPS", "2"))):
    t0 = time.perf_counter(); ctx.scenarios(shocked, pnl.data_ptr()); ctx.sync(); print(f"{1e3 * (time.perf_counter() - t0):.3f} ms")
print("ok", float(pnl[0, :10].sum().item()))
