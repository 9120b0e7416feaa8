"""PV + delta (BASELINE config 2) step time on the device-built 1M-trade book: CUDA events over back-to-back valuations and the
per-stage times of one profiled valuation; checksum for bit-identity across kernel variants (CAV_EXPAND_ROWS=1|2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cv, curve = load_curve()
ctx = _native.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
make_array_book(curve, n).upload(ctx, tiles=False)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
M = 3
for _ in range(3):
    ctx.portfolio_value(M, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(20):
    ctx.portfolio_value(M, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr())
b.record(stream); b.synchronize()
hot = a.elapsed_time(b) / 20
cold = []
for _ in range(10):
    flush.zero_()
    a.record(stream)
    ctx.portfolio_value(M, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr())
    b.record(stream); b.synchronize()
    cold.append(a.elapsed_time(b))
ctx.profile(True)
ctx.portfolio_value(M, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr()); ctx.sync()
print(f"PV+delta n={n}: back-to-back {hot:.4f} ms, L2 flushed {min(cold):.4f} ms (median {sorted(cold)[5]:.4f}); stages units/expand/totals {ctx.last_kernel_ms()}")
print("checksum", float(dl.sum().item()), float(pv.sum().item()), float(agg[1:33].sum().item()))
