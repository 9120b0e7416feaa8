"""Curve (re)build latency: cav_curve_rebuild_dev (bootstrap + tangents + tables) with CAV_BOOTSTRAP=1 (single CTA) / 2 (entry-parallel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from bench import load_curve
cv, curve = load_curve()
ctx = _native.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
d, J, H = ctx.curve_read()
r = torch.tensor(np.pad(curve.swap_rates, (0, 0)), dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.curve_rebuild_dev(r.data_ptr())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(50):
    ctx.curve_rebuild_dev(r.data_ptr())
b.record(stream); b.synchronize()
print(f"curve rebuild (G={len(d)}): {a.elapsed_time(b) / 50 * 1e3:.1f} us   checksums {d.sum()!r} {J.sum()!r} {H.sum()!r}")
