"""Where the e2e step goes: upload (H2D + validation + tile plan) vs valuation vs first-use table builds."""
import os, sys, time, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_book, flatten_book
from bench import load_curve
n = 1_000_000
cv, curve = load_curve()
flat = flatten_book(make_book(curve, n), dedup=True)
fp = copy.copy(flat)
for k in ("unit_offsets", "amt", "weight", "node", "comp_weight", "group_offsets", "group_units", "out_index", "unit_weight"):
    a = getattr(flat, k)
    if a is not None:
        setattr(fp, k, torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy())
ctx = _native.Context(0)
ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = np.empty(1057)
def t(fn, reps=10):
    fn(); ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync(); return (time.perf_counter() - t0) / reps * 1e3
up = t(lambda: ctx.portfolio_upload(fp))
tp = fp.tile_plan; fp2 = copy.copy(fp); fp2.tile_plan = None
up_notiles = t(lambda: ctx.portfolio_upload(fp2))
ctx.portfolio_upload(fp)
val = t(lambda: ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg))
both = t(lambda: (ctx.portfolio_upload(fp), ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg)))
print(f"h2d bytes {flat.h2d_bytes()/1e6:.1f} MB; upload {up:.3f} ms (without tile plan {up_notiles:.3f}); value_host (tables cached) {val:.3f} ms; upload+value {both:.3f} ms -> {n/both/1e3:.1f} M trades/s")

# ---- phase timing of the end-to-end sequence with and without the pipelined upload
def phases(async_on, reps=20):
    ctx.set_async_upload(async_on)
    acc = np.zeros(4)
    for i in range(reps + 2):
        ctx.sync(); t0 = time.perf_counter()
        fp_nt = copy.copy(fp); fp_nt.tile_plan = None
        ctx.portfolio_upload(fp_nt); t1 = time.perf_counter()
        ctx.portfolio_set_tiles(fp.tile_plan); t2 = time.perf_counter()
        ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg); t3 = time.perf_counter()
        if i >= 2:
            acc += [t1 - t0, t2 - t1, t3 - t2, t3 - t0]
    acc *= 1e3 / reps
    print(f"async={async_on}: upload {acc[0]:.3f} ms, set_tiles {acc[1]:.3f} ms, value_host {acc[2]:.3f} ms, total {acc[3]:.3f} ms -> {n/acc[3]/1e3:.1f} M trades/s")
phases(False); phases(True); phases(False); phases(True)
