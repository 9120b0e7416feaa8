"""Sweep of the pipelined upload's chunking (count, growth, expansion streams): e2e ms per 1M trades from host buffers."""
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
ctx.set_async_upload(True)
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = np.empty(1057)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    ctx.portfolio_upload(fp)
    ctx.portfolio_value_host(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg)
def run(reps=30):
    for _ in range(3): step()
    tot = 0.0
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize(); ctx.sync()
        t0 = time.perf_counter(); step(); tot += time.perf_counter() - t0
    return tot / reps * 1e3
ref = None
for chunks, streams, growth in [(2, 1, 2), (2, 2, 2), (3, 2, 2), (4, 2, 2), (4, 1, 2), (4, 2, 1), (6, 2, 2), (6, 2, 1), (8, 2, 2), (8, 2, 1), (2, 1, 2)]:
    os.environ["CAV_UP_CHUNKS"] = str(chunks); os.environ["CAV_EXPAND_STREAMS"] = str(streams); os.environ["CAV_UP_GROWTH"] = str(growth)
    ms = run()
    g = gm.sum().item()
    if ref is None: ref = (g, agg.copy())
    ok = (g == ref[0]) and np.array_equal(agg, ref[1])
    print(f"chunks {chunks} streams {streams} growth {growth}: {ms:.3f} ms  identical={ok}", flush=True)
