"""Feasibility probe for pipelining books: does the device flattener (a chain of ~100 small kernels with four host syncs)
keep its pace on a high-priority stream while another context's HBM-bound expansion runs on a low-priority stream?"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adrates_b200 import _native
from adrates_b200.synthetic import make_array_book
from bench import load_curve
n = 1_000_000
cv, curve = load_curve()
prio_b = int(sys.argv[1]) if len(sys.argv) > 1 else -1
sA = torch.cuda.Stream(priority=0); sB = torch.cuda.Stream(priority=prio_b)
ctxA = _native.Context(0); ctxA.set_stream(sA.cuda_stream)
ctxB = _native.Context(0); ctxB.set_stream(sB.cuda_stream)
for c in (ctxA, ctxB):
    c.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
book = make_array_book(curve, n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
for k in ("effective", "_tenor", "fixed_sign", "coupon", "notional"):
    setattr(book, k, pin(getattr(book, k)))
book.upload(ctxA); ctxA.sync()
pv = torch.empty(n, dtype=torch.float64, device="cuda"); dl = torch.empty(n, 32, dtype=torch.float64, device="cuda")
gm = torch.empty(n, 32, 32, dtype=torch.float64, device="cuda"); agg = torch.zeros(1057, dtype=torch.float64, device="cuda")
stop = False; cntA = [0]
def loopA():
    while not stop:
        ctxA.portfolio_value(7, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr()); ctxA.sync(); cntA[0] += 1
def timeB(reps=40):
    book.upload(ctxB); ctxB.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        book.upload(ctxB); ctxB.sync()
    return (time.perf_counter() - t0) / reps * 1e3
print(f"priority of B {prio_b}: flatten alone {timeB():.3f} ms")
t0 = time.perf_counter(); c0 = cntA[0]
th = threading.Thread(target=loopA); th.start()
time.sleep(0.2)
c1 = cntA[0]; t1 = time.perf_counter()
fb = timeB()
c2 = cntA[0]; t2 = time.perf_counter()
stop = True; th.join()
print(f"valuation loop alone: {1e3 * 0.2 / max(1, c1 - c0):.3f} ms/step;  with flatten beside it: {1e3 * (t2 - t1) / max(1, c2 - c1):.3f} ms/step;  flatten beside valuation {fb:.3f} ms")
