"""Wall-clock of OISBook.from_arrays(...).compute() for the BASELINE-size book, pageable and pinned inputs
(CAV_BOOK_TRACE=1 prints the phases of cav_book_from_arrays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from adrates_b200 import RequestTypes
from adrates_b200.batch import OISBook, add_weekdays
from adrates_b200.dates import BusDayAdjustTypes, DayCountTypes, FrequencyTypes
from adrates_b200.position import CurveSession
from bench import load_curve

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cv, curve = load_curve()
rng = np.random.Generator(np.random.PCG64(20240430))
tenor = rng.integers(1, 51, n).astype(np.int32)
offset = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 251, n))
eff = add_weekdays(np.full(n, curve._value_dt._n), offset)
coupon = rng.uniform(0.01, 0.06, n)
notional = np.exp(rng.uniform(np.log(1e5), np.log(1e8), n))
sign = np.where(rng.random(n) < 0.5, 1.0, -1.0)
conv = dict(fixed_freq_type=FrequencyTypes.ANNUAL, fixed_dc_type=DayCountTypes.ACT_365F, float_freq_type=FrequencyTypes.ANNUAL,
            float_dc_type=DayCountTypes.ACT_365F, bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING)
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
for label, arrs in (("pageable", (eff, tenor, sign, coupon, notional)), ("pinned", tuple(pin(a) for a in (eff, tenor, sign, coupon, notional)))):
    e, t, s, c, no = arrs
    sess = CurveSession.get(curve, 0)
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        book = OISBook.from_arrays(curve, e, tenor_years=t, fixed_sign=s, fixed_coupon=c, notional=no, **conv)
        t1 = time.perf_counter()
        where = book.upload(sess.ctx)
        sess.ctx.sync()
        t2 = time.perf_counter()
        res, rows = book.compute(ALL)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        res2, _ = book.compute(ALL, per_trade=False)
        t4 = time.perf_counter()
        print(f"{label:9s} rep {rep}: from_arrays {1e3*(t1-t0):.3f} ms | upload({where}) {1e3*(t2-t1):.3f} ms | compute(rows) "
              f"{1e3*(t3-t2):.3f} ms | compute(totals only) {1e3*(t4-t3):.3f} ms | PV {res.value.amount:.6e}", flush=True)
        del rows
