"""One device flatten + valuation of the BASELINE-size array book (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adrates_b200 import RequestTypes
from adrates_b200.synthetic import make_array_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cv, curve = load_curve()
book = make_array_book(curve, n)
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
for _ in range(2):
    res, rows = book.compute(ALL)
    torch.cuda.synchronize()
print("PV", res.value.amount)
