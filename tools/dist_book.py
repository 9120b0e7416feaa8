"""torchrun --nproc-per-node N tools/dist_book.py [trades]: OISBook.compute_distributed over NCCL, totals vs rank-0 single-GPU run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from adrates_b200 import RequestTypes
from adrates_b200.synthetic import make_array_book
from bench import load_curve
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
cv, curve = load_curve()
book = make_array_book(curve, n)
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
book.compute_distributed(ALL)                      # warm-up (tables, allocations, NCCL communicator)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
res, rows, (lo, hi) = book.compute_distributed(ALL)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if dist.get_rank() == 0:
    ref, _ = book.compute(ALL, device=local)
    err = max(abs(res.value.amount - ref.value.amount) / abs(ref.value.amount),
              float(np.max(np.abs(res.gamma.risk_ladder - ref.gamma.risk_ladder)) / np.max(np.abs(ref.gamma.risk_ladder))))
    print(f"world {dist.get_world_size()}: {n} trades incl. array flattening {dt*1e3:.1f} ms; shard0 [{lo},{hi}); totals vs single GPU rel err {err:.2e}")
dist.destroy_process_group()
