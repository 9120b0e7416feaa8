"""torchrun --nproc-per-node N tools/dist_scen.py [scenarios] [trades]: OISBook.scenario_values_distributed over N GPUs
(scenarios shard by rank, no data-path collective); the slices are all-gathered over NCCL here only to check them
against rank 0's single-GPU matrix."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from adrates_b200.synthetic import make_array_book, shocked_rate_scenarios
from bench import load_curve
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
cv, curve = load_curve()
book = make_array_book(curve, n)
rates = shocked_rate_scenarios(curve, S)
book.scenario_values_distributed(rates)                         # warm-up
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
rows, (lo, hi) = book.scenario_values_distributed(rates)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
sizes = [0] * world
dist.all_gather_object(sizes, (lo, hi))
parts = [torch.empty(h - l, n, dtype=torch.float64, device="cuda") for l, h in sizes]
if len({h - l for l, h in sizes}) == 1:
    dist.all_gather(parts, rows)
else:                                                           # ragged slices: one broadcast per rank
    for r in range(world):
        if r == rank:
            parts[r].copy_(rows)
        dist.broadcast(parts[r], src=r)
if rank == 0:
    full = book.scenario_values(rates, device=local)
    same = torch.equal(torch.cat(parts, 0), full)
    print(f"world {world}: {S} scenarios x {n} trades incl. array flattening {dt*1e3:.1f} ms; slices {sizes}; "
          f"gathered matrix identical to the single-GPU one: {same}")
    assert same
dist.destroy_process_group()
