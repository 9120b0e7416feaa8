"""Host-link probe under torchrun: every rank copies 12 MB and 36 MB of pinned memory to its GPU at the same time (barrier first);
prints the per-rank times - do the ranks share PCIe bandwidth on this box?"""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (12, 36):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"rank {local}/{world}: {mb} MB H2D median {1e3 * ts[5]:.3f} ms = {mb * 1.048576 / ts[5] / 1e3:.1f} GB/s", flush=True)
# launch latency of an empty kernel stream sync round trip
x = torch.zeros(1, device="cuda")
ts = []
for _ in range(200):
    t0 = time.perf_counter(); x.add_(1); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
ts.sort()
print(f"rank {local}/{world}: launch + sync round trip median {1e6 * ts[100]:.1f} us, p90 {1e6 * ts[180]:.1f} us", flush=True)
if world > 1: dist.destroy_process_group()
