"""Multi-GPU sharding of the valuation path.

Trades are independent, so a book shards across ranks with no data-path collective; the only
exchange is one all-reduce of the 1057 portfolio totals [PV, ladder(32), gamma(32x32)] -
`Portfolio.compute` sum semantics (cavour/market/portfolio/portfolio.py:48-65).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(cost_per_trade, world: int):
    """Contiguous trade ranges [lo, hi) per rank with balanced total cost (e.g. cashflow counts)."""
    cost = np.asarray(cost_per_trade, dtype=np.float64)
    n = cost.shape[0]
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (world - 1)
    csum = np.cumsum(cost)
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left") + 1
    cuts = np.clip(cuts, 0, n)
    edges = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(edges)):
        edges[i] = max(edges[i], edges[i - 1])
    return [(edges[i], edges[i + 1]) for i in range(world)]


def all_reduce_totals(totals):
    """Sum the [1057] totals tensor over all ranks in place (NCCL on GPU tensors, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(totals)
    return totals


def exchange_handles(local: bytes, rank: int, world: int):
    """All ranks' comm handles in rank order (torch.distributed all_gather of fixed-size byte blobs; any backend)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != world or dist.get_rank() != rank:
        raise RuntimeError("exchange_handles: torch.distributed must be initialised with this rank / world size")
    mine = torch.frombuffer(bytearray(local), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        mine = mine.cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


def init_device_allreduce(ctx, rank: int | None = None, world: int | None = None):
    """Connect the context to its peers (one process per GPU of one NVSwitch box) for the in-kernel all-reduce of the
    portfolio totals: afterwards valuations with REQ_ALLREDUCE return whole-job totals with no collective call in the step
    (csrc/cav_comm.cu).  No-op for a single rank.  Returns True when the device path is active."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    if world <= 1:
        return False
    if getattr(ctx, "_comm_world", 0) == world:
        return True
    handles = exchange_handles(ctx.comm_local_handle(), rank, world)
    ctx.comm_init(rank, world, handles)
    return True
