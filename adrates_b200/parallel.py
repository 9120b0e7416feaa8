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
