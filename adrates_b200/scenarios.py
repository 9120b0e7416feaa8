"""Batched scenario revaluation: many shocked curves x one book in one device pass.

The reference revalues a scenario as `Model.scenario(curve, shock)` (a new Model with the curve rebuilt from the
shocked quotes, models.py:507-557) followed by `Position.compute([VALUE])` per trade (position.py:62-80), i.e.
S x N Python round trips and S x N path-B bootstraps (engine.py:2246-2360).  Here the S shocked par-rate vectors
(`Model.scenario_rates`) go to the device once, every curve is re-bootstrapped there (DFs only, `k_scen_bootstrap`),
and the flat book is valued against all of them (`cav_scenarios`): the result is the [S, N] matrix of trade values
in trade order, or the P&L against the base curve.  Scenarios are independent: under torch.distributed each rank takes
a contiguous slice of them (no collective, SURVEY 8e / BASELINE config 4).
"""
from __future__ import annotations

import numpy as np

from .curves import OISCurve
from .error import LibError


def scenario_bounds(n_scen: int, world: int):
    """Contiguous scenario ranges [lo, hi) per rank, sizes differing by at most one."""
    world = max(int(world), 1)
    base, extra = divmod(int(n_scen), world)
    edges = [0]
    for r in range(world):
        edges.append(edges[-1] + base + (1 if r < extra else 0))
    return [(edges[r], edges[r + 1]) for r in range(world)]


def check_rates(curve: OISCurve, rates) -> np.ndarray:
    r = np.ascontiguousarray(rates, dtype=np.float64)
    if r.ndim != 2 or r.shape[1] != len(curve.swap_rates):
        raise LibError(f"shocked rates must be [S, {len(curve.swap_rates)}] (decimal par rates), got {r.shape}")
    if not np.all(np.isfinite(r)):
        raise LibError("shocked rates must be finite")
    return r


def scenario_values_flat(curve: OISCurve, flat, rates, device: int = 0, pnl: bool = False, out=None):
    """Values of the trades of `flat` (a FlatPortfolio on `curve`) under every row of `rates` [S, R]: torch CUDA tensor
    [S, n_trades], FP64.  pnl=True subtracts the values on the unshocked curve.  `out` (optional) is a preallocated
    contiguous [S, n_trades] CUDA tensor."""
    return scenario_values_uploaded(curve, flat.n_trades, lambda ctx: ctx.portfolio_upload(flat), rates, device, pnl, out)


def scenario_values_uploaded(curve: OISCurve, n_trades: int, upload, rates, device: int = 0, pnl: bool = False, out=None):
    """As scenario_values_flat, with the portfolio put on the device by `upload(ctx)` (a host upload of a FlatPortfolio or
    the device flattener of an array book)."""
    import torch
    from . import _native
    from .position import CurveSession
    rates = check_rates(curve, rates)
    S, n = rates.shape[0], n_trades
    dev = torch.device("cuda", device)
    if out is None:
        out = torch.empty(S, n, dtype=torch.float64, device=dev)
    elif out.shape != (S, n) or out.dtype != torch.float64 or not out.is_contiguous() or out.device != dev:
        raise LibError(f"out must be a contiguous float64 [{S}, {n}] tensor on {dev}")
    if S == 0 or n == 0:
        return out
    sess = CurveSession.get(curve, device)
    upload(sess.ctx)
    sess.ctx.scenarios(rates, out.data_ptr())
    if pnl:
        base = torch.empty(n, dtype=torch.float64, device=dev)
        sess.ctx.portfolio_value(_native.REQ_VALUE, base.data_ptr())
        sess.ctx.sync()
        out -= base[None, :]
    else:
        sess.ctx.sync()
    return out
