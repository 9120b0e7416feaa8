"""Differentiable portfolio PV: the CUDA valuation as a node of an autograd graph.

The north star asks for the kernels to sit behind a `jax.ffi` custom call with `custom_jvp` / `custom_vjp` rules so
that `jax.grad` / `jax.hessian` keep composing (INTEGRATION.md section 3 shows that binding; JAX is not installed in
this image, so it cannot be compiled or run here).  PyTorch is, and the same composition is provided for it: the
portfolio PV as a function of the par rates is a `torch.autograd.Function` whose backward is the delta ladder and
whose double-backward is the gamma matrix - one kernel launch delivers PV, dPV/dr and d2PV/dr2 (the analytic
Greeks of engine.py:2541-2574), so `torch.autograd.grad`, `torch.autograd.functional.hessian` and any further
composition through the rates need no finite differences and no second valuation algorithm.
"""
from __future__ import annotations

import torch

from . import _native


class RateSession:
    """A context with a bootstrap plan and a flat portfolio uploaded; `totals(rates)` re-bootstraps the curve from
    device-resident rates (cav_curve_rebuild_dev) and returns the 1057 totals [PV | ladder | gamma] on the device."""

    def __init__(self, curve, flat, device: int = 0):
        self.device = torch.device("cuda", device)
        self.R = len(curve.swap_rates)
        self.ctx = _native.Context(device)
        self.ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
        self.ctx.portfolio_upload(flat)
        self._agg = torch.zeros(_native.NOUT, dtype=torch.float64, device=self.device)
        self._rates32 = torch.zeros(32, dtype=torch.float64, device=self.device)

    def totals(self, rates: torch.Tensor) -> torch.Tensor:
        if rates.dtype != torch.float64 or rates.device != self.device or rates.numel() != self.R:
            raise ValueError(f"rates must be {self.R} float64 values on {self.device}")
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)   # order with the autograd graph
        self._rates32[:self.R] = rates.detach().reshape(-1)
        self.ctx.curve_rebuild_dev(self._rates32.data_ptr())
        mask = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA
        self.ctx.portfolio_value(mask, None, None, None, self._agg.data_ptr())
        return self._agg.clone()


class _Gradient(torch.autograd.Function):
    """dPV/dr as a differentiable function of r: its backward is the Hessian-vector product with the gamma matrix."""

    @staticmethod
    def forward(ctx, rates, session):
        agg = session.totals(rates)
        R = session.R
        ctx.hess = 1e8 * agg[33:].reshape(32, 32)[:R, :R]          # reported gamma is per bp^2
        return 1e4 * agg[1:1 + R]                                   # reported ladder is per bp

    @staticmethod
    def backward(ctx, g):
        return ctx.hess @ g, None


class _PresentValue(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rates, session):
        ctx.session = session
        ctx.save_for_backward(rates)
        return session.totals(rates)[0]

    @staticmethod
    def backward(ctx, g):
        (rates,) = ctx.saved_tensors
        return g * _Gradient.apply(rates, ctx.session), None


def portfolio_pv(rates: torch.Tensor, session: RateSession) -> torch.Tensor:
    """Portfolio PV as a twice-differentiable function of the par rates (decimal, one per curve pillar)."""
    return _PresentValue.apply(rates, session)
