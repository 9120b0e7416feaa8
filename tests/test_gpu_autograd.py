"""torch.autograd composition through the CUDA valuation (the PyTorch analogue of the jax.ffi + custom_jvp binding):
grad = 1e4 x ladder, hessian = 1e8 x gamma, and both compose with ordinary tensor code."""
import numpy as np
import pytest
import torch

from adrates_b200 import RequestTypes, Portfolio
from adrates_b200.autograd import RateSession, portfolio_pv
from adrates_b200.flatten import Flattener
from tests.util_trades import build_model, make_trade

pytestmark = pytest.mark.gpu
ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]


def test_grad_and_hessian_compose(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    swaps = [make_trade(s, cv) for s in ref_trades if s["curve"] == "gbp_readme_lzr"][:12]
    ref = Portfolio([s.position(model) for s in swaps]).compute(ALL)
    fl = Flattener(curve)
    for s in swaps:
        fl.add_trade(s)
    sess = RateSession(curve, fl.finalize(dedup=True))
    r = torch.tensor(curve.swap_rates, dtype=torch.float64, device="cuda", requires_grad=True)
    pv = portfolio_pv(r, sess)
    assert abs(float(pv.detach()) - ref.value.amount) <= 1e-10 * 1e8
    (g,) = torch.autograd.grad(pv, r, create_graph=True)
    assert np.max(np.abs(g.detach().cpu().numpy() * 1e-4 - ref.risk.risk_ladder)) <= 1e-10 * 1e8 * 1e-4 * 50
    H = torch.autograd.functional.hessian(lambda x: portfolio_pv(x, sess), r.detach())
    assert np.max(np.abs(H.cpu().numpy() * 1e-8 - ref.gamma.risk_ladder)) <= 1e-10 * 1e8 * 1e-8 * 2500
    # composition: d/dr [ PV(r)^2 + sum(r) ] = 2 PV dPV/dr + 1, and a Hessian-vector product through grad
    (g2,) = torch.autograd.grad(portfolio_pv(r, sess) ** 2 + r.sum(), r)
    assert torch.allclose(g2, 2.0 * pv.detach() * g.detach() + 1.0, rtol=1e-12, atol=0.0)
    v = torch.linspace(-1.0, 1.0, r.numel(), dtype=torch.float64, device="cuda")
    (hv,) = torch.autograd.grad(g @ v, r)
    assert torch.allclose(hv, H @ v, rtol=1e-11, atol=1e-6)
    # the rates are a real input: a bumped curve revalues (first-order Taylor within second-order error)
    bump = torch.zeros_like(r); bump[24] = 1e-4
    pv_b = float(portfolio_pv(r.detach() + bump, sess))
    taylor = float(pv.detach()) + float(g.detach() @ bump) + 0.5 * float(bump @ H @ bump)
    assert abs(pv_b - taylor) <= 1e-6 * abs(float(g.detach() @ bump))
