"""Public scenario-revaluation calls against the reference's own way of doing it: Model.scenario (rebuilt curve,
models.py:507-557) + Position.compute([VALUE]) per trade and scenario (position.py:62-80)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from adrates_b200 import RequestTypes, batch as B  # noqa: E402
from adrates_b200.error import LibError  # noqa: E402
from adrates_b200.position import Portfolio  # noqa: E402
from tests.test_batch_cpu import CONVS, _random_book  # noqa: E402
from tests.util_trades import build_model, make_trade  # noqa: E402

TOL = 1e-10
SHOCKS = [0.01, -0.25, {"10Y": 0.05, "2Y": -0.03}, {"1W": 0.2, "50Y": -0.1}, 0.0]


def test_portfolio_scenario_values_match_the_scenario_loop(ref_curves, ref_trades):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    specs = [s for s in ref_trades if s["curve"] == "gbp_readme_lzr"]
    swaps = [make_trade(s, cv) for s in specs]
    pf = Portfolio([sw.position(model) for sw in swaps])
    got = pf.scenario_values(cv["name"], SHOCKS).cpu().numpy()
    assert got.shape == (len(SHOCKS), len(swaps))
    for s, shock in enumerate(SHOCKS):
        shocked = model.scenario(cv["name"], shock)
        for i, sw in enumerate(swaps):
            ref = sw.position(shocked).compute([RequestTypes.VALUE]).value.amount
            assert abs(got[s, i] - ref) <= TOL * max(abs(ref), specs[i]["notional"]), (shock, specs[i]["id"])
    pnl = pf.scenario_values(cv["name"], SHOCKS, pnl=True).cpu().numpy()
    base = got[-1]                                              # the 0.0 shock
    scale = np.maximum(np.abs(got), np.array([s["notional"] for s in specs])[None, :])
    assert np.max(np.abs(pnl - (got - base[None, :])) / scale) < TOL
    assert np.max(np.abs(pnl[-1]) / scale[-1]) < 1e-13          # zero shock: the two kernels differ by summation order only
    assert Portfolio([]).scenario_values(cv["name"], SHOCKS).shape == (len(SHOCKS), 0)
    other = build_model(cv)
    with pytest.raises(LibError, match="share one Model"):
        Portfolio([swaps[0].position(model), swaps[1].position(other)]).scenario_values(cv["name"], SHOCKS)


@pytest.mark.parametrize("conv", ["annual_act365", "lagged"])
def test_array_book_scenario_values_match_books_on_rebuilt_curves(ref_curves, conv):
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves[cv["name"]]
    rng = np.random.default_rng(5)
    n = 400
    spec = _random_book(curve, n, rng, spread=(conv != "annual_act365"))
    book = B.OISBook.from_arrays(curve, **spec, **CONVS[conv])
    rates = model.scenario_rates(cv["name"], SHOCKS)
    for dedup in ((True,) if conv == "lagged" else (True, False)):
        got = book.scenario_values(rates, dedup=dedup).cpu().numpy()
        assert got.shape == (len(SHOCKS), n)
        for s, shock in enumerate(SHOCKS):
            shocked_curve = model.scenario(cv["name"], shock).curves[cv["name"]]
            _, rows = B.OISBook.from_arrays(shocked_curve, **spec, **CONVS[conv]).compute([RequestTypes.VALUE], dedup=dedup)
            ref = rows["pv"].cpu().numpy()
            assert np.max(np.abs(got[s] - ref) / np.maximum(np.abs(ref), spec["notional"])) < TOL, (shock, dedup)
    # no process group: the distributed call values every scenario on this rank
    rows, (lo, hi) = book.scenario_values_distributed(rates, device=0)
    assert (lo, hi) == (0, len(SHOCKS)) and np.array_equal(rows.cpu().numpy(), book.scenario_values(rates).cpu().numpy())
    out = torch.empty(len(SHOCKS), n, dtype=torch.float64, device="cuda")
    assert book.scenario_values(rates, out=out) is out
    with pytest.raises(LibError):
        book.scenario_values(rates[:, :5])
    with pytest.raises(LibError):
        book.scenario_values(rates, out=torch.empty(2, n, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("n_scen", [2, 38, 130, 256])
@pytest.mark.parametrize("n", [1237, 1238])
def test_scenario_expansion_kernels_agree_bitwise(ref_curves, monkeypatch, n_scen, n):
    """Even scenario counts take the bulk-store expansion kernel (128 x 32 tiles written by cp.async.bulk; even trade counts
    only) or the 32x128 kernel with 16-byte reads, odd ones (and CAV_SCEN_EXPAND=1) the 64x64 kernel: same sums in the same
    order, so the matrices are identical, ragged tile edges included (1237 / 1238 trades: neither a multiple of 32 nor of 128)."""
    from adrates_b200.synthetic import shocked_rate_scenarios
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves[cv["name"]]
    rng = np.random.default_rng(8)
    spec = _random_book(curve, n, rng)
    book = B.OISBook.from_arrays(curve, **spec, **CONVS["annual_act365"])
    rates = shocked_rate_scenarios(curve, n_scen)
    out = {}
    for variant in ("1", "2", "3"):
        monkeypatch.setenv("CAV_SCEN_EXPAND", variant)
        monkeypatch.setenv("CAV_SCEN_UNITS", variant)             # ... and the unit sums (2: two scenarios per thread, 3: chains)
        out[variant] = book.scenario_values(rates).cpu().numpy()
    assert np.array_equal(out["1"], out["2"])
    assert np.array_equal(out["1"], out["3"])
    odd = book.scenario_values(rates[: n_scen - 1]).cpu().numpy()     # odd count: always the 64x64 kernel
    assert np.array_equal(odd, out["2"][: n_scen - 1])


@pytest.mark.parametrize("device_flatten", [True, False])
def test_prefix_chain_unit_sums_agree_bitwise(ref_curves, monkeypatch, device_flatten):
    """A book shaped like BASELINE config 2 / 4 (a few start dates x tenors 1..50Y): the annuity units of one start date are
    prefixes of each other, k_scen_units_chain walks one list per chain and must give the per-unit kernels' values bit for
    bit - for the device-built and the host-built book alike - while doing a fraction of their gathers."""
    from adrates_b200.position import CurveSession
    from adrates_b200.synthetic import make_array_book, shocked_rate_scenarios
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves[cv["name"]]
    book = make_array_book(curve, 6000, seed=5)
    rates = shocked_rate_scenarios(curve, 66)
    out = {}
    for variant in ("2", "3"):
        monkeypatch.setenv("CAV_SCEN_EXPAND", variant)
        monkeypatch.setenv("CAV_SCEN_UNITS", variant)
        sess = CurveSession.get(curve, 0)
        book.upload(sess.ctx, tiles=False, device_flatten=device_flatten)
        pnl = torch.empty(66, book.n_trades, dtype=torch.float64, device="cuda")
        sess.ctx.scenarios(rates, pnl.data_ptr())
        sess.ctx.sync()
        out[variant] = pnl.cpu().numpy()
        info = sess.ctx.scenarios_info()
        assert info["chain_kernel"] == (variant == "3")
        if variant == "3":
            n_terms = sess.ctx.book_info()["n_terms"] if device_flatten else None
            assert info["chains"] > 0 and info["queries"] > 0
            if n_terms:
                assert info["chain_terms"] * 3 <= n_terms       # at least three times fewer gathers on this book
    assert np.array_equal(out["2"], out["3"])
    api = book.scenario_values(rates).cpu().numpy()                 # the public call, default kernels, output in trade order
    assert api.shape == out["3"].shape
