"""Array-based bond books on the GPU (bond_book.BondBook -> C ABI): per-bond rows against bonds valued by the unmodified
reference engine (tests/golden/ref_bonds.json), against the numpy evaluation of the same flat arrays, and the book
totals against Portfolio([...Bond objects...]).compute."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import cavour_oracle as orc  # noqa: E402
from adrates_b200 import Portfolio, RequestTypes  # noqa: E402
from adrates_b200.bond_book import BondBook  # noqa: E402
from adrates_b200.credit import Bond  # noqa: E402
from adrates_b200.dates import Date, DayCountTypes, FrequencyTypes  # noqa: E402
from adrates_b200.global_types import CurrencyTypes  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.flat_eval import eval_flat  # noqa: E402
from tests.test_bond_book_cpu import BOND_CONVS, _random_bonds  # noqa: E402
from tests.util_bonds import build_bond_model  # noqa: E402

REQ = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
TOL = 1e-10


def test_bond_book_rows_match_the_reference_engine():
    """Every vanilla golden bond sits at row 0 of a book of 80 bonds with its conventions."""
    g = load_golden("ref_bonds.json")
    m = build_bond_model(g)
    curve_of = {"GBP": m.curves.GBP_OIS_SONIA, "USD": m.curves.USD_OIS_SOFR}
    rng = np.random.default_rng(12)
    seen = 0
    for b in g["bonds"]:
        if b["amortization"] is not None or b["coupon"] == 0.0:
            continue
        curve = curve_of[b["currency"]]
        iss = Date(*b["issue"])
        mat = iss.add_tenor(b["maturity"]) if isinstance(b["maturity"], str) else Date(*b["maturity"])
        fill = _random_bonds(curve, 79, rng)
        from adrates_b200 import batch as B
        issue = np.concatenate([[iss._n], fill["issue"]])
        maturity = np.concatenate([[mat._n], B.add_tenor(fill["issue"], fill["tenor_months"], "M")])
        book = BondBook.from_arrays(curve, issue=issue, maturity=maturity,
                                    coupon=np.concatenate([[b["coupon"]], fill["coupon"]]),
                                    face_value=np.concatenate([[b["face"]], fill["face_value"]]),
                                    freq_type=FrequencyTypes[b["freq"]], dc_type=DayCountTypes[b["dc"]],
                                    payment_lag=b["payment_lag"])
        T = max(len(b["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[b["freq"]], 1.0)
        ref_d, ref_g = np.array(b["delta"]), np.array(b["gamma"])
        R = ref_d.shape[0]
        for dedup in (True, False):
            _, rows = book.compute(REQ, dedup=dedup)
            pv, dl, gm = (rows[k].cpu().numpy() for k in ("pv", "delta", "gamma"))
            assert abs(pv[0] - b["value"]) <= TOL * max(abs(b["value"]), b["face"]), b["id"]
            assert np.max(np.abs(dl[0, :R] - ref_d) / np.maximum(np.abs(ref_d), b["face"] * 1e-4 * T)) < TOL, b["id"]
            assert np.max(np.abs(gm[0, :R, :R] - ref_g) / np.maximum(np.abs(ref_g), b["face"] * 1e-8 * T * T)) < TOL, b["id"]
        seen += 1
    assert seen >= 6


@pytest.mark.parametrize("conv", list(BOND_CONVS))
def test_bond_book_compute_matches_flat_arrays_and_portfolio(ref_curves, conv):
    from tests.util_trades import build_model
    cv = ref_curves["gbp_readme_lzr"]
    model = build_model(cv)
    curve = model.curves.GBP_OIS_SONIA
    rng = np.random.default_rng(31)
    n = 160
    spec = _random_bonds(curve, n, rng)
    c = BOND_CONVS[conv]
    book = BondBook.from_arrays(curve, **spec, **c)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    positions = []
    for i in range(n):
        iss = Date._of(int(spec["issue"][i]))
        kw = {k: c[k] for k in ("bd_type", "dg_type", "end_of_month") if k in c}
        positions.append(Bond(iss, iss.add_tenor(f"{int(spec['tenor_months'][i])}M"), float(spec["coupon"][i]),
                              c["freq_type"], c["dc_type"], CurrencyTypes.GBP, float(spec["face_value"][i]),
                              c.get("payment_lag", 0), **kw).position(model))
    ref_tot = Portfolio(positions).compute(REQ)
    face = spec["face_value"]
    F = float(np.sum(face))
    for dedup in (True, False):
        res, rows = book.compute(REQ, dedup=dedup)
        exp = eval_flat(book.flatten(dedup=dedup, tiles=False), d, J, C)
        for key, e, scale in zip(("pv", "delta", "gamma"), exp, (face, face * 1e-4 * 40, face * 1e-8 * 1600)):
            got = rows[key].cpu().numpy()
            s = scale.reshape((-1,) + (1,) * (got.ndim - 1))
            assert np.max(np.abs(got - e) / np.maximum(np.abs(e), s)) < TOL, (key, dedup)
        assert abs(res.value.amount - ref_tot.value.amount) <= TOL * F
        assert np.max(np.abs(res.risk.risk_ladder - ref_tot.risk.risk_ladder)) <= TOL * F * 1e-4 * 40
        assert np.max(np.abs(res.gamma.risk_ladder - ref_tot.gamma.risk_ladder)) <= TOL * F * 1e-8 * 1600
    # scenario values of the book: row s = the book valued on the curve rebuilt from the shocked quotes
    shocks = [0.05, {"10Y": -0.1}]
    rates = model.scenario_rates(cv["name"], shocks)
    got = book.scenario_values(rates).cpu().numpy()
    for s, shock in enumerate(shocks):
        shocked_curve = model.scenario(cv["name"], shock).curves[cv["name"]]
        _, rows = BondBook.from_arrays(shocked_curve, **spec, **c).compute([RequestTypes.VALUE])
        ref = rows["pv"].cpu().numpy()
        assert np.max(np.abs(got[s] - ref) / np.maximum(np.abs(ref), face)) < TOL
