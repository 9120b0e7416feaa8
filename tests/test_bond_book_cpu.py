"""Array-based bond books (bond_book.BondBook) against the object route (credit.Bond -> Flattener): the two flat
books must give the same per-bond PV / delta / gamma when evaluated with the same tables (tests/flat_eval.py is the
numpy evaluation of exactly the arrays the device receives)."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200 import batch as B
from adrates_b200.bond_book import BondBook
from adrates_b200.credit import Bond
from adrates_b200.dates import BusDayAdjustTypes, Date, DateGenRuleTypes, DayCountTypes, FrequencyTypes
from adrates_b200.error import LibError
from adrates_b200.flatten import Flattener
from adrates_b200.global_types import CurrencyTypes
from tests.flat_eval import eval_flat
from tests.test_batch_cpu import _curve

BOND_CONVS = {
    "semi_act365": dict(freq_type=FrequencyTypes.SEMI_ANNUAL, dc_type=DayCountTypes.ACT_365F),
    "annual_30e360_lag2": dict(freq_type=FrequencyTypes.ANNUAL, dc_type=DayCountTypes.THIRTY_E_360, payment_lag=2,
                               bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING),
    "quarterly_forward_eom": dict(freq_type=FrequencyTypes.QUARTERLY, dc_type=DayCountTypes.ACT_360,
                                  dg_type=DateGenRuleTypes.FORWARD, end_of_month=True),
}


def _random_bonds(curve, n, rng):
    vd = curve._value_dt._n
    issue = B.add_weekdays(np.full(n, vd), 0) + rng.integers(-3000, 200, n)
    months = rng.integers(3, 420, n)
    months[: n // 2] = 12 * rng.integers(1, 35, n // 2)
    issue[::4] = issue[1::4][: issue[::4].shape[0]]              # repeated schedules
    months[::4] = months[1::4][: months[::4].shape[0]]
    return dict(issue=issue, tenor_months=months, coupon=rng.uniform(0.005, 0.08, n),
                face_value=np.exp(rng.uniform(4, 16, n)))


def _object_flat(curve, spec, conv, dedup):
    fl = Flattener(curve)
    n = spec["issue"].shape[0]
    for i in range(n):
        iss = Date._of(int(spec["issue"][i]))
        fl.add_trade(Bond(iss, iss.add_tenor(f"{int(spec['tenor_months'][i])}M"), float(spec["coupon"][i]),
                          conv["freq_type"], conv["dc_type"], CurrencyTypes.GBP, float(spec["face_value"][i]),
                          conv.get("payment_lag", 0), bd_type=conv.get("bd_type", BusDayAdjustTypes.FOLLOWING),
                          dg_type=conv.get("dg_type", DateGenRuleTypes.BACKWARD),
                          end_of_month=conv.get("end_of_month", False)))
    return fl.finalize(dedup=dedup)


@pytest.mark.parametrize("conv", list(BOND_CONVS))
@pytest.mark.parametrize("dedup", [True, False])
def test_bond_book_flatten_equals_object_flatten(ref_curves, conv, dedup):
    cv = ref_curves["gbp_readme_lzr"]
    curve = _curve(cv)
    rng = np.random.default_rng(29)
    n = 120
    spec = _random_bonds(curve, n, rng)
    book = BondBook.from_arrays(curve, **spec, **BOND_CONVS[conv])
    flat = book.flatten(dedup=dedup, tiles=False)
    ref = _object_flat(curve, spec, BOND_CONVS[conv], dedup)
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    got = eval_flat(flat, d, J, C)
    exp = eval_flat(ref, d, J, C)
    assert flat.n_trades == ref.n_trades == n and flat.n_pairs == ref.n_pairs == 2
    assert np.any(exp[0] == 0.0) and np.any(exp[0] != 0.0)          # matured and live bonds are both in the sample
    face = spec["face_value"]
    for g, e, scale in zip(got, exp, (face, face * 1e-4 * 40, face * 1e-8 * 1600)):
        s = scale.reshape((-1,) + (1,) * (g.ndim - 1))
        assert np.max(np.abs(g - e) / np.maximum(np.abs(e), s)) < 1e-12
    if dedup:                                                       # sharing by schedule class, not by identical bond
        assert flat.n_units <= 2 * np.unique(np.stack([spec["issue"], spec["tenor_months"]]), axis=1).shape[1]
        assert np.all(np.diff(flat.group_offsets) <= 256) and sorted(flat.out_index.tolist()) == list(range(n))
    tiled = book.flatten(dedup=dedup, tiles=True)
    assert tiled.tile_plan is not None


def test_bond_book_errors(ref_curves):
    curve = _curve(ref_curves["gbp_readme_lzr"])
    vd = curve._value_dt
    with pytest.raises(LibError):
        BondBook.from_arrays(curve, issue=[vd], coupon=0.03)
    with pytest.raises(LibError):
        BondBook.from_arrays(curve, issue=[vd], maturity=[vd.add_days(-5)], coupon=0.03)
    with pytest.raises(LibError):
        BondBook.from_arrays(curve, issue=[vd], tenor_years=5, coupon=0.0)
    with pytest.raises(LibError):
        BondBook.from_arrays(curve, issue=[vd], tenor_years=5)
    gone = BondBook.from_arrays(curve, issue=np.full(3, vd._n - 5000), tenor_years=2, coupon=0.04)
    z = gone.flatten(tiles=False)                                   # everything matured: valid all-zero layout
    assert z.n_units == 1 and np.all(z.comp_weight == 0.0) and np.all(z.amt == 0.0)


def _golden_books():
    """One single-bond book per vanilla (fixed-coupon, bullet) bond of the reference goldens, with its curve."""
    from tests.conftest import load_golden
    from tests.util_bonds import build_bond_model
    g = load_golden("ref_bonds.json")
    m = build_bond_model(g)
    curve_of = {"GBP": m.curves.GBP_OIS_SONIA, "USD": m.curves.USD_OIS_SOFR}
    for b in g["bonds"]:
        if b["amortization"] is not None or b["coupon"] == 0.0:
            continue
        curve = curve_of[b["currency"]]
        iss = Date(*b["issue"])
        mat = iss.add_tenor(b["maturity"]) if isinstance(b["maturity"], str) else Date(*b["maturity"])
        book = BondBook.from_arrays(curve, issue=[iss], maturity=[mat], coupon=b["coupon"], face_value=b["face"],
                                    freq_type=FrequencyTypes[b["freq"]], dc_type=DayCountTypes[b["dc"]],
                                    payment_lag=b["payment_lag"])
        yield b, curve, book


def test_bond_book_matches_bonds_valued_by_the_reference_engine():
    """tests/golden/ref_bonds.json holds VALUE / DELTA / GAMMA of Engine._compute_bond of the unmodified reference."""
    seen = 0
    for b, curve, book in _golden_books():
        plan = orc.plan_path_b(curve.swap_times, curve.year_fracs)
        d, J, C = orc.bootstrap_tables(curve.swap_rates, plan)
        T = max(len(b["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[b["freq"]], 1.0)
        for dedup in (True, False):
            pv, dl, gm = eval_flat(book.flatten(dedup=dedup, tiles=False), d, J, C)
            ref_d, ref_g = np.array(b["delta"]), np.array(b["gamma"])
            R = ref_d.shape[0]
            assert abs(pv[0] - b["value"]) <= 1e-10 * max(abs(b["value"]), b["face"]), b["id"]
            assert np.max(np.abs(dl[0, :R] - ref_d) / np.maximum(np.abs(ref_d), b["face"] * 1e-4 * T)) < 1e-10, b["id"]
            assert np.max(np.abs(gm[0, :R, :R] - ref_g) / np.maximum(np.abs(ref_g), b["face"] * 1e-8 * T * T)) < 1e-10, b["id"]
        seen += 1
    assert seen >= 6
