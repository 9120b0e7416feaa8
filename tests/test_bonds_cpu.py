"""Bond host layer (schedule, coupon payments) and the flattener's bond units, evaluated per term on the CPU with
the oracle's curve tables, against Position(bond, model).compute of the unmodified reference (ref_bonds.json)."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200 import Date, LibError, FrequencyTypes, DayCountTypes, CurrencyTypes, Bond
from adrates_b200.flatten import Flattener
from tests.conftest import load_golden
from tests.flat_eval import eval_flat
from tests.util_bonds import build_bond_model, make_bond, make_frn


@pytest.fixture(scope="module")
def g():
    return load_golden("ref_bonds.json")


def test_bond_schedule_and_coupons_match_reference(g):
    for b in g["bonds"]:
        bond = make_bond(b)
        assert [[d.d(), d.m(), d.y()] for d in bond._payment_dts] == b["payment_dts"], b["id"]
        assert np.allclose(bond._coupon_payments, b["coupon_payments"], rtol=1e-15, atol=0), b["id"]


def test_bond_units_match_reference_engine(g):
    """Flattened bond -> per-term evaluation with the oracle's bootstrap tables = reference VALUE / DELTA / GAMMA."""
    m = build_bond_model(g)
    for b in g["bonds"]:
        bond = make_bond(b)
        curve = getattr(m.curves, bond._floating_index.name)
        plan = orc.plan_path_b(curve.swap_times, curve.year_fracs)
        d, J, C = orc.bootstrap_tables(curve.swap_rates, plan)
        fl = Flattener(curve)
        fl.add_trade(bond)
        pv, dl, gm = eval_flat(fl.finalize(dedup=False), d, J, C)
        N, T = b["face"], max(len(b["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[b["freq"]], 1.0)
        R = len(b["delta"])
        assert abs(pv[0] - b["value"]) <= 1e-10 * max(abs(b["value"]), N), b["id"]
        assert np.max(np.abs(dl[0][:R] - b["delta"]) / np.maximum(np.abs(b["delta"]), N * 1e-4 * T)) < 1e-10, b["id"]
        ref_g = np.array(b["gamma"])
        assert np.max(np.abs(gm[0][:R, :R] - ref_g) / np.maximum(np.abs(ref_g), N * 1e-8 * T * T)) < 1e-10, b["id"]


def test_bond_errors_like_reference():
    with pytest.raises(LibError, match="Issue date must be before maturity date"):
        Bond(Date(1, 1, 2025), Date(1, 1, 2024), 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.GBP)
    with pytest.raises(LibError, match="Amortization schedule length"):
        Bond(Date(1, 1, 2024), "5Y", 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.GBP,
             amortization_schedule=[50.0, 0.0])
    b = Bond(Date(1, 1, 2024), "5Y", 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.JPY)
    with pytest.raises(LibError, match="No default OIS curve for currency"):
        b._floating_index


def test_frn_units_match_reference_engine():
    """Engine._compute_frn (single curve): schedule, then flattened FRN -> per-term evaluation = reference Greeks."""
    g = load_golden("ref_frn.json")
    m = build_bond_model(g)
    for f in g["frns"]:
        frn = make_frn(f)
        assert [[d.d(), d.m(), d.y()] for d in frn._payment_dts] == f["payment_dts"], f["id"]
        assert np.allclose(frn._year_fracs, f["year_fracs"], rtol=1e-15, atol=0), f["id"]
        curve = getattr(m.curves, frn._floating_index.name)
        plan = orc.plan_path_b(curve.swap_times, curve.year_fracs)
        d, J, C = orc.bootstrap_tables(curve.swap_rates, plan)
        fl = Flattener(curve)
        fl.add_trade(frn)
        pv, dl, gm = eval_flat(fl.finalize(dedup=False), d, J, C)
        N, T = f["face"], max(len(f["payment_dts"]) / {"ANNUAL": 1, "SEMI_ANNUAL": 2, "QUARTERLY": 4}[f["freq"]], 1.0)
        R = len(f["delta"])
        assert abs(pv[0] - f["value"]) <= 1e-10 * max(abs(f["value"]), N), f["id"]
        assert np.max(np.abs(dl[0][:R] - f["delta"]) / np.maximum(np.abs(f["delta"]), N * 1e-4 * T)) < 1e-10, f["id"]
        ref_g = np.array(f["gamma"])
        assert np.max(np.abs(gm[0][:R, :R] - ref_g) / np.maximum(np.abs(ref_g), N * 1e-8 * T * T)) < 1e-10, f["id"]
