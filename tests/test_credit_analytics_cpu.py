"""Stand-alone bond / FRN analytics (adrates_b200/credit_analytics.py: value, accrued interest, dirty / clean price, yield,
z-spread, duration, convexity, DV01 / CS01, discount margin, modified duration, amortisation schedules) against known answers
of the unmodified reference's methods (tests/golden/ref_credit_analytics.json, tests/golden/gen/make_golden_credit_analytics.py)
on the bonds / notes of the engine goldens.  Host code in the reference and here: path-A curve look-ups and a scalar root search."""
import numpy as np
import pytest

from adrates_b200 import Bond, CurrencyTypes, CurveTypes, Date, DayCountTypes, FRN, FrequencyTypes, LibError
from adrates_b200.credit import BOND_CURVE
from tests.conftest import load_golden
from tests.util_bonds import build_bond_model, make_bond, make_frn

TOL = 1e-12          # of the face value (prices: of 100)
ROOT_TOL = 1e-9      # yields / spreads come out of a root search (brentq xtol 2e-12, discount margin xtol 1e-8)


@pytest.fixture(scope="module")
def setup():
    g = load_golden("ref_credit_analytics.json")
    bonds = load_golden("ref_bonds.json")
    frns = load_golden("ref_frn.json")
    return g, bonds, frns, build_bond_model(bonds)


def _near(a, b, tol, scale=1.0):
    assert np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))) <= tol * scale, (a, b)


def test_bond_analytics_match_reference(setup):
    g, bonds, _, model = setup
    vd, settle = model.value_dt, Date(*g["settle"])
    specs = {b["id"]: b for b in bonds["bonds"]}
    assert len(g["bonds"]) == 9
    for rec in g["bonds"]:
        b = make_bond(specs[rec["id"]])
        c = model.curves[BOND_CURVE[b._currency].name]
        face = b._face_value
        for tag, s in (("vd", vd), ("settle", settle)):
            r = rec[tag]
            _near(b.value(vd, c, 0.0, s), r["value"], TOL, face)
            _near(b.value(vd, c, 0.0075, s), r["value_z75"], TOL, face)
            _near(b._payment_dfs, r["payment_dfs"], TOL)              # the lists the last valuation left (z-spread 75 bp)
            _near(b._coupon_pvs, r["coupon_pvs"], TOL, face)
            _near(b._principal_pvs, r["principal_pvs"], TOL, face)
            assert b.accrued_interest(s) == r["accrued"]
            _near(b.dirty_price(vd, c, 0.0, s), r["dirty"], TOL, 100)
            clean = b.clean_price(vd, c, 0.0, s)
            _near(clean, r["clean"], TOL, 100)
            _near(b.clean_price(vd, c, 0.0075, s), r["clean_z75"], TOL, 100)
            _near(b.yield_to_maturity(s, clean), r["ytm"], ROOT_TOL)
            _near(b.yield_to_maturity(s, clean - 2.0), r["ytm_minus2"], ROOT_TOL)
            _near(b.z_spread(s, c, clean - 2.0), r["z_spread_minus2"], ROOT_TOL)
            _near(b.duration(s, c), r["duration"], 1e-8)
            _near(b.duration(s, c, "macaulay", 0.0075), r["macaulay_z75"], 1e-8)
            _near(b.convexity(s, c), r["convexity"], 1e-7)
            _near(b.dv01(s, c), r["dv01"], TOL, face)
            _near(b.cs01(s, c, 0.0075), r["cs01_z75"], TOL, face)
        assert b.current_yield() == rec["current_yield"]
        if "g_spread" in rec:
            other = model.curves["USD_OIS_SOFR" if b._currency == CurrencyTypes.GBP else "GBP_OIS_SONIA"]
            _near(b.g_spread(settle, other, rec["settle"]["clean"] - 1.0), rec["g_spread"], ROOT_TOL)
            _near(b.i_spread(settle, c, rec["settle"]["clean"] - 1.0), rec["i_spread"], ROOT_TOL)
        if tag == "settle" and s is settle:
            assert b.value(vd, c) == b.value(vd, c, 0.0, vd)          # settlement defaults to the value date
    with pytest.raises(ValueError, match="Unknown duration type"):
        b.duration(vd, c, "effective")
    after = b._maturity_dt.add_days(5)                 # a matured bond: zero value, and the duration sums divide by zero as the reference's
    assert b.value(vd, c, 0.0, after) == 0.0
    with pytest.raises(ZeroDivisionError):
        b.duration(after, c)


def test_amortisation_schedules_match_reference(setup):
    g = setup[0]
    for key, ref in g["equal_principal"].items():
        face, n = key.split("/")
        assert Bond.generate_equal_principal_schedule(float(face), int(n)) == ref
    for key, ref in g["annuity"].items():
        face, n, cpn, fq = key.split("/")
        _near(Bond.generate_annuity_schedule(float(face), int(n), float(cpn), FrequencyTypes[fq]), ref, 1e-15, float(face))
    with pytest.raises(LibError, match="Number of periods must be positive"):
        Bond.generate_equal_principal_schedule(100.0, 0)
    with pytest.raises(LibError, match="Number of periods must be positive"):
        Bond.generate_annuity_schedule(100.0, 0, 0.04, FrequencyTypes.ANNUAL)
    # a schedule made by the generator is accepted by the constructor
    sched = Bond.generate_annuity_schedule(100.0, 5, 0.04, FrequencyTypes.ANNUAL)
    b = Bond(Date(30, 4, 2024), "5Y", 0.04, FrequencyTypes.ANNUAL, DayCountTypes.ACT_365F, CurrencyTypes.GBP, amortization_schedule=sched)
    assert abs(sum(b._principal_payments) - 100.0) < 1e-12


def test_frn_analytics_match_reference(setup):
    g, _, frns, model = setup
    vd, settle = model.value_dt, Date(*g["settle"])
    specs = {f["id"]: f for f in frns["frns"] + frns.get("dual", [])}
    n_err = 0
    for rec in g["frns"]:
        f = make_frn({**specs[rec["id"]]})
        disc, idx = model.curves[BOND_CURVE[f._currency].name], model.curves[f._floating_index.name]
        face = f._face_value
        for tag, s in (("vd", vd), ("settle", settle)):
            r = rec[tag]
            if "error" in r:
                with pytest.raises(LibError) as ex:
                    f.value(vd, disc, idx, 0.0, s)
                assert "LibError: " + str(ex.value) == r["error"]
                n_err += 1
                continue
            _near(f.value(vd, disc, idx, 0.0, s), r["value"], TOL, face)
            _near(f.value(vd, disc, idx, 0.002, s), r["value_dm20"], TOL, face)
            _near(f._rates, r["rates"], TOL)                          # the lists the last valuation left (margin 20 bp)
            _near(f._coupon_payments, r["coupon_payments"], TOL, face)
            _near(f._payment_dfs, r["payment_dfs"], TOL)
            _near(f._payment_pvs, r["payment_pvs"], TOL, face)
            _near(f.dirty_price(vd, disc, idx, 0.0, s), r["dirty"], TOL, 100)
            assert abs(f.accrued_interest(s) - r["accrued"]) <= 1e-15
            clean = f.clean_price(vd, disc, idx, 0.0, s)
            _near(clean, r["clean"], TOL, 100)
            _near(f.discount_margin(s, disc, idx, clean - 0.5), r["dm_minus_half"], 2e-8)
            _near(f.modified_duration(vd, disc, idx, 0.0, s), r["mod_duration"], 1e-8)
            _near(f.dv01(vd, disc, idx, 0.002, s), r["dv01_dm20"], TOL, face)
    assert n_err == 2 and len(g["frns"]) == 10
    collar = FRN(vd, "4Y", 0.003, FrequencyTypes.QUARTERLY, DayCountTypes.ACT_365F, CurrencyTypes.GBP, CurveTypes.GBP_OIS_SONIA,
                 cap_rate=0.045, floor_rate=0.04)
    gbp = model.curves.GBP_OIS_SONIA
    _near(collar.value(vd, gbp), g["collar"]["value"], TOL, 100)
    _near(collar._rates, g["collar"]["rates"], TOL)
    assert min(collar._rates) >= 0.04 and max(collar._rates) <= 0.045
    _near(collar.clean_price(vd, gbp), g["collar"]["clean"], TOL, 100)
    with pytest.raises(LibError, match="Discount curve is required"):
        collar.value(vd, None)
