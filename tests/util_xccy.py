"""Helpers for the cross-currency tests: build the golden model with the host layer."""
from adrates_b200 import (Date, DayCountTypes, FrequencyTypes, BusDayAdjustTypes, SwapTypes, InterpTypes, CurveTypes,
                          CurrencyTypes, XccyBasisSwap)
from adrates_b200.models import Model


def build_xccy_model(g, ois_interp=InterpTypes.LINEAR_ZERO_RATES, xccy_name="GBP_USD_BASIS"):
    vd = Date(*g["value_dt"])
    m = Model(vd)
    for name, px, dc in (("GBP_OIS_SONIA", g["gbp_px"], DayCountTypes.ACT_365F),
                         ("USD_OIS_SOFR", g["usd_px"], DayCountTypes.ACT_360)):
        m.build_curve(name=name, px_list=px, tenor_list=g["tenors"], spot_days=0, swap_type=SwapTypes.PAY,
                      fixed_dcc_type=dc, fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                      float_dc_type=dc, bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                      interp_type=ois_interp)
    m.build_xccy_curve(name=xccy_name, domestic_curve_name="USD_OIS_SOFR", foreign_curve_name="GBP_OIS_SONIA",
                       basis_spreads=g["basis_bps"], tenor_list=g["basis_tenors"], spot_fx=g["spot_fx"],
                       domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.QUARTERLY)
    return m


def make_xccy_trade(t):
    return XccyBasisSwap(effective_dt=Date(*t["effective"]), term_dt_or_tenor=t["tenor"],
                         domestic_notional=t["domestic_notional"], foreign_notional=t["foreign_notional"],
                         domestic_spread=t["domestic_spread"], foreign_spread=t["foreign_spread"],
                         domestic_freq_type=FrequencyTypes[t["domestic_freq"]],
                         foreign_freq_type=FrequencyTypes[t["foreign_freq"]], domestic_dc_type=DayCountTypes.ACT_360,
                         foreign_dc_type=DayCountTypes.ACT_365F, domestic_floating_index=CurveTypes.USD_OIS_SOFR,
                         foreign_floating_index=CurveTypes.GBP_OIS_SONIA, domestic_currency=CurrencyTypes.USD,
                         foreign_currency=CurrencyTypes.GBP)
