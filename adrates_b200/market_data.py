"""Market data of the reference's README / intro notebook (quotes, not code): the 32-pillar GBP SONIA curve of
30-Apr-2024 (README section 1; BASELINE configs 1-4), the USD SOFR quotes and GBP/USD basis spreads of README section 7
(BASELINE config 5), and builders that turn them into curves / models through the reference-facing API.  bench.py,
__graft_entry__.smoke() and the examples use these; the tests keep their own golden copies (tests/golden/)."""
from __future__ import annotations

VALUE_DT = (30, 4, 2024)
TENORS = ['1D', '1W', '2W', '1M', '2M', '3M', '4M', '5M', '6M', '7M', '8M', '9M', '10M', '11M', '1Y', '18M', '2Y', '3Y', '4Y', '5Y', '6Y', '7Y', '8Y', '9Y', '10Y', '12Y', '15Y', '20Y', '25Y', '30Y', '40Y', '50Y']
GBP_SONIA_PX = [5.1998, 5.2014, 5.2003, 5.2027, 5.2023, 5.19281, 5.1656, 5.1482, 5.1342, 5.1173, 5.1013, 5.0862, 5.0701, 5.054, 5.0394, 4.8707, 4.75483, 4.532, 4.3628, 4.2428, 4.16225, 4.1132, 4.08505, 4.0762, 4.078, 4.0961, 4.12195, 4.1315, 4.113, 4.07724, 3.984, 3.88]
USD_SOFR_PX = [5.35, 5.32, 5.31, 5.29, 5.27, 5.25, 5.23, 5.21, 5.19, 5.17, 5.15, 5.13, 5.11, 5.09, 5.07, 4.95, 4.85, 4.7, 4.58, 4.48, 4.41, 4.36, 4.32, 4.29, 4.27, 4.28, 4.3, 4.32, 4.31, 4.29, 4.24, 4.18]
BASIS_TENORS = ['1Y', '2Y', '3Y', '5Y', '7Y', '10Y']
BASIS_BPS = [-5.0, -8.0, -10.0, -12.0, -13.5, -15.0]
SPOT_FX = 1.25


def readme_gbp_curve(interp: str = "LINEAR_ZERO_RATES"):
    """OISCurve of the README SONIA quotes (annual ACT/365F calibration swaps from the value date, MODIFIED_FOLLOWING)."""
    from .curves import OISCurve
    from .dates import BusDayAdjustTypes, Date, DayCountTypes, FrequencyTypes
    from .global_types import CurrencyTypes, CurveTypes, InterpTypes, SwapTypes
    from .trades import OIS
    vd = Date(*VALUE_DT)
    swaps = [OIS(effective_dt=vd, term_dt_or_tenor=t, fixed_leg_type=SwapTypes.PAY, fixed_coupon=px / 100,
                 fixed_freq_type=FrequencyTypes.ANNUAL, fixed_dc_type=DayCountTypes.ACT_365F,
                 floating_index=CurveTypes.GBP_OIS_SONIA, currency=CurrencyTypes.GBP,
                 bd_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, float_freq_type=FrequencyTypes.ANNUAL,
                 float_dc_type=DayCountTypes.ACT_365F) for t, px in zip(TENORS, GBP_SONIA_PX)]
    return OISCurve(vd, swaps, InterpTypes[interp])


def readme_model(with_usd: bool = False, with_basis: bool = False, interp: str = "LINEAR_ZERO_RATES"):
    """Model(value_dt) with GBP_OIS_SONIA (and optionally USD_OIS_SOFR and the GBP_USD_BASIS curve) built through
    Model.build_curve / build_xccy_curve exactly as the README does."""
    from .dates import BusDayAdjustTypes, Date, DayCountTypes, FrequencyTypes
    from .global_types import InterpTypes, SwapTypes
    from .models import Model
    m = Model(Date(*VALUE_DT))
    curves = [("GBP_OIS_SONIA", GBP_SONIA_PX, DayCountTypes.ACT_365F)]
    if with_usd or with_basis:
        curves.append(("USD_OIS_SOFR", USD_SOFR_PX, DayCountTypes.ACT_360))
    for name, px, dc in curves:
        m.build_curve(name=name, px_list=px, tenor_list=TENORS, spot_days=0, swap_type=SwapTypes.PAY, fixed_dcc_type=dc,
                      fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL, float_dc_type=dc,
                      bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING, interp_type=InterpTypes[interp])
    if with_basis:
        m.build_xccy_curve(name="GBP_USD_BASIS", domestic_curve_name="USD_OIS_SOFR", foreign_curve_name="GBP_OIS_SONIA",
                           basis_spreads=BASIS_BPS, tenor_list=BASIS_TENORS, spot_fx=SPOT_FX,
                           domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.QUARTERLY)
    return m
