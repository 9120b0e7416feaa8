"""adrates_b200 - B200-native valuation-and-Greeks path behind the Cavour/ADRates API.

Host side: Python mirror of the reference interface for this path (Model.build_curve,
curve.df_ad, OIS(...).position(model).compute([VALUE, DELTA, GAMMA]), Portfolio.compute).
Device side: hand-written sm_100a CUDA kernels behind the C ABI in include/adrates_b200.h
(adrates_b200/csrc), loaded with ctypes.  There is no CPU fallback for valuation.
"""
from .error import LibError
from .dates import (Date, DateFormatTypes, set_date_format, Calendar, CalendarTypes, create_calendar_intersection,
                    BusDayAdjustTypes, DateGenRuleTypes, DayCount, DayCountTypes, FrequencyTypes, Schedule, to_tenor,
                    times_from_dates)
from .global_types import (SwapTypes, InstrumentTypes, RequestTypes, InterpTypes, CurveTypes,
                           CurrencyTypes, CollateralType)

from .trades import OIS, SwapFixedLeg, SwapFloatLeg, XccyBasisSwap, XccyFixFloat, XccyFixFix
from .xccy_curve import XccyCurve
from .curves import OISCurve, DiscountCurve
from .models import Model
from .position import Position, Portfolio, Engine
from .results import Valuation, Delta, Gamma, CrossGamma, Risk, AnalyticsResult
from .cashflows import CashflowItem, Cashflows
from .credit import Bond, FRN
from .inflation import (InflationIndex, InflationCurve, InflationIndexTypes, InflationInterpTypes, SwapInflationLeg,
                        ZeroCouponInflationSwap, SwapYoYInflationLeg, YoYInflationSwap)

__all__ = [
    "OIS", "SwapFixedLeg", "SwapFloatLeg", "XccyBasisSwap", "XccyFixFloat", "XccyFixFix", "XccyCurve", "OISCurve", "DiscountCurve", "Model", "Position", "Portfolio", "Engine",
    "Valuation", "Delta", "Gamma", "CrossGamma", "Risk", "AnalyticsResult", "CashflowItem", "Cashflows", "InflationIndex", "InflationCurve", "InflationIndexTypes",
    "InflationInterpTypes", "SwapInflationLeg", "ZeroCouponInflationSwap", "SwapYoYInflationLeg", "YoYInflationSwap", "Bond", "FRN",
    "LibError", "Date", "DateFormatTypes", "set_date_format", "Calendar", "CalendarTypes", "create_calendar_intersection", "BusDayAdjustTypes", "DateGenRuleTypes", "DayCount",
    "DayCountTypes", "FrequencyTypes", "Schedule", "to_tenor", "times_from_dates", "SwapTypes",
    "InstrumentTypes", "RequestTypes", "InterpTypes", "CurveTypes", "CurrencyTypes", "CollateralType",
]
