"""Enumerations shared by the valuation path.

Names and numeric values follow cavour/utils/global_types.py:53-155 and
cavour/utils/currency.py so that user code and serialised enums are interchangeable.
"""
from enum import Enum


class SwapTypes(Enum):
    PAY = 1
    RECEIVE = 2


class InstrumentTypes(Enum):
    SWAP_FIXED_LEG = 1
    SWAP_FLOAT_LEG = 2
    OIS_SWAP = 3
    XCCY_SWAP = 4
    ZCIS = 5
    SWAP_INFLATION_LEG = 6
    BOND = 7
    FRN = 8
    YOY_INFLATION_SWAP = 9
    SWAP_YOY_INFLATION_LEG = 10


class RequestTypes(Enum):
    VALUE = 1
    DELTA = 2
    GAMMA = 3
    SPEED = 4
    CASHFLOWS = 5


class InterpTypes(Enum):
    FLAT_FWD_RATES = 1
    LINEAR_FWD_RATES = 2
    LINEAR_ZERO_RATES = 4
    FINCUBIC_ZERO_RATES = 7
    NATCUBIC_LOG_DISCOUNT = 8
    NATCUBIC_ZERO_RATES = 9
    PCHIP_ZERO_RATES = 10
    PCHIP_LOG_DISCOUNT = 11


class CurveTypes(Enum):
    GBP_OIS_SONIA = 1
    USD_OIS_SOFR = 2
    EUR_OIS_ESTR = 3
    USD_GBP_BASIS = 4
    GBP_RPI_INFLATION = 5
    GBP_CPI_INFLATION = 6
    USD_CPI_INFLATION = 7
    EUR_HICP_INFLATION = 8


class CurrencyTypes(Enum):
    USD = 1
    EUR = 2
    GBP = 3
    CHF = 4
    CAD = 5
    AUD = 6
    NZD = 7
    DKK = 8
    SEK = 9
    HKD = 10
    JPY = 11
    NOK = 12
    PLN = 13
    RON = 14
    NONE = 15


class CollateralType(Enum):
    USD = 1
    GBP = 2
    EUR = 3
    JPY = 4
    CHF = 5
    AUD = 6
    CAD = 7
    UNCOLLATERALIZED = 99


def collateral_to_currency(collateral_type: CollateralType) -> CurrencyTypes:
    """cavour/utils/global_types.py:157-190 (currency collateral only)."""
    if collateral_type == CollateralType.UNCOLLATERALIZED or collateral_type.name not in CurrencyTypes.__members__:
        raise ValueError(f"Cannot convert {collateral_type} to currency. Use is_currency_collateral() to check first.")
    return CurrencyTypes[collateral_type.name]


ONE_MILLION = 1000000
