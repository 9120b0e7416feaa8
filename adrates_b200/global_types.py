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


class InflationIndexTypes(Enum):
    """cavour/utils/global_types.py:93-112"""
    UK_RPI = 1
    UK_CPI = 2
    UK_CPIH = 3
    US_CPI_U = 4
    EUR_HICP = 5
    EUR_HICP_EX = 6


class InflationInterpTypes(Enum):
    """Daily CPI between monthly fixings (cavour/utils/global_types.py:113-123)"""
    FLAT = 1
    LINEAR = 2
    COMPOUND = 3


class CollateralType(Enum):
    """Collateral of a CSA (cavour/utils/global_types.py:124-148): currencies discount on OIS / XCCY curves; the bond kinds are
    names the reference reserves for later."""
    USD = 1
    GBP = 2
    EUR = 3
    JPY = 4
    CHF = 5
    AUD = 6
    CAD = 7
    USD_TIPS = 10
    EUR_OATS = 11
    EUR_BUNDS = 12
    GBP_GILTS = 13
    JGB = 14
    UNCOLLATERALIZED = 99


_BOND_COLLATERAL_CCY = {"USD_TIPS": "USD", "EUR_OATS": "EUR", "EUR_BUNDS": "EUR", "GBP_GILTS": "GBP", "JGB": "JPY"}
_OIS_CURVE_NAME = {"USD": "USD_OIS_SOFR", "GBP": "GBP_OIS_SONIA", "EUR": "EUR_OIS_ESTR", "JPY": "JPY_OIS_TONAR",
                   "CHF": "CHF_OIS_SARON", "AUD": "AUD_OIS_AONIA", "CAD": "CAD_OIS_CORRA"}


def is_currency_collateral(collateral_type: CollateralType) -> bool:
    return collateral_type in (CollateralType.USD, CollateralType.GBP, CollateralType.EUR, CollateralType.JPY,
                               CollateralType.CHF, CollateralType.AUD, CollateralType.CAD)


def is_bond_collateral(collateral_type: CollateralType) -> bool:
    return isinstance(collateral_type, CollateralType) and collateral_type.name in _BOND_COLLATERAL_CCY


def collateral_to_currency(collateral_type: CollateralType) -> CurrencyTypes:
    """Currency a collateral type is denominated in (cavour/utils/global_types.py:157-190)."""
    if is_currency_collateral(collateral_type):
        return CurrencyTypes[collateral_type.name]
    if is_bond_collateral(collateral_type):
        return CurrencyTypes[_BOND_COLLATERAL_CCY[collateral_type.name]]
    raise ValueError(f"Cannot convert {collateral_type} to currency. Use is_currency_collateral() to check first.")


def get_discount_curve_name(cashflow_currency: CurrencyTypes, collateral_type: CollateralType) -> str:
    """Name of the curve that discounts cashflows of one currency under a collateral type (global_types.py:227-300): the
    currency's OIS curve under its own collateral, `<CCY>_<COLLATERAL>_XCCY` otherwise."""
    if is_currency_collateral(collateral_type):
        collateral_ccy = collateral_to_currency(collateral_type)
        if cashflow_currency != collateral_ccy:
            return f"{cashflow_currency.name}_{collateral_ccy.name}_XCCY"
        if cashflow_currency.name not in _OIS_CURVE_NAME:
            raise ValueError(f"No OIS curve defined for {cashflow_currency}")
        return _OIS_CURVE_NAME[cashflow_currency.name]
    if is_bond_collateral(collateral_type):
        return f"{cashflow_currency.name}_{collateral_type.name}_XCCY"
    if collateral_type == CollateralType.UNCOLLATERALIZED:
        raise ValueError("Cannot generate curve name for UNCOLLATERALIZED. "
                         "Uncollateralized discounting requires separate handling.")
    raise ValueError(f"Unsupported collateral type: {collateral_type}")


ONE_MILLION = 1000000
