"""Result containers returned by Position.compute / Portfolio.compute.

Field names and semantics follow cavour/requests/results.py: Valuation (:37-164),
Delta (:228-380: `risk_ladder[R]`, `tenors`, `.value` = sum, `.ladder` tenor->value dict with
the reference's label collisions), Gamma (:383-605: `risk_ladder[R,R]`, `.value` = sum of
all entries), Risk (:839-942) and AnalyticsResult (:1124-1202).  Arrays are numpy float64.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional

import numpy as np

from .global_types import CurrencyTypes, CurveTypes


class _Export:
    """to_csv / to_excel over the container's `df` (the reference's export helpers, results.py:133-158 and per class)."""
    _sheet = "Sheet1"

    def to_csv(self, filepath: Optional[str] = None) -> Optional[str]:
        if filepath:
            self.df.to_csv(filepath)
            return None
        return self.df.to_csv()

    def to_excel(self, filepath: str, sheet_name: Optional[str] = None):
        self.df.to_excel(filepath, sheet_name=sheet_name or self._sheet)


@dataclass(frozen=True)
class Valuation(_Export):
    _sheet = "Valuation"

    amount: float
    currency: CurrencyTypes = CurrencyTypes.NONE

    def __post_init__(self):
        if not isinstance(self.currency, CurrencyTypes):
            raise TypeError(f"currency must be a CurrencyTypes enum, got {type(self.currency)}")

    def __repr__(self):
        return f"{self.amount:.2f} {self.currency.name}"

    def _same(self, other, verb):
        if self.currency is not other.currency:
            raise ValueError(f"Cannot {verb} {self.currency.name} and {other.currency.name}")

    def __add__(self, other: Any):
        if not isinstance(other, Valuation):
            return NotImplemented
        self._same(other, "add")
        return Valuation(self.amount + other.amount, self.currency)

    def __radd__(self, other: Any):
        return self if other == 0 else self.__add__(other)

    def __sub__(self, other: Any):
        if not isinstance(other, Valuation):
            return NotImplemented
        self._same(other, "subtract")
        return Valuation(self.amount - other.amount, self.currency)

    def __mul__(self, factor: float):
        return Valuation(self.amount * factor, self.currency)

    __rmul__ = __mul__

    def __truediv__(self, divisor: float):
        return Valuation(self.amount / divisor, self.currency)

    def to_dict(self):
        return {"amount": float(self.amount), "currency": self.currency.name}

    def to_json(self, indent: Optional[int] = 2) -> str:
        import json
        return json.dumps(self.to_dict(), indent=indent)

    @property
    def df(self):
        import pandas as pd
        return pd.DataFrame([self.to_dict()])


Value = Valuation


class Ladder:
    def __init__(self, data: dict, curve_name: str):
        self.data = data
        self._curve_name = curve_name

    @property
    def df(self):
        """One column `<curve>_Risk` indexed by tenor label (results.py:204-217)"""
        import pandas as pd
        out = pd.DataFrame.from_dict(self.data, orient="index", columns=[f"{self._curve_name}_Risk"])
        out.index.name = "Tenor"
        return out

    def to_dict(self) -> dict:
        return dict(self.data)

    def __repr__(self):
        return f"Ladder(curve={self._curve_name}, points={len(self.data)}, curve_data={self.data})"


class _Sensitivity(_Export):
    risk_ladder: np.ndarray
    tenors: List[str]
    currency: CurrencyTypes
    curve_type: CurveTypes

    def _validate(self):
        arr = np.asarray(self.risk_ladder, dtype=np.float64)
        object.__setattr__(self, "risk_ladder", arr)
        if arr.shape[-1] != len(self.tenors):
            raise ValueError(f"Expected {arr.shape[-1]} tenors, got {len(self.tenors)}")
        if not isinstance(self.currency, CurrencyTypes):
            raise TypeError(f"currency must be CurrencyTypes, got {type(self.currency)}")
        if not isinstance(self.curve_type, CurveTypes):
            raise TypeError(f"curve_type must be CurveTypes, got {type(self.curve_type)}")

    @property
    def value(self) -> Valuation:
        return Valuation(float(np.sum(self.risk_ladder)), self.currency)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.curve_type.name}: {self.value.amount:.6g} "
                f"{self.currency.name}, points={len(self.tenors)})")

    def __add__(self, other: Any):
        if not isinstance(other, self.__class__):
            return NotImplemented
        if (self.curve_type != other.curve_type or self.currency != other.currency or self.tenors != other.tenors):
            raise ValueError(f"Cannot add {self.__class__.__name__} with mismatched curve_type, currency, or tenors")
        return self.__class__(self.risk_ladder + other.risk_ladder, self.tenors, self.currency, self.curve_type)

    def __radd__(self, other: Any):
        return self if other == 0 else self.__add__(other)


@dataclass(frozen=True, repr=False)
class Delta(_Sensitivity):
    """Per-bp ladder: risk_ladder[k] = 1e-4 * dPV/d(par rate k) (engine.py:2554-2555)."""
    risk_ladder: np.ndarray
    tenors: List[str]
    currency: CurrencyTypes
    curve_type: CurveTypes

    def __post_init__(self):
        self._validate()

    @property
    def ladder(self) -> Ladder:
        return Ladder(dict(zip(self.tenors, self.risk_ladder.tolist())), self.curve_type.name)

    def to_dict(self):
        return {"risk_ladder": self.risk_ladder.tolist(), "tenors": self.tenors, "currency": self.currency.name,
                "curve_type": self.curve_type.name, "total": float(np.sum(self.risk_ladder))}

    def to_json(self, indent: Optional[int] = 2) -> str:
        import json
        return json.dumps(self.to_dict(), indent=indent)

    _sheet = "Delta"

    @property
    def df(self):
        return self.ladder.df


@dataclass(frozen=True, repr=False)
class Gamma(_Sensitivity):
    """Per-bp^2 matrix: risk_ladder[j,k] = 1e-8 * d2PV/d(rate j)d(rate k) (engine.py:2565-2568)."""
    risk_ladder: np.ndarray
    tenors: List[str]
    currency: CurrencyTypes
    curve_type: CurveTypes

    def __post_init__(self):
        self._validate()

    @property
    def to_dict(self) -> dict:
        g = np.asarray(self.risk_ladder)
        if g.ndim != 2:
            raise ValueError("Gamma risk_ladder must be 2D to access matrix")
        return {rt: {ct: float(g[i, j]) for j, ct in enumerate(self.tenors)} for i, rt in enumerate(self.tenors)}

    _sheet = "Gamma"

    @property
    def df(self):
        """Matrix indexed by tenor labels on both axes; a 1-D ladder becomes the diagonal (results.py:598-605)"""
        import pandas as pd
        g = np.asarray(self.risk_ladder)
        return pd.DataFrame(np.diag(g) if g.ndim == 1 else g, index=self.tenors, columns=self.tenors)

    @property
    def matrix(self) -> None:
        """Prints the matrix without its all-zero rows and columns as a grid table (results.py:443-462).  Built from the
        tenor -> {tenor -> value} dictionary, so - as in the reference - pillars that share a tenor label collapse."""
        import pandas as pd
        from tabulate import tabulate
        out = pd.DataFrame(self.to_dict)
        out = out.loc[~(out == 0).all(axis=1)]
        out = out.loc[:, ~(out == 0).all(axis=0)]
        out.index = [f"{float(i):.2f}" if _is_number(i) else i for i in out.index]
        out.columns = [f"{float(c):.2f}" if _is_number(c) else c for c in out.columns]
        out.index.name = "Tenors"
        print(tabulate(out, headers="keys", tablefmt="grid", floatfmt=".2f"))

    def to_json(self, indent: Optional[int] = 2) -> str:
        import json
        return json.dumps({"matrix": self.to_dict, "tenors": self.tenors, "currency": self.currency.name,
                           "curve_type": self.curve_type.name, "total": float(np.sum(self.risk_ladder))}, indent=indent)


def _is_number(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating))


@dataclass(frozen=True, repr=False)
class CrossGamma(_Export):
    """Cross-curve second-order sensitivity: risk_matrix[i, j] = 1e-8 * d2PV / d(curve-1 rate i) d(curve-2 rate j)
    (cavour/requests/results.py:608-836: value, to_dict, df, addition, JSON / CSV export; plotting helpers are not
    part of the valuation path)."""
    risk_matrix: np.ndarray
    tenors_curve1: List[str]
    tenors_curve2: List[str]
    curve_type_1: CurveTypes
    curve_type_2: CurveTypes
    currency: CurrencyTypes

    def __post_init__(self):
        arr = np.asarray(self.risk_matrix, dtype=np.float64)
        object.__setattr__(self, "risk_matrix", arr)
        if arr.ndim != 2:
            raise ValueError(f"CrossGamma risk_matrix must be 2D, got {arr.ndim}D")
        n1, n2 = arr.shape
        if n1 != len(self.tenors_curve1):
            raise ValueError(f"Expected {n1} tenors for curve 1, got {len(self.tenors_curve1)}")
        if n2 != len(self.tenors_curve2):
            raise ValueError(f"Expected {n2} tenors for curve 2, got {len(self.tenors_curve2)}")
        if not isinstance(self.currency, CurrencyTypes):
            raise TypeError(f"currency must be CurrencyTypes, got {type(self.currency)}")
        if not isinstance(self.curve_type_1, CurveTypes):
            raise TypeError(f"curve_type_1 must be CurveTypes, got {type(self.curve_type_1)}")
        if not isinstance(self.curve_type_2, CurveTypes):
            raise TypeError(f"curve_type_2 must be CurveTypes, got {type(self.curve_type_2)}")

    @property
    def value(self) -> Valuation:
        return Valuation(float(np.sum(self.risk_matrix)), self.currency)

    @property
    def to_dict(self) -> dict:
        g = self.risk_matrix
        return {rt: {ct: float(g[i, j]) for j, ct in enumerate(self.tenors_curve2)} for i, rt in enumerate(self.tenors_curve1)}

    @property
    def df(self):
        import pandas as pd
        out = pd.DataFrame(self.risk_matrix, index=self.tenors_curve1, columns=self.tenors_curve2)
        out.index.name = f"{self.curve_type_1.name}_Tenors"
        out.columns.name = f"{self.curve_type_2.name}_Tenors"
        return out

    def to_json(self, indent: Optional[int] = 2) -> str:
        import json
        return json.dumps({"matrix": self.to_dict, "tenors_curve1": self.tenors_curve1, "tenors_curve2": self.tenors_curve2,
                           "curve_type_1": self.curve_type_1.name, "curve_type_2": self.curve_type_2.name,
                           "currency": self.currency.name, "total": float(np.sum(self.risk_matrix))}, indent=indent)

    _sheet = "CrossGamma"

    @property
    def matrix(self) -> None:
        """Prints the full matrix as a grid table (results.py:683-697)"""
        import pandas as pd
        from tabulate import tabulate
        out = pd.DataFrame(self.to_dict)
        out.index.name = f"{self.curve_type_1.name} Tenors"
        out.columns.name = f"{self.curve_type_2.name} Tenors"
        print(tabulate(out, headers="keys", tablefmt="grid", floatfmt=".4f"))

    def __add__(self, other: Any) -> "CrossGamma":
        if not isinstance(other, CrossGamma):
            return NotImplemented
        if (self.curve_type_1 != other.curve_type_1 or self.curve_type_2 != other.curve_type_2 or self.currency != other.currency
                or self.tenors_curve1 != other.tenors_curve1 or self.tenors_curve2 != other.tenors_curve2):
            raise ValueError("Cannot add CrossGamma with mismatched curves, currency, or tenors")
        return CrossGamma(self.risk_matrix + other.risk_matrix, self.tenors_curve1, self.tenors_curve2, self.curve_type_1,
                          self.curve_type_2, self.currency)

    __radd__ = __add__

    def __repr__(self):
        n1, n2 = len(self.tenors_curve1), len(self.tenors_curve2)
        return (f"CrossGamma({self.curve_type_1.name} x {self.curve_type_2.name}: {self.value.amount:.6g} "
                f"{self.currency.name}, shape=[{n1}, {n2}])")


class Risk:
    """Multi-curve container: attribute access by curve name, call with a CurveTypes
    (cavour/requests/results.py:839-942)."""

    def __init__(self, ladders, cross_gammas=None):
        self._by_curve = {}
        for it in ladders:
            name = it.curve_type.name
            if name in self._by_curve:
                raise ValueError(f"Duplicate curve {name} in Risk")
            self._by_curve[name] = it
        self._cross_gammas = {}
        for cg in cross_gammas or []:
            key = (cg.curve_type_1.name, cg.curve_type_2.name)
            if key in self._cross_gammas:
                raise ValueError(f"Duplicate cross-gamma for {key}")
            self._cross_gammas[key] = cg

    def __call__(self, curve_type: CurveTypes):
        try:
            return self._by_curve[curve_type.name]
        except KeyError:
            raise ValueError(f"No risk data for curve: {curve_type.name}")

    def __getattr__(self, name):
        try:
            return self.__dict__["_by_curve"][name]
        except KeyError:
            raise AttributeError(f"No risk for curve {name}")

    def cross_gamma(self, curve_type_1: CurveTypes, curve_type_2: CurveTypes):
        return self._cross_gammas.get((curve_type_1.name, curve_type_2.name), None)

    def has_cross_gamma(self, curve_type_1: CurveTypes, curve_type_2: CurveTypes) -> bool:
        return (curve_type_1.name, curve_type_2.name) in self._cross_gammas

    @property
    def all_cross_gammas(self) -> dict:
        return self._cross_gammas.copy()

    def __repr__(self):
        """Risk(<curve>=<total> <ccy>, ...) - results.py:937-942"""
        parts = [f"{k}={v.value.amount:.6g} {v.value.currency.name}" for k, v in self._by_curve.items()]
        return f"{self.__class__.__name__}({', '.join(parts)})"


class AnalyticsResult:
    def __init__(self, value: Optional[Valuation] = None, risk=None, gamma=None, cashflows=None):
        self._value, self._risk, self._gamma, self._cashflows = value, risk, gamma, cashflows

    @property
    def value(self):
        return self._value

    @property
    def risk(self):
        return self._risk

    @property
    def gamma(self):
        return self._gamma

    @property
    def cashflows(self):
        return self._cashflows

    def __repr__(self):
        parts = [f"{k}={v!r}" for k, v in (("value", self._value), ("risk", self._risk), ("gamma", self._gamma),
                                           ("cashflows", self._cashflows)) if v is not None]
        return f"AnalyticsResult({', '.join(parts)})"
