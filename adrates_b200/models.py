"""Model: market container and shocked-copy factory.

Mirror of cavour/models/models.py: Model(value_dt), build_curve (:142-228, px in percent),
build_fx (:230-265), scenario (:507-557, shocks in percentage points: 1bp = 0.01),
`curves` accessor (:23-49, :559-572).  Bloomberg `prebuilt_*` and the cross-currency
curve builder are outside this path.
"""
from __future__ import annotations

import numpy as np
from typing import Dict, List

from .curves import OISCurve
from .dates import BusDayAdjustTypes, Date, DayCountTypes, FrequencyTypes
from .error import LibError
from .global_types import CurrencyTypes, CurveTypes, InterpTypes, SwapTypes
from .trades import OIS


class CurveAccessor:
    def __init__(self, curves: Dict[str, OISCurve]):
        self._curves = curves

    def __getattr__(self, item):
        try:
            return self.__dict__["_curves"][item]
        except KeyError:
            raise AttributeError(f"No such curve: {item}")

    def __getitem__(self, item):
        return self._curves[item]


class Model:
    def __init__(self, value_dt: Date, _curves_dict=None, _curve_params_dict=None, _fx_params_dict=None):
        self.value_dt = value_dt
        self._curves_dict: Dict[str, OISCurve] = dict(_curves_dict or {})
        self._curve_params_dict: Dict[str, dict] = dict(_curve_params_dict or {})
        self._fx_params_dict: Dict[str, dict] = dict(_fx_params_dict or {})

    def build_curve(self, name: str, px_list: List[float], tenor_list: List[str], spot_days: int = 0,
                    swap_type=SwapTypes.PAY, fixed_dcc_type=DayCountTypes.ACT_360,
                    fixed_freq_type=FrequencyTypes.ANNUAL, float_freq_type=FrequencyTypes.ANNUAL,
                    float_dc_type=DayCountTypes.ACT_360, bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                    interp_type=InterpTypes.LINEAR_ZERO_RATES, payment_lag: int = 0):
        settle_dt = self.value_dt.add_weekdays(spot_days)
        curve_type = CurveTypes[name]                    # KeyError for unknown names, as the reference
        currency = CurrencyTypes[name.split("_")[0]]
        swaps = [OIS(effective_dt=settle_dt, term_dt_or_tenor=tenor, fixed_leg_type=swap_type,
                     fixed_coupon=px / 100, fixed_freq_type=fixed_freq_type, fixed_dc_type=fixed_dcc_type,
                     floating_index=curve_type, currency=currency, bd_type=bus_day_type,
                     float_freq_type=float_freq_type, float_dc_type=float_dc_type, payment_lag=payment_lag)
                 for tenor, px in zip(tenor_list, px_list)]
        self._curves_dict[name] = OISCurve(value_dt=self.value_dt, ois_swaps=swaps, interp_type=interp_type,
                                           check_refit=True)
        self._curve_params_dict[name] = {
            "tenor_list": tenor_list, "px_list": px_list, "spot_days": spot_days, "swap_type": swap_type,
            "fixed_dcc_type": fixed_dcc_type, "fixed_freq_type": fixed_freq_type, "float_freq_type": float_freq_type,
            "float_dc_type": float_dc_type, "bus_day_type": bus_day_type, "interp_type": interp_type,
        }

    def build_fx(self, currency_pairs: List[str], pxs: List[float]):
        for pair, price in zip(currency_pairs, pxs):
            try:
                base, quote = CurrencyTypes[pair[:3]], CurrencyTypes[pair[3:]]
            except KeyError:
                raise ValueError(f"Invalid currency code in pair: {pair}")
            self._fx_params_dict[pair] = {"base": base, "quote": quote, "ticker": f"{pair} Curncy",
                                          "price": float(price)}

    def build_xccy_curve(self, name: str, domestic_curve_name: str, foreign_curve_name: str,
                         basis_spreads: List[float], tenor_list: List[str], spot_fx: float,
                         domestic_notional: float = 100_000_000,
                         domestic_freq_type=FrequencyTypes.ANNUAL, foreign_freq_type=FrequencyTypes.ANNUAL,
                         domestic_dc_type=DayCountTypes.ACT_360, foreign_dc_type=DayCountTypes.ACT_365F,
                         bus_day_type=BusDayAdjustTypes.MODIFIED_FOLLOWING,
                         interp_type=InterpTypes.FLAT_FWD_RATES, use_ad: bool = True):
        """Cross-currency basis curve from basis spreads in bp (models.py:267-391).  As in the reference the
        calibration swaps use the XccyBasisSwap default business-day rule and XccyCurve receives 1/spot_fx."""
        from .trades import XccyBasisSwap
        from .xccy_curve import XccyCurve
        for nm, what in ((domestic_curve_name, "Domestic"), (foreign_curve_name, "Foreign")):
            if nm not in self._curves_dict:
                raise ValueError(f"{what} curve '{nm}' not found in model. Build it first using build_curve() or "
                                 f"prebuilt_curve().")
        dom_ccy = CurrencyTypes[domestic_curve_name.split("_")[0]]
        for_ccy = CurrencyTypes[foreign_curve_name.split("_")[0]]
        foreign_notional = domestic_notional / spot_fx
        swaps = [XccyBasisSwap(effective_dt=self.value_dt, term_dt_or_tenor=tenor, domestic_notional=domestic_notional,
                               foreign_notional=foreign_notional, domestic_spread=0.0, foreign_spread=bps / 10000.0,
                               domestic_freq_type=domestic_freq_type, foreign_freq_type=foreign_freq_type,
                               domestic_dc_type=domestic_dc_type, foreign_dc_type=foreign_dc_type,
                               domestic_floating_index=CurveTypes[domestic_curve_name],
                               foreign_floating_index=CurveTypes[foreign_curve_name],
                               domestic_currency=dom_ccy, foreign_currency=for_ccy)
                 for tenor, bps in zip(tenor_list, basis_spreads)]
        self._curves_dict[name] = XccyCurve(value_dt=self.value_dt, basis_swaps=swaps,
                                            domestic_curve=self._curves_dict[domestic_curve_name],
                                            foreign_curve=self._curves_dict[foreign_curve_name],
                                            spot_fx=1 / spot_fx, interp_type=interp_type, use_ad=use_ad)
        self._curve_params_dict[name] = {
            "domestic_curve_name": domestic_curve_name, "foreign_curve_name": foreign_curve_name,
            "basis_spreads": basis_spreads, "tenor_list": tenor_list, "spot_fx": spot_fx,
            "domestic_notional": domestic_notional, "domestic_freq_type": domestic_freq_type,
            "foreign_freq_type": foreign_freq_type, "domestic_dc_type": domestic_dc_type,
            "foreign_dc_type": foreign_dc_type, "bus_day_type": bus_day_type, "interp_type": interp_type,
            "use_ad": use_ad,
        }

    def scenario(self, curve_name: str, shock, new_name: str = None) -> "Model":
        """New Model holding `curve_name` rebuilt from shocked quotes (float = parallel,
        dict = tenor -> shock; percentage points) (models.py:507-557)."""
        if curve_name not in self._curve_params_dict:
            raise ValueError(f"No stored parameters found for curve '{curve_name}'")
        params = dict(self._curve_params_dict[curve_name])
        tenors, px = params.pop("tenor_list"), params.pop("px_list")
        if isinstance(shock, dict):
            shocked = [p + shock.get(t, 0.0) for t, p in zip(tenors, px)]
        else:
            shocked = [p + shock for p in px]
        new = Model(self.value_dt)   # like the reference, the copy holds only the shocked curve
        new.build_curve(name=new_name or curve_name, px_list=shocked, tenor_list=tenors, **params)
        return new

    def scenario_rates(self, curve_name: str, shocks) -> np.ndarray:
        """Par rates (decimal) of `curve_name` under a batch of shocks: row s is what
        `self.scenario(curve_name, shocks[s]).curves[curve_name].swap_rates` would hold, without rebuilding S curves on
        the host (models.py:507-557 applied S times).  `shocks`: a sequence whose items are floats (parallel) or
        dicts tenor -> shock, or an array [S] (parallel) / [S, R] (per pillar, quote order); percentage points
        (1bp = 0.01).  Input of the batched revaluation (`OISBook.scenario_values`, `Portfolio.scenario_values`)."""
        if curve_name not in self._curve_params_dict:
            raise ValueError(f"No stored parameters found for curve '{curve_name}'")
        params = self._curve_params_dict[curve_name]
        tenors, px = list(params["tenor_list"]), np.asarray(params["px_list"], dtype=np.float64)
        if isinstance(shocks, np.ndarray) and shocks.dtype != object:
            sh = np.asarray(shocks, dtype=np.float64)
            if sh.ndim == 1:
                sh = sh[:, None]
            if sh.ndim != 2 or sh.shape[1] not in (1, len(px)):
                raise ValueError(f"shocks must be [S] or [S, {len(px)}], got {shocks.shape}")
            shocked = px[None, :] + sh
        else:
            shocked = np.empty((len(shocks), len(px)))
            for s, shock in enumerate(shocks):
                if isinstance(shock, dict):
                    shocked[s] = [p + shock.get(t, 0.0) for t, p in zip(tenors, px)]
                else:
                    shocked[s] = px + float(shock)
        return shocked / 100

    @property
    def curves(self) -> CurveAccessor:
        return CurveAccessor(self._curves_dict)
