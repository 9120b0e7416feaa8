#!/usr/bin/env python
"""Runs the REFERENCE'S OWN test files (/root/reference/tests/test_*.py, unmodified, read in place) against this package:
the module names the tests import (`cavour.utils.date`, `cavour.trades.rates.ois`, `cavour.models.models` ...) are aliased to
the modules of adrates_b200 that mirror them, so every `from cavour... import X` resolves to OUR class.  Build container only
(/root/reference does not exist on the GPU box; nothing in tests/, smoke() or bench.py uses this).  TEST INFRASTRUCTURE.

    python tools/reftests/run_reference_tests.py [test_date_arithmetic.py ...]      # default: every reference test file

Tests that reach the device (`Position.compute`, `df_ad`) fail loudly here - there is no CPU fallback - and are listed as such;
the committed summary is tools/reftests/REPORT.md.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHIM = os.path.join(ROOT, "tests", "golden", "gen", "refshim")      # `import jax.numpy as jnp` of the test files themselves
REF_TESTS = "/root/reference/tests"

# reference module -> the modules of this package whose public names stand in for it (first one wins on a clash)
ALIASES = {
    "cavour.utils.date": ["adrates_b200.dates"],
    "cavour.utils.calendar": ["adrates_b200.dates"],
    "cavour.utils.day_count": ["adrates_b200.dates"],
    "cavour.utils.frequency": ["adrates_b200.dates"],
    "cavour.utils.schedule": ["adrates_b200.dates"],
    "cavour.utils.helpers": ["adrates_b200.dates"],
    "cavour.utils.error": ["adrates_b200.error"],
    "cavour.utils.currency": ["adrates_b200.global_types"],
    "cavour.utils.global_types": ["adrates_b200.global_types", "adrates_b200.inflation"],
    "cavour.market.curves.interpolator": ["adrates_b200.interpolator", "adrates_b200.global_types"],
    "cavour.market.curves.discount_curve": ["adrates_b200.curves"],
    "cavour.market.curves.inflation_curve": ["adrates_b200.inflation"],
    "cavour.market.indices.inflation_index": ["adrates_b200.inflation"],
    "cavour.market.position.engine": ["adrates_b200.position"],
    "cavour.market.position.position": ["adrates_b200.position"],
    "cavour.market.portfolio.portfolio": ["adrates_b200.position"],
    "cavour.requests.results": ["adrates_b200.results", "adrates_b200.cashflows"],
    "cavour.models.models": ["adrates_b200.models"],
    "cavour.trades.rates.ois": ["adrates_b200.trades"],
    "cavour.trades.rates.ois_curve": ["adrates_b200.curves"],
    "cavour.trades.rates.swap_fixed_leg": ["adrates_b200.trades"],
    "cavour.trades.rates.swap_float_leg": ["adrates_b200.trades"],
    "cavour.trades.rates.xccy_basis_swap": ["adrates_b200.trades"],
    "cavour.trades.rates.xccy_fix_float_swap": ["adrates_b200.trades"],
    "cavour.trades.rates.xccy_fix_fix_swap": ["adrates_b200.trades"],
    "cavour.trades.rates.xccy_curve": ["adrates_b200.xccy_curve"],
    "cavour.trades.rates.zcis": ["adrates_b200.inflation"],
    "cavour.trades.rates.swap_inflation_leg": ["adrates_b200.inflation"],
    "cavour.trades.rates.yoy_inflation_swap": ["adrates_b200.inflation"],
    "cavour.trades.rates.swap_yoy_inflation_leg": ["adrates_b200.inflation"],
    "cavour.trades.credit.bond": ["adrates_b200.credit"],
    "cavour.trades.credit.frn": ["adrates_b200.credit"],
}


def install_aliases():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, SHIM)
    made = {}

    def package(name):
        if name not in made:
            m = types.ModuleType(name)
            m.__path__ = []                # a package with nothing on disk: un-aliased submodules raise ModuleNotFoundError
            sys.modules[name] = made[name] = m
            if "." in name:
                parent, _, leaf = name.rpartition(".")
                setattr(package(parent), leaf, m)
        return made[name]

    for ref_name, ours in ALIASES.items():
        parent, _, leaf = ref_name.rpartition(".")
        m = types.ModuleType(ref_name)
        for mod_name in reversed(ours):
            src = importlib.import_module(mod_name)
            for k, v in vars(src).items():
                if not k.startswith("__"):
                    setattr(m, k, v)
        sys.modules[ref_name] = m
        setattr(package(parent), leaf, m)


def main():
    install_aliases()
    import pytest
    files = sys.argv[1:] or sorted(f for f in os.listdir(REF_TESTS) if f.startswith("test_") and f.endswith(".py"))
    args = [os.path.join(REF_TESTS, f) for f in files]
    # the reference tree is read-only: no cache, no bytecode; its own conftest.py (fixtures) is used as it is
    sys.dont_write_bytecode = True
    return pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS, "-c", "/dev/null", "--tb=short",
                        "-o", "addopts=", *args])


if __name__ == "__main__":
    sys.exit(main())
