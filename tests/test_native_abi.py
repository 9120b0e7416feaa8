"""The C-ABI library builds, loads and exports every symbol include/*.h declares (no
compute calls: there is no GPU in the CPU test run)."""
import ctypes
import os
import re

import pytest

from adrates_b200 import _native, build as b
from adrates_b200.error import LibError
from tests.conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "adrates_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cav_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    lib = b.build()
    assert os.path.exists(lib)
    dll = ctypes.CDLL(lib)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(dll, s), s
    assert sorted(_native.EXPORTS) == syms
    dll.cav_version.restype = ctypes.c_int
    assert dll.cav_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(LibError):
        _native.Context(0)


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", b.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
