import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a box without a CUDA device skips the gpu-marked tests (they need the CUDA library and a
    device: there is no CPU fallback to run them on).  An explicit `-m gpu` is never skipped: without a device it fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on a B200 box)")
        for it in gpu_items:
            it.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_curves():
    return load_golden("ref_curves.json")


@pytest.fixture(scope="session")
def ref_trades():
    return load_golden("ref_trades.json")


@pytest.fixture(scope="session")
def ref_schedules():
    return load_golden("ref_schedules.json")
