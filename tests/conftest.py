import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """On a machine without a CUDA device the `gpu` tests are skipped (plain `pytest tests` stays a usable CPU gate).
    On a GPU box they run and fail loudly when the product library is missing: there is no CPU fallback
    (tests/test_abi.py::test_no_cpu_fallback_without_a_device covers the error path itself)."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def product_lib():
    """The C-ABI product library; GPU tests fail loudly if it is absent (no fallback)."""
    from hamers_b200 import abi

    return abi.load_library()
