import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
