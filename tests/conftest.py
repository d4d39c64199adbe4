import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def capi():
    """The C-ABI library.  GPU tests go through this and fail loudly if it is missing."""
    from polychordlite_b200 import _capi
    _capi.lib()
    return _capi


@pytest.fixture(scope="session")
def gpu(capi):
    if capi.device_count() <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box (no CPU fallback exists)")
    return capi


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib
