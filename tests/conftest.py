import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (g++) and, when nvcc is present, the CUDA library."""
    from oracle import oracle as orc
    orc.build()
    from regcm_b200 import build as B
    try:
        B.build_library()
        B.build_fast()
    except RuntimeError:
        if not os.path.exists(B.LIB):
            raise
    yield
