import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (no-op when up to date) and load the CUDA library."""
    from paillier_halo2_b200 import build as b
    from paillier_halo2_b200 import _lib

    b.build()
    return _lib.load()
