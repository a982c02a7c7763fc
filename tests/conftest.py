import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a) and the built libavdn.so")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libavdn.so exists (nvcc cross-compiles on a CPU-only box)."""
    import __graft_entry__ as g
    from avdn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.LIB_PATH
