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
    """Make sure the checker library and libgndt.so exist (compiled in-tree)."""
    from grid_ndt_b200._lib import LIB_PATH, build_library
    from oracle import oracle as O
    O.build()
    if not os.path.exists(LIB_PATH):
        build_library()
    yield


def have_reference_lib():
    from oracle import oracle as O
    return os.path.exists(O.REF_SO)
