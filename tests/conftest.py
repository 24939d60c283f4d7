import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the in-tree shared library; CPU-side tests only load it / inspect symbols."""
    import __graft_entry__ as g
    g.build(only_if_missing=True)
    return True
