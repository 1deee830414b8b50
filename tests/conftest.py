import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import pkgload
    return pkgload.load()


@pytest.fixture(scope="session")
def built():
    """Native libraries are built once per session (no-op when up to date)."""
    import __graft_entry__ as ge
    ge.build()
    return True
