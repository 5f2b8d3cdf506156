import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds (once) and loads the CPU oracle -- the checker, never the product."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def product_lib():
    """The C-ABI library under test; built by __graft_entry__.build()."""
    import __graft_entry__ as g
    from mrhyde_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        g.build()
    return capi
