import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build (if stale) the CUDA library and the oracle once per session."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import orc_binding
    orc_binding.lib()
    return orc_binding


@pytest.fixture(scope="session")
def ctx(built):
    import pitt_object_table_segmentation_b200 as pkg
    c = pkg.Context(0)
    yield c
    c.close()
