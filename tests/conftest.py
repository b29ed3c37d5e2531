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
    """libmetafem_b200.so must exist (built in-tree by __graft_entry__.build()); never JIT-built here."""
    import metafem_b200 as m
    if not os.path.exists(m.lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return m.lib.load()
