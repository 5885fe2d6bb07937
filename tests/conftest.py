import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle_lib import Orc
    return Orc()


@pytest.fixture(scope="session")
def ref():
    """the reference compiled into oracle/_ref (None where it was never built)"""
    from oracle_lib import Ref
    r = Ref()
    return r if r.available else None


@pytest.fixture(scope="session")
def mgb():
    import multigrid_parallel_b200 as m
    m.load_library()
    return m
