import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build(ref=os.path.isdir("/root/reference"))
    return oracle


@pytest.fixture(scope="session")
def mavg():
    """The product package with libmavg.so built (nvcc cross-compiles without a GPU)."""
    from digital_signal_processsing_b200 import build
    build.build_lib()
    import digital_signal_processsing_b200 as m
    return m
