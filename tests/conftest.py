import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.lib()
    O.set_transc_mode(1)  # FP64-then-round transcendental policy shared with the CUDA product
    return O


@pytest.fixture(scope="session")
def handle():
    from wps_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()
