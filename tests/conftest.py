import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests._util import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own CUDA code (oracle/_ref/libnmref.so); GPU tests only."""
    from tests._util import load_reflib
    lib = load_reflib()
    if lib is None:
        pytest.skip("oracle/_ref/libnmref.so not built")
    return lib
