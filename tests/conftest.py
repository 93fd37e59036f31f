import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def vor():
    from harness import abi
    return abi.backend("vor")


@pytest.fixture(scope="session")
def vref():
    from harness import abi
    if not abi.available("vref"):
        pytest.skip("oracle/_ref/libvisor_ref.so not built (needs /root/reference at build time)")
    return abi.backend("vref", 0)


@pytest.fixture(scope="session")
def gpu():
    from harness import abi
    return abi.backend("vb200", 0)
