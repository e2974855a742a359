import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    if not o.available() and not o.build():
        pytest.skip("oracle library not built and /root/reference not mounted")
    return o


@pytest.fixture(scope="session")
def cfx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import cuttlefish_b200 as c
    c.init(0)
    return c
