import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from emlight_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from emlight_b200 import _lib
    _lib.check(lib.eml_device_ok(), "eml_device_ok")
    return torch.device("cuda:0")
