import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def mmx_lib():
    """The built C-ABI library (compiled on demand with nvcc; works without a GPU)."""
    from micromix_b200 import build as B
    B.build()
    from micromix_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a CUDA device is required for -m gpu tests (no CPU fallback exists)")
    return torch.device("cuda:0")
