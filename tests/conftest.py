import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the kernel library; GPU tests fail loudly when it is missing."""
    import torch

    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from univst_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    _lib.require_device()
    return _lib.lib()
