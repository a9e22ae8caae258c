import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """libspe.so built in-tree; building is part of the fixture so a missing/stale .so fails loudly."""
    from satellite_pose_estimation_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device (there is no CPU fallback to test)"
    return "cuda:0"
