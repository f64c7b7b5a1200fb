import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle.build import build
    build()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guassianhand_b200 import _native, build
    # a stale binary would pass or fail against old kernels: libghr.so must be newer than every source
    # (the box has the sources and the prebuilt .so of the same snapshot; nvcc is there too)
    if build._stale():
        build.build()
    _native.lib()     # raises if libghr.so is missing: the GPU tests must never pass on a fallback
    return "cuda:0"
