import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """Build (or reuse) the in-tree native libraries once per session.  nvcc cross-compiles
    sm_100a without a GPU; on the GPU box the prebuilt .so files are newer than the sources."""
    from navigation2_b200 import build
    build.build_all()
    yield
