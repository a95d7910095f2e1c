import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (this container only)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the CUDA library; GPU tests fail loudly without it."""
    from pydem_b200 import build, _lib
    build.build()      # no-op when the library is newer than its sources
    _lib.load()
    _lib.init(0)
    return _lib
