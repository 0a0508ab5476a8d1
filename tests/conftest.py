import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from graphbix_b200 import _lib
        return _lib.load().gbx_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    """GPU tests never skip silently on a GPU box: the library must load and see a device."""
    from graphbix_b200 import _lib
    n = _lib.load().gbx_device_count()
    assert n > 0, "no CUDA device visible to libgraphbix_cuda.so"
    return n
