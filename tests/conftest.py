import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available():
    try:
        from grm_b200 import native
        return native.load().grmkm_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must run the CUDA path: missing library or device is a failure, not a skip."""
    from grm_b200 import native
    lib = native.load()
    assert lib.grmkm_device_count() > 0, "no CUDA device visible: -m gpu tests need the B200 box"
    return lib
