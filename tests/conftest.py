import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        from copra_b200 import capi
        return capi.load().copra_b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    from copra_b200 import capi
    if not _has_gpu():
        pytest.fail("GPU test selected but no CUDA device / libcopra_b200.so: the product has no CPU fallback")
    eng = capi.Engine(0)
    yield eng
    eng.close()
