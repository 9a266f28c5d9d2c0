import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_available():
    try:
        import ugemm_b200 as u
        u.sgemm_cuda_init()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def u():
    """The product library, initialised.  GPU tests FAIL (not skip) when the CUDA path is unavailable."""
    import ugemm_b200 as mod
    mod.sgemm_cuda_init()
    return mod
