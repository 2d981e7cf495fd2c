import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on a B200")


@pytest.fixture(scope="session")
def crux():
    import crux_b200
    return crux_b200


@pytest.fixture(scope="session")
def ctx(crux):
    """One device context for the whole GPU session (fails loudly without a GPU)."""
    return crux.default_context()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _release_device_temporaries():
    yield
    try:
        import gpu_util
        if gpu_util._KEEP:
            import torch
            torch.cuda.synchronize()
            gpu_util._KEEP.clear()
    except ImportError:
        pass
