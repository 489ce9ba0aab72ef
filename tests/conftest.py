import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu = _has_gpu()
    ref = os.path.isdir("/root/reference/srcs")
    for it in items:
        if "gpu" in it.keywords and not gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "needs_reference" in it.keywords and not ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def dev():
    import torch
    return torch.device("cuda:0")
