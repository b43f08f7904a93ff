import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


# Test hooks of the sort's size thresholds (radix_sort.cu, SortEnv): the product takes the speculative kernels from 2^24 keys
# and the warp-specialised kernel from 2^27; the parity tests compare against the CPU oracle, so they move the thresholds
# down to 2^20 / 2^23 to push every kernel through oracle-sized inputs.  (Read once, when the library first sorts.)
os.environ.setdefault("BCB_SORT_SPEC_MIN_LOG2", "20")
os.environ.setdefault("BCB_SORT_WS_MIN_LOG2", "23")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
