import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


HAVE_GPU = None


@pytest.fixture(scope="session")
def have_gpu():
    global HAVE_GPU
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    return HAVE_GPU


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must fail loudly on a box without a GPU rather than skip silently;
    # nothing to do here -- the marker expression selects them.
    pass
