import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle_py import build, lib
    build()
    return lib()


@pytest.fixture(scope="session")
def built_cuda_lib():
    from cosmomc_galileon_b200 import build as b
    return b.build()


def _cuda_device_present():
    try:
        import ctypes
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcudart.so")
        return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:  # noqa: BLE001
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:  # noqa: BLE001
            return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them (the product has no CPU
    fallback to run them on)."""
    if any(it.get_closest_marker("gpu") for it in items) and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device: the CUDA path has no CPU fallback")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)
