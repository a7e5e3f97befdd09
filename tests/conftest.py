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
