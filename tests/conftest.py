import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="module")
def device_fdiv():
    """Installs the device's __fdividef as the oracle's division (cudapoa_nw_banded.cuh:207 under -use_fast_math). GPU tests only."""
    import ctypes as C
    import oracle_lib as ol
    from genomeworks_b200 import cudapoa
    cache = {}

    def fdiv(a, b):
        k = (a, b)
        if k not in cache:
            cache[k] = float(cudapoa.device_fdividef([a], [b])[0])
        return cache[k]

    cb = ol.FDIV_T(fdiv)
    ol.lib().oracle_set_fdiv(C.cast(cb, C.c_void_p))
    yield cb
    ol.lib().oracle_set_fdiv(None)
