import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not errored) on a box without a CUDA device; the product itself still fails
    loudly there (frcs_ctx_create -> FRCS_E_CUDA), see tests/test_abi.py."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (cudaGetDeviceCount == 0)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


_CIRC = {}


@pytest.fixture(scope="session")
def circuits(oracle):
    """Oracle circuits keyed by (logn, kind), built lazily."""
    def get(logn, kind=0):
        if (logn, kind) not in _CIRC:
            _CIRC[(logn, kind)] = oracle.Circuit(logn, kind)
        return _CIRC[(logn, kind)]
    return get


_CTX = {}


@pytest.fixture(scope="session")
def contexts():
    """Product contexts (GPU) keyed by logn."""
    from falcon_r1cs_b200 import api

    def get(logn):
        if logn not in _CTX:
            _CTX[logn] = api.Context(logn)
        return _CTX[logn]
    yield get
    for c in _CTX.values():
        c.close()
    _CTX.clear()
