import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


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
