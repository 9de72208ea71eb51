import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _build_oracle():
    """compile oracle/liboracle*.so if missing (gcc only; the checker, not the product)"""
    import subprocess
    from oracle import pyoracle
    if not (pyoracle.available("oracle", False) and pyoracle.available("oracle", True)):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


@pytest.fixture(scope="session")
def oracle():
    _build_oracle()
    from oracle.pyoracle import Oracle
    return Oracle(False)


@pytest.fixture(scope="session")
def oracle_fma():
    _build_oracle()
    from oracle.pyoracle import Oracle
    return Oracle(True)


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    if not pyoracle.available("ref", False):
        pytest.skip("oracle/_ref/libinovesa_ref.so not built (needs /root/reference)")
    return pyoracle.Ref(False)


@pytest.fixture(scope="session")
def ref_fma():
    from oracle import pyoracle
    if not pyoracle.available("ref", True):
        pytest.skip("oracle/_ref/libinovesa_ref_fma.so not built (needs /root/reference)")
    return pyoracle.Ref(True)


@pytest.fixture(scope="session")
def capi():
    from inovesa_b200 import capi
    capi.load()
    return capi


@pytest.fixture(scope="session")
def ctx_strict(capi):
    c = capi.Context(0, capi.ARITH_STRICT)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_fma(capi):
    c = capi.Context(0, capi.ARITH_FMA)
    yield c
    c.close()


@pytest.fixture
def rng():
    return np.random.default_rng(20261017)


def rel_l1(a, b):
    """relative L1 distance of a to the reference b (north_star's parity measure)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).sum() / np.abs(b).sum()
