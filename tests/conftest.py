import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import dymu_b200
    return dymu_b200.load()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build(ref=True, port=True)
    return oracle


@pytest.fixture(scope="session")
def ref_lib(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libdymu_ref.so not built (needs /root/reference at build time)")
    return oracle_mod.reference()


def rel_err(a, b):
    """max |a-b| / |b| over cells where b is finite and non-zero; masks must agree."""
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b) & (b != 0)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.abs(b[fin])))
