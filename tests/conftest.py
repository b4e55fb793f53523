import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle library and libdpq.so exist (both compile without a GPU)."""
    from oracle import pyoracle as po
    import deltapq_b200 as dpq
    if not os.path.exists(os.path.join(ROOT, "oracle", "libdpq_oracle.so")):
        po.build(ref=False)
    if not os.path.exists(dpq.LIB_PATH):
        dpq.build()


def load_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden4000():
    return load_golden("sift_n4000_m8")


@pytest.fixture(scope="session")
def golden1501():
    return load_golden("sift_n1501_m8")


@pytest.fixture(scope="session")
def golden_m16():
    return load_golden("sift_n600_m16")
