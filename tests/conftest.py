import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libicg_oracle.so).  Test infrastructure only."""
    path = os.path.join(ROOT, "oracle", "libicg_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = C.CDLL(path)
    from tests import oracle_api
    oracle_api.declare(lib)
    return lib


@pytest.fixture(scope="session")
def klt_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "klt_golden.npz"))
