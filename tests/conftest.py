import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import orc as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def pkg():
    import _pkg
    return _pkg.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    """GPU context of the product library; GPU tests only."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def hostsim():
    import ctypes as C
    import subprocess
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    return C.CDLL(os.path.join(d, "libhostsim.so"))


@pytest.fixture(scope="session")
def fhew_setup(orc):
    """FHEW-T key (boolean.rs:225-239) from the oracle, seeded."""
    P = orc.fhew_testing_param()
    K = orc.FhewKey(P, 0x5EED0001)
    return P, K, K.export()
