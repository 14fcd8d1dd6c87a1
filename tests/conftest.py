import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/): the checker every parity test compares against."""
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def emu():
    """Host emulation of the kernel bodies (tests/emu): same source as the CUDA kernels, one virtual thread."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    from tests.emu_binding import Emu
    return Emu()


@pytest.fixture(scope="session")
def engine():
    """The real engine on cuda:0 through the C ABI.  No fallback: fails if the library or the GPU is missing."""
    import addapt_b200
    e = addapt_b200.Engine(0)
    yield e
    e.close()
