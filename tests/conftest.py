import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test (oracle set computations)")


@pytest.fixture(autouse=True)
def _lp_backend(request):
    """Set pipeline LPs: tests marked ``gpu`` run them on the GPU (the package's default); CPU tests select the host
    backend explicitly - one HiGHS LP per call, which is what the reference does (``utils_polytope.py:19``)."""
    from rtmpc_b200 import polytope as pc
    pc.set_lp_backend("gpu" if request.node.get_closest_marker("gpu") else "highs")
    yield
    pc.set_lp_backend("gpu")
