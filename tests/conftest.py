import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import pyoracle
    return pyoracle.load("port")


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference compiled in oracle/_ref (prebuilt .so travels to the GPU box)."""
    from oracle import pyoracle
    if not pyoracle.available("reference") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libcsm_ref.so not built and /root/reference absent")
    return pyoracle.load("reference")


@pytest.fixture(scope="session")
def checker(request):
    """Strongest checker available: the compiled reference, else the port."""
    from oracle import pyoracle
    if pyoracle.available("reference") or os.path.isdir("/root/reference"):
        return pyoracle.load("reference")
    return pyoracle.load("port")


@pytest.fixture(scope="session")
def handle():
    from my_lidar_graph_slam_v2_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()
