import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    """The package with the CUDA library loaded; fails loudly when the library or the device is missing."""
    L = pkg._lib.load()
    assert L.b2c_device_count() > 0, "no sm_100 device visible: the CUDA path cannot run (there is no CPU fallback)"
    return pkg
