import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """a GPU test that hangs (a kernel waiting for a peer, a lost NCCL rank) must end the run, not hold the box until the driver's
    limit: 15 minutes per test, far above the slowest one (pytest-timeout; without the plugin the marker is inert)"""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def gpu_backend_lib():
    """The CUDA backend; GPU tests must fail loudly (not skip) when it is missing."""
    import wumingpic_b200 as w
    return w.load_library()
