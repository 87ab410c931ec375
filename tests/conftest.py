import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that deadlocks a kernel must fail, not hang the run: 300 s per test (the slowest takes ~10 s) when
    pytest-timeout is installed (it is in this image; without it the marker is inert)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(300, method="thread"))    # a hung kernel blocks inside a C call: SIGALRM would never be served


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built():
    """Make sure libsuo_b200.so and the oracle are built (no GPU needed)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from suo_slam_b200 import _lib
    ctx = _lib.Context(device=0, max_crops=8, crop_res=64, num_kp=41)
    yield ctx
    ctx.close()
