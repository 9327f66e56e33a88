import os
import sys

import pytest

# Some tests run several ranks of the partitioned path on ONE device (separate streams, ordering by spinning flag kernels):
# every stream needs a hardware queue of its own, or a wait kernel queued in front of another rank's dispatch never ends.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (cheap no-op when they are up to date)."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "blight_b200", "lib", "libblight_b200.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "oracle", "libblight_oracle.so")):
        g.build()
    yield
