import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def built():
    """The native libraries; built on demand so a fresh checkout can run the suite."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def cuda_device(built):
    from tweeker_raytracer_b200 import core
    n = core.device_count()
    if n < 1:
        pytest.fail("a test marked gpu ran without a CUDA device: " + core.lib().rtc_last_error().decode())
    return n
