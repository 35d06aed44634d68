import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libira.so, building it if this checkout has not been built yet."""
    from irotavg_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def solver(built_lib):
    import irotavg_b200 as ira
    if ira.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    s = ira.Solver()
    yield s
    s.close()
