import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the native libraries once per session (nvcc cross-compiles without a GPU)."""
    from fyrox_b200 import _lib, build

    if not (os.path.exists(_lib.LIB_PATH) and os.path.exists(_lib.SCENEGEN_PATH)):
        build.build_all()
    import oracle_binding

    oracle_binding.lib()


@pytest.fixture()
def ctx():
    import fyrox_b200 as fb

    c = fb.Context()
    yield c
    c.close()
