import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build (or re-use) the in-tree native libraries once per session."""
    from duckstation_b200 import build

    build.build_generator()
    build.build_oracle()
    build.build_hostsim()
    build.build_reference()
    try:
        build.build_product()
    except RuntimeError:
        # nvcc missing: the prebuilt libpsxgpu.so that travelled with the snapshot is used as-is
        if not os.path.exists(os.path.join(ROOT, "duckstation_b200", "libpsxgpu.so")):
            raise
    yield
