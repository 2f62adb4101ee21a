import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from stair_step_detector_b200 import build
    build.build()
    import stair_step_detector_b200 as S
    return S.lib()


@pytest.fixture(scope="session")
def S(built_lib):
    import stair_step_detector_b200 as S
    return S


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.load_oracle()
