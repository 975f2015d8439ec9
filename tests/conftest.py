import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return REPO / "tests" / "golden"


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if needed) and loads the C-ABI library; used by both CPU and GPU suites."""
    from protstruc_b200 import _cabi, build

    build.build()
    return _cabi.load()
