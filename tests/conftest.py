import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def built():
    """Both shared libraries, built in-tree (nvcc cross-compiles without a GPU)."""
    from base_b200 import build
    build.build_all()
    return build


@pytest.fixture(scope="session")
def ref(built):
    from tests import _ref
    return _ref.load(built.REF_LIB)
