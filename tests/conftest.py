import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; without a device they fail loudly instead of silently passing.
    pass


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own CUDA kernels compiled for sm_100 (oracle/_ref), when the prebuilt .so travelled."""
    import ctypes

    p = ROOT / "oracle" / "_ref" / "libnt_ref.so"
    if not p.exists():
        pytest.skip("oracle/_ref/libnt_ref.so not built")
    return ctypes.CDLL(str(p))
