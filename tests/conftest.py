import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        from tsta_b200 import api
        return api.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The CUDA path.  Fails (not skips) when the library is missing on a GPU box."""
    from tsta_b200 import api
    api.load()
    if api.device_count() <= 0:
        pytest.fail("no CUDA device: gpu-marked tests must run on the B200 box")
    return api


GOLDEN = ROOT / "tests" / "golden"
