import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture
def tmp_ibu(tmp_path):
    """Unique temp file path (the reference's tests use fixed names in the CWD, mmap.rs:342-348)."""
    return str(tmp_path / "t.ibu")
