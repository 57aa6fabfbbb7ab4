import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the product library and the oracle exist (prebuilt files are reused)."""
    from vg_b200 import build
    build.build()
    if not (ROOT / "oracle" / "liboracle.so").exists():
        build.build_oracle()
    yield
