import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def zzb():
    """The product package (host side); importing it never touches the GPU."""
    return graft.load_package()


@pytest.fixture(scope="session")
def gpu(zzb):
    """Initialised device; GPU tests fail (not skip) when the CUDA path is unavailable."""
    zzb.init(0)
    return zzb
