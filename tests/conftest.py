import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")
    config.addinivalue_line("markers", "auto_schedule: leave the choice of the schedule (windowed / sequential chains) to the library")


@pytest.fixture(scope="session")
def zzb():
    """The product package (host side); importing it never touches the GPU."""
    return graft.load_package()


@pytest.fixture(scope="session")
def gpu(zzb):
    """Initialised device; GPU tests fail (not skip) when the CUDA path is unavailable."""
    zzb.init(0)
    return zzb


@pytest.fixture(autouse=True)
def _windowed_schedule_by_default(request, monkeypatch):
    """Small or densely coupled problems run as sequential chains by default (zz_seq.cuh).  The parity tests of the windowed
    kernels use exactly such problems, so they pin the windowed schedule unless a test asks for a schedule itself
    (zzb_run_set("schedule") wins over the environment); tests of the automatic choice opt out with the marker below."""
    if "auto_schedule" in request.keywords:
        monkeypatch.delenv("ZZB200_DEFAULT_SCHEDULE", raising=False)
    else:
        monkeypatch.setenv("ZZB200_DEFAULT_SCHEDULE", "1")
