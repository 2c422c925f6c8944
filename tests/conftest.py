import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refhost: needs /root/reference (build container only); skipped elsewhere")


@pytest.fixture(params=["offline", "jit"])
def jit_mode(request, monkeypatch):
    """Runs a structured-path test twice: with the offline runtime-table colour-pass kernels (MCG_JIT=0) and with the
    NVRTC-specialised build of the same source (MCG_JIT=1: mcg_pass_m0/m1, the kernels the bench number comes from).
    The variable is read when a kernel is first launched for a system (structured.cu: jit_enabled)."""
    monkeypatch.setenv("MCG_JIT", "1" if request.param == "jit" else "0")
    return request.param


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
