import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hostsim():
    """CPU twin of the kernels (tests/hostsim) -- test infrastructure, never used by the package."""
    from swinvox_b200 import _lib
    so = os.path.join(ROOT, "tests", "hostsim", "libsvx_hostsim.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim", "svx_hostsim.cpp"),
            os.path.join(ROOT, "swinvox_b200", "csrc", "svx_api.cu"),
            os.path.join(ROOT, "include", "swinvox_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["bash", os.path.join(ROOT, "tests", "hostsim", "build.sh")])
    return _lib.bind(so)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200 import _lib
    return _lib.get()
