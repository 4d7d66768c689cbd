import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, "model-based-policy-optimizers_b200")
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def c_oracle():
    """The oracle's plain-C twin (oracle/c), built on demand with the committed Makefile."""
    import ctypes
    import subprocess
    cdir = os.path.join(ROOT, "oracle", "c")
    lib = os.path.join(cdir, "libmbpo_oracle.so")
    src = os.path.join(cdir, "mbpo_oracle.c")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", cdir], check=True)
    return ctypes.CDLL(lib)
