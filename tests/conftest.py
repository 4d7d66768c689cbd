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


_BUDGET_REPORT = {}


@pytest.fixture(scope="session")
def budget_report():
    """Collects the measured error fractions of the float64-budget tests (tests/f64_truth.py); written at the end
    of the session to gpurun_out/f64_budget_report_{cpu,gpu}.json (a copy per round lives under profiles/)."""
    def record(name, **fields):
        _BUDGET_REPORT[name] = fields
    return record


def pytest_sessionfinish(session, exitstatus):
    if not _BUDGET_REPORT:
        return
    import json
    try:
        import torch
        where = "gpu" if torch.cuda.is_available() else "cpu"
    except Exception:
        where = "cpu"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "f64_budget_report_%s.json" % where), "w") as f:
        json.dump(_BUDGET_REPORT, f, indent=1, sort_keys=True)
