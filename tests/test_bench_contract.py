"""CPU tests of bench.py's contract: the reference arm (`--impl reference`) of the workloads that finish in seconds
prints ONE JSON line with the keys the driver reads, runs no GPU code, and under a multi-rank launch only rank 0
prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e")


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


@pytest.mark.parametrize("workload", ["config3_replay_insert", "config3_collect_experience"])
def test_reference_arm_prints_one_contract_line(workload):
    lines = _run(["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["config"]["workload"] == workload and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["higher_is_better"] is True


def test_reference_arm_is_silent_on_other_ranks():
    lines = _run(["--impl", "reference", "--workload", "config3_replay_insert", "--steps", "1", "--warmup", "1",
                  "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []
