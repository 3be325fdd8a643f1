"""bench.py's `--impl reference` arm runs without a GPU: its one JSON line must carry the keys
the driver reads, for both CPU baselines (the unmodified reference package when baseline/_ref is
installed, the numpy port of oracle/ otherwise)."""

import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_line(kind):
    if kind == "reference" and not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "fast_forward")):
        pytest.skip("baseline/_ref is not installed here")
    env = dict(os.environ, FFX_CPU_Q_PER_CORE="1", FFX_CPU_DOCS="10000", FFX_CPU_BASELINE="port" if kind == "port" else "")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout  # exactly one line on stdout
    line = json.loads(lines[0])
    assert REQUIRED <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["config"]["workload"] == "c3_msmarco_doc_maxp"
    base = line["cpu_baseline"]
    assert base["kind"] == kind and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone measures and prints; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
