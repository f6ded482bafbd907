"""bench.py's reference arm (task contract: `bench.py --impl reference` times the CPU restatement on host cores and prints
ONE JSON line with the base keys + impl / cpu_baseline / e2e) and the shape of the committed GPU-arm line.  CPU only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config"}


def _baseline_metric():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)["metric"]


def _run_reference(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run_reference()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == _baseline_metric() and d["unit"] == "examples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert set(d["config"]) >= {"workload"} and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["index"]["value"] > 0 and d["index"]["unit"] == "queries/s"


def test_reference_arm_other_ranks_exit_quietly():
    assert _run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_committed_gpu_line_has_the_contract_keys():
    prof = os.path.join(ROOT, "profiles")
    lines = sorted(f for f in os.listdir(prof) if f.endswith("_bench_line.json"))
    if not lines:
        pytest.skip("no committed bench line yet")
    with open(os.path.join(prof, lines[-1])) as f:
        d = json.load(f)
    assert BASE_KEYS <= set(d) and d["metric"] == _baseline_metric()
    assert {"e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
