"""bench.py contract checks that need no GPU: the reference arm runs end to end and prints ONE JSON line
with the agreed keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(env_extra=None, *args):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_line():
    lines = _run(None, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--cpu-sample", "16")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "layer_band_solves_per_s" and d["unit"] == "layer*band solves/s" and d["higher_is_better"] is True
    assert d["value"] > 1e5 and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "crt1d", "solvers", "_solve_2s.py"))
    assert cb["kind"] == ("reference" if staged else "port")  # the unmodified reference when build() staged it
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "scenarios" in cb["sample"]
    c1 = cb["cfg1_default_case"]  # BASELINE.json configs[0]: default case, 2s, one core, best of 5
    assert c1["cores"] == 1 and c1["kind"] == cb["kind"] and abs(c1["checksum_F"] - 4.778896367937e4) < 1e-6
    assert d["scaling"] == "strong" and "configs[2]" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    lines = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
                 "--warmup", "1")
    assert lines == []
