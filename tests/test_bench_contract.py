"""The reference arm of bench.py (the driver computes the headline ratio from it): JSON contract on CPU, and under a
2-rank launch only rank 0 prints.  (The GPU arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line():
    out = run({})
    line = json.loads(out.splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "spectra/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("scalar C_l spectra/sec") and line["value"] > 0
    assert line["steps"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_are_silent():
    out = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out == ""
