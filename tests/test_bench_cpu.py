"""The reference arm of bench.py (the CPU restatement timed on the host cores) and its JSON contract; no GPU needed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # ONE JSON line on stdout, everything else goes to stderr
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "images/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("images/sec") and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "batch 16" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]
    assert j["gpu_launches"] == 0 and j["vs_baseline"] is None
