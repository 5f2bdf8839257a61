"""bench.py's output contract on the CPU: the reference arm (`--impl reference`, the oracle port timed on the host cores)
prints exactly one JSON line on stdout with the keys the driver reads, whatever the libraries write elsewhere."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in rec, key
    assert rec["unit"] == "images/s" and rec["value"] > 0 and rec["steps"] == 1
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["value"] == rec["value"] and rec["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in rec["config"]
