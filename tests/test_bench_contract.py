"""bench.py's contract on a box without a GPU: the reference arm and the blocked line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(*args):
    return subprocess.run([sys.executable, "bench.py", *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=300)


def test_reference_arm_reports_unavailable_and_exits_zero():
    r = run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "not staged" in line["unavailable"]


def test_own_arm_never_reports_a_number_for_the_blocked_metric(built):
    r = run("--steps", "2", "--warmup", "1")
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["metric"] == json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    assert line["value"] is None and line["e2e"] is None and line["vs_baseline"] is None
    assert line["roofline"] is None and line["cpu_baseline"] is None
    assert "BLOCKED" in line["blocked"]
    if "error" in line:          # no GPU here: loud, non-zero, and still no fallback number
        assert r.returncode == 1 and line["groundwork"] is None and line["gpu_launches"] == 0
    else:
        assert r.returncode == 0 and line["gpu_launches"] > 0
