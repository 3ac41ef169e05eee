"""CPU check of bench.py's JSON contract through its reference arm (the GPU arm needs a B200): one
line, the keys the driver reads, and a cpu_baseline / e2e block of the documented shape."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--model", "tiny",
                        "--cpu-config", "tiny", "--seq", "700", "--steps", "4", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # the reference arm runs the UNMODIFIED reference classes (oracle/_ref or /root/reference), whole steps only
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["steps"] == len(cb["step_s"]) == 4 and d["steps_requested"] == 4 and "extrapolat" not in cb["sample"].replace("no extrapolation", "")
    assert abs(sum(cb["step_s"]) / len(cb["step_s"]) * 1000 - d["ms_per_step"]) < 1.0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000


def test_gpu_arm_fails_loudly_without_cuda():
    """No CPU fallback: on a box without a GPU the product arm must error out, not print a number."""
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--model", "tiny", "--seq", "700", "--steps", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
