"""bench.py's reference arm runs on the CPU: check the JSON-line contract of both workloads (keys the driver reads)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,metric", [("pretrain", "CCD pretrain images/sec (ViT-Small, 3x32x128)"),
                                             ("finetune", "CCD finetune images/sec (ViT-Small, 3x32x128)")])
def test_reference_arm_json_line(workload, metric):
    env = dict(os.environ, CCD_CPU_THREADS="4", CCD_CPU_BUDGET_S="60")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                        "--warmup", "1", "--cpu-batch", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                     # exactly ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    import ref_import
    # the UNMODIFIED reference modules whenever the reference tree (or its oracle/_ref copy) is present
    want_kind = "reference" if ref_import.reference_available() else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
