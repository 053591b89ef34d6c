"""bench.py contract, the part that runs without a GPU: the reference arm prints exactly ONE JSON line on
stdout with the keys the driver reads (metric, value, unit, n_gpus, steps, warmup, ms_per_step, ...)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_emits_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-800:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:400]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ecdsa_verifies_per_sec" and d["unit"] == "verifies/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["steps"] == 1 and d["config"]["workload"].startswith("ECDSA verify batch 2^20")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the product arm must exit non-zero, not print numbers."""
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
