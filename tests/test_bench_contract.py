"""bench.py's JSON contract, exercised on the CPU through the reference arm
(`--impl reference` times the oracle port on a bounded sample; no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--halos", "1",
                        "--n-total", "120000", "--box", "40", "--cells", "1", "--steps", "1", "--warmup", "0",
                        "--cpu-halos", "1", "--order", "2", "--subsample", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "halos_per_s" and d["unit"] == "halos/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["order"] == 2 and d["config"]["subsample"] == 2 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "halos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"


def test_ours_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--halos", "1", "--n-total", "50000", "--box", "40",
                        "--cells", "1", "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0   # no CPU fallback: the product path needs the CUDA library and a B200
