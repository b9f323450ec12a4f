"""bench.py's reference arm and the shape of its JSON line (the contract the driver reads).  The CUDA arm needs a GPU
and is exercised by the driver; here: the CPU reference arm runs, prints one JSON line with the agreed keys, and the
CUDA arm refuses to run without a GPU instead of falling back."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-filters", "256"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "filter-steps/s" and line["higher_is_better"] is True
    assert line["metric"] == "PoseUKF predict+update filter-steps/s" and line["dtype"] == "f64"
    assert line["value"] > 0 and line["steps"] == 2 and line["gpu_launches"] == 0
    from oracle import oracle_lib as O
    assert line["cpu_baseline"]["kind"] == ("reference" if os.path.exists(O.REF_LIB) else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("C4")


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU path" in r.stdout
