"""bench.py's reference arm runs on host cores: its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line_with_the_contract_keys():
    from oracle import refharness as rh
    if not rh.have_ref_engine():
        pytest.skip("oracle/_ref not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--ref-L", "8", "--ref-sweeps", "4", "--ref-L2", "8"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:500]          # the reference's PyInit printf and any library banner stay off stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "attempts/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["reference_arm_runs"]["lattice"] == "simple cubic 8^3"      # the arm names the lattice it really ran


def test_reference_arm_is_silent_on_other_ranks():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert p.returncode == 0 and p.stdout.strip() == ""
