"""CPU tier: bench.py's driver contract on the arm that runs without a GPU (`--impl reference` = the oracle port on the host
cores), and its refusal to run the product arm without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["unit"] == "systems/s" and out["higher_is_better"] is True
    assert out["metric"] == "almost-banded QR+solve systems/sec" and out["dtype"] == "f64" and out["value"] > 0
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1 and out["cpu_baseline"]["value"] == out["value"]
    assert out["e2e"] == {"value": out["value"], "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in out["config"] and "model" not in out["config"]
    # the product arm prints the very same `config` object (the driver compares them) and the arm honours --steps / --warmup as given
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    assert out["config"] == mod.workload_config(1)
    assert out["steps"] == 1 and out["warmup"] == 0


def test_reference_arm_uses_all_cores_under_a_launcher_that_exports_one_thread():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every host core (round-1 finding)."""
    import os
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads([ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")][0])
    cores = len(os.sched_getaffinity(0))
    assert out["cpu_baseline"]["cores"] == cores, (out["cpu_baseline"]["cores"], cores)


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
