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


@pytest.mark.gpu
def test_product_arm_prints_the_contract_line_with_every_leg():
    """GPU tier: a short run of the product arm — one JSON line with the base-contract keys, roofline / cpu_baseline / e2e / clocks / parity, the
    same `config` object as the reference arm, and the legs for the other BASELINE.json configs (small step counts; the numbers are not judged here)."""
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--legs", "c2f,e2e,cpu,c3,c4,c1", "--e2e-seconds", "1"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches", "parity", "configs"):
        assert k in out, k
    assert out["dtype"] == "f64" and out["gpu_launches"] == 9 and out["value"] > 1e6
    assert set(out["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and 0.3 < out["roofline"]["frac"] < 1.2
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1
    assert out["e2e"]["h2d_bytes_per_step"] > 0 and out["e2e"]["d2h_bytes_per_step"] > 0 and out["e2e"]["matches_device_path"] is True
    assert out["parity"]["relres_max"] <= 1e-13 and out["parity"]["relerr_sample"] <= 1e-12
    for leg in ("c2_fused", "c3", "c4", "c1"):
        assert leg in out["configs"] and "error" not in out["configs"][leg], out["configs"].get(leg)
    assert out["configs"]["c3"]["parity"]["relerr_sample"] <= 1e-12 and out["configs"]["c4"]["parity"]["relerr"] <= 1e-12
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    assert out["config"] == mod.workload_config(1)
