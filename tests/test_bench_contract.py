"""CPU tier: bench.py's output contract on the arm that runs without a GPU (--impl reference times the C restatement of the
reference path on the host cores): stdout is exactly ONE JSON line carrying the keys the driver reads; the GPU arm
refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "instances/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    r = _run("--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--other-batch", "256", "--mpc-streams", "64", "--mpc-resolves", "5",
             "--cpu-sample-other", "16", "--quad-batch", "8")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] >= 3
    assert d["converged_fraction"] == 1.0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["roofline"]["bound"] == "fp64-issue" and d["roofline"]["unit"] == "TFLOP/s" and d["roofline"]["peak"] > 10.0
    assert 0.0 < d["roofline"]["hbm"]["frac"] < 0.05          # real DRAM traffic: a fraction of a per cent of the HBM peak
    oc = d["other_configs"]
    assert isinstance(oc, list) and [o["workload"][0] for o in oc] == ["E", "D", "C", "Q"], oc
    for o in oc:
        assert o["value"] > 0 and o["e2e"]["value"] > 0 and 0.5 < o["converged_fraction"] <= 1.0
        assert o["cpu_baseline"]["value"] > 0 and o["cpu_baseline"]["kind"] == "port"
