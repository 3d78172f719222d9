"""bench.py's JSON-line contract, as far as it can be checked without a GPU: the reference arm (CPU restatement) prints
one line with the required keys; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300, env=e)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--groups", "3000", "--steps", "1", "--warmup", "1", env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "logdensity+gradient obs-evals/s" and d["unit"] == "obs-evals/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    # every core of the affinity mask, whatever OMP_NUM_THREADS says (torchrun exports 1)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--groups", "3000", "--steps", "1", "--warmup", "1", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--groups", "1000", "--chains", "8", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_our_arm_line_on_the_gpu():
    r = _run("--groups", "200000", "--chains", "128", "--e2e-chains", "64", "--steps", "3", "--warmup", "3", "--sustain-s", "0.6")  # 0.8 GB: larger than L2
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "roofline", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert "impl" not in d or d["impl"] != "reference"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "f64"
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0.0 < rf["frac"] < 1.5
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["algorithmic_bytes"] == 2 * 8 * 128 * d["config"]["D"]
    assert d["gpu_launches"] >= 3 * 3          # at least prologue + plate + finalize per timed step
    e = d["e2e"]
    # every chain of the batch crosses PCIe (two slabs of 64), in both host layouts
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 8 * d["config"]["D"] * 128 and e["matches_resident"] is True
    assert e["chains"] == 128 and e["slabs_per_step"] == 2 and e["param_fast"]["value"] > 0
    assert e["value"] < d["value"]             # host buffers cross PCIe inside the timed region
    # the plate kernel is timed in EVERY timed step (library events); its mean cannot exceed the step it is part of
    assert rf["kernel_ms_samples"] == 3 and rf["kernel_ms"] <= d["ms_per_step"] * 1.001 and rf["kernel_ms"] <= rf["kernel_ms_max"]
    assert d["ms_per_step_p50"] <= d["ms_per_step_p95"]
    assert d["sustained"]["seconds"] >= 0.1 and d["sustained"]["kernel_ms_mean"] <= d["sustained"]["ms_per_step_mean"] * 1.001
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
