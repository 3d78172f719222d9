"""bench.py's JSON-line contract, as far as it can be checked without a GPU: the reference arm (CPU restatement) prints
one line with the required keys; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

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
