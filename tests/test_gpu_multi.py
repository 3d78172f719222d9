"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU through torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_evaluation_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_two_devices_in_one_process():
    """one process holding plans on two GPUs (the ABI takes a device ordinal): the shared-memory opt-ins of the
    kernels are per device, so the dense GEMM path must work on the second device as on the first."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    import jbugs
    from c_oracle import COracle
    from helpers import random_thetas
    from test_gpu_parity import _check
    rng = np.random.default_rng(3)
    N, P, C = 3000, 33, 70
    X = np.asfortranarray(rng.standard_normal((N, P)))
    bt = rng.normal(0.0, 0.3, P)
    data = {"N": N, "P": P, "X": X, "y": (rng.random(N) < 1.0 / (1.0 + np.exp(-X @ bt))).astype(float)}
    theta = random_thetas(np.concatenate([[0.5, 0.1], bt]), C, seed=3, scale=0.1)
    for dev in (0, 1, 0):
        m = jbugs.set_evaluation_mode(jbugs.compile(jbugs.examples.DENSE_LOGISTIC, data), jbugs.UseCUDABatch(device=dev))
        _check(m.cuda, COracle(m.plan), theta)
