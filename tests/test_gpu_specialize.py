"""Run-time specialisation (jbugs_plan_specialize): the plate kernel compiled for the exact shape of a plate that has no
in-tree specialisation must give the interpreter's results (and the C restatement's)."""
import os

import numpy as np
import pytest

import jbugs
from c_oracle import COracle
from helpers import random_thetas
from random_models import make
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


def _nvcc():
    return os.environ.get("JBUGS_NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not os.access(_nvcc(), os.X_OK), reason="nvcc is not installed on this box")
@pytest.mark.parametrize("which", ["random-3", "nested-poisson"])   # ~25 s of nvcc each
def test_specialised_kernel_matches_interpreter(which, tmp_path_factory):
    cache = str(tmp_path_factory.getbasetemp().parent / "jbugs_jit_cache")   # shared by the cases of one session
    if which.startswith("random"):
        text, data, desc = make(int(which.split("-")[1]), n_max=400)
    else:
        from test_families import NESTED, NESTED_KINDS, nested_data
        text, data = NESTED % NESTED_KINDS["poisson"], nested_data("poisson", 3000, 120, seed=5)
    m = jbugs.compile(text, data)
    cp = m.cuda
    theta = random_thetas(np.zeros(m.plan.D), 150, seed=3, scale=0.2)
    before = cp.eval_host(theta)
    assert "dynamic" in cp.describe()
    cp.specialize(cache)
    assert "specialised_at_run_time" in cp.describe() and "dynamic" not in cp.describe()
    after = cp.eval_host(theta)
    assert np.array_equal(before[2], after[2])
    assert np.max(np.abs(after[0] - before[0]) / np.abs(before[0])) < 1e-12
    _check(cp, COracle(m.plan), theta)
    _check(cp, COracle(m.plan), theta[:, :7])          # 32-chain CTAs
    cp.set_option("dynamic", 1)                         # the interpreter can still be forced
    assert "dynamic" in cp.describe()
    _check(cp, COracle(m.plan), theta)
