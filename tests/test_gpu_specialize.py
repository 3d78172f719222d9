"""Run-time specialisation (jbugs_plan_specialize): the plate kernel compiled for the exact shape of a plate that has no
in-tree specialisation must give the interpreter's results (and the C restatement's)."""
import os
import stat
import time

import numpy as np
import pytest

import jbugs
from c_oracle import COracle
from helpers import random_thetas
from random_models import make
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["random-3", "nested-poisson"])   # NVRTC in-process: a second or two each
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


def test_cold_compile_time_and_cache(tmp_path):
    """no toolchain on the box: NVRTC in-process, headers embedded in the library; <= 3 s cold for a Rats variant, a
    warm start loads the cubins from a per-user 0700 cache; a cache directory others can write to is refused"""
    from helpers import synth_rats
    text = jbugs.examples.RATS.replace("mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar)",
                                       "mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar) + gam * z[i]")
    text = text.replace("alpha.c ~ dnorm(0.0, 1.0E-6)", "alpha.c ~ dnorm(0.0, 1.0E-6)\n    gam ~ dnorm(0.0, 0.01)")
    data, _, _ = synth_rats(3000)
    data["z"] = np.random.default_rng(1).standard_normal(3000)
    m = jbugs.compile(text, data)
    cp = m.cuda
    cache = tmp_path / "cache"
    t0 = time.perf_counter()
    cp.specialize(str(cache))
    cold = time.perf_counter() - t0
    assert "specialised_at_run_time" in cp.describe()
    assert cp.last_compile_seconds() > 0.0
    print(f"cold specialisation: {cold:.2f} s (NVRTC {cp.last_compile_seconds():.2f} s)")
    assert cold < 5.0, f"cold specialisation took {cold:.2f} s"   # (2.8 s measured: the quick variant, three widths in parallel)
    theta0 = random_thetas(np.zeros(m.plan.D), 130, seed=2, scale=0.05)
    _check(cp, COracle(m.plan), theta0)                              # the quick variant is a correct kernel in its own right
    assert stat.S_IMODE(os.stat(cache).st_mode) == 0o700
    assert len(list(cache.glob("plate_*.cubin"))) >= 3
    # tiered: the call returned with the quick-to-compile variant (non-inlined remainder groups); the best variant is
    # compiled on a worker thread and takes over at a later evaluation
    deadline = time.perf_counter() + 120
    while "quick_variant" in cp.describe() and time.perf_counter() < deadline:
        time.sleep(0.25)
        cp.eval_host(np.zeros((m.plan.D, 8)))
    assert "quick_variant" not in cp.describe() and "specialised_at_run_time" in cp.describe()
    assert len(list(cache.glob("plate_*.cubin"))) == 6
    theta = random_thetas(np.zeros(m.plan.D), 130, seed=2, scale=0.05)
    _check(cp, COracle(m.plan), theta)
    m2 = jbugs.compile(text, data)
    t0 = time.perf_counter()
    m2.cuda.specialize(str(cache))
    warm = time.perf_counter() - t0
    assert m2.cuda.last_compile_seconds() == 0.0 and warm < 0.5
    _check(m2.cuda, COracle(m2.plan), theta)
    # without blocking: the interpreter answers until the worker thread is done, then the specialised kernels take over
    m4 = jbugs.compile(text, data)
    t0 = time.perf_counter()
    m4.cuda.specialize_async("none")
    assert time.perf_counter() - t0 < 0.5
    assert "dynamic" in m4.cuda.describe()
    _check(m4.cuda, COracle(m4.plan), theta)
    deadline = time.perf_counter() + 120
    while "dynamic" in m4.cuda.describe() and time.perf_counter() < deadline:
        time.sleep(0.25)
        m4.cuda.eval_host(theta[:, :40])
    assert "specialised_at_run_time" in m4.cuda.describe()
    _check(m4.cuda, COracle(m4.plan), theta)
    shared = tmp_path / "shared"
    shared.mkdir()
    os.chmod(shared, 0o777)
    m3 = jbugs.compile(text, data)
    m3.cuda.specialize(str(shared))           # not trusted: compiled again, nothing read from or written to it
    assert m3.cuda.last_compile_seconds() > 0.0 and not list(shared.glob("*.cubin"))


@pytest.mark.parametrize("name", ["bones", "dugongs", "lsat", "air", "leuk", "magnesium", "leukfr"])
def test_expression_tapes_compile_to_straight_line_code(fixtures, name):
    """expression plates: the tape written out as code (registers instead of the interpreter's local-memory arrays)"""
    from helpers import ZOO, start_theta
    import graph_oracle as go
    _, text, fx, ini = next(z for z in ZOO if z[0] == name)
    data, inits = fixtures[fx]["data"], fixtures[fx][ini]
    m = jbugs.compile(text, data, inits)
    cp = m.cuda
    gm = go.compile(text, data, inits).use_generated_order()
    theta = random_thetas(start_theta(name, m.plan.D, gm), 70, seed=4, scale=0.05)
    before = cp.eval_host(theta)
    assert "expression_tape" in cp.describe()
    cp.specialize()
    assert "expression_tape_specialised_at_run_time" in cp.describe()
    after = cp.eval_host(theta)
    assert np.array_equal(before[2], after[2])
    ok = before[2] == 0
    assert np.max(np.abs(after[0][ok] - before[0][ok]) / np.abs(before[0][ok])) < 1e-12
    _check(cp, COracle(m.plan), theta)
    _check(cp, COracle(m.plan), theta[:, :5])


def test_marginalised_mixture_tape_compiles():
    from test_marginalization import GMM3, POISMIX, _random_cases
    for name in ("gmm3", "poisson_mixture", "gmm_weights_from_a_parameter"):
        text, data = _random_cases()[name]
        m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
        cp = m.cuda
        theta = 0.4 * np.random.default_rng(5).standard_normal((m.plan.D, 70))
        before = cp.eval_host(theta)
        cp.specialize()
        assert "expression_tape_specialised_at_run_time" in cp.describe()
        after = cp.eval_host(theta)
        assert np.max(np.abs(after[0] - before[0]) / np.maximum(1.0, np.abs(before[0]))) < 1e-12
        _check(cp, COracle(m.plan), theta)


def test_cervix_tape_compiles():
    """Cervix under UseAutoMarginalization: Bernoulli latents with a folded second child, selects on data leaves and
    PICK_K in one tape -- interpreter == generated straight-line code == C restatement"""
    from jbugs import examples as ex
    from test_marginalization import _cervix_data
    m = jbugs.compile(ex.CERVIX, _cervix_data(), evaluation_mode=jbugs.UseAutoMarginalization())
    cp = m.cuda
    theta = 0.5 * np.random.default_rng(5).standard_normal((m.plan.D, 70))
    before = cp.eval_host(theta)
    cp.specialize()
    assert "expression_tape_specialised_at_run_time" in cp.describe()
    after = cp.eval_host(theta)
    assert np.max(np.abs(after[0] - before[0]) / np.maximum(1.0, np.abs(before[0]))) < 1e-12
    _check(cp, COracle(m.plan), theta)
