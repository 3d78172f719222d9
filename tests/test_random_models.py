"""Randomly generated models of the plate class: per-node oracle == lowered plan through the C restatement (CPU), and
CUDA kernels (interpreter and, when one matches, static spec) == C restatement (GPU).  Combinations of families, links,
loop nests, random-effect levels and prior argument kinds that no fixture covers."""
import numpy as np
import pytest

import graph_oracle as go
import jbugs
from c_oracle import COracle
from conftest import grad_err, rel_err
from helpers import random_thetas
from random_models import make, make_nonlinear, make_structured

SEEDS = list(range(96))


@pytest.mark.parametrize("seed", SEEDS)
def test_random_model_plan_matches_node_loop(seed):
    text, data, desc = make(seed)
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert plan.D == gm.D and lw.param_names() == gm.param_names(), desc
    theta = random_thetas(np.zeros(plan.D), 3, seed=seed, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        if not np.isfinite(lp_g):
            assert lp[c] == lp_g, desc
            continue
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-11 * max(1.0, abs(lp_g)), (desc, text)
        assert grad_err(g[:, c], g_g) < 1e-10, (desc, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_random_model_cuda_matches_c_oracle(seed):
    from test_gpu_parity import _check
    text, data, desc = make(seed)
    m = jbugs.compile(text, data)
    co = COracle(m.plan)
    theta = random_thetas(np.zeros(m.plan.D), 37, seed=seed, scale=0.3)
    _check(m.cuda, co, theta)
    _check(m.cuda, co, theta, layout="param_fast")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(100, 124)))
def test_random_model_cuda_many_groups(seed):
    """the same generator with up to 700 groups: several tiles, staged chunks with ragged ends, 300 chains (three CTA
    columns, the last one partly idle)."""
    from test_gpu_parity import _check
    text, data, desc = make(seed, n_max=700)
    m = jbugs.compile(text, data)
    theta = random_thetas(np.zeros(m.plan.D), 300, seed=seed, scale=0.2)
    _check(m.cuda, COracle(m.plan), theta)


@pytest.mark.parametrize("seed", list(range(40)))
def test_structured_random_model_plan_matches_node_loop(seed):
    """data-indexed random effects (regrouped, padded, weighted plate) and sibling observation statements (merged plate)."""
    text, data, desc = make_structured(seed)
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert len(plan.plates) == 1 and plan.D == gm.D and lw.param_names() == gm.param_names(), desc
    theta = random_thetas(np.zeros(plan.D), 3, seed=seed, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        if not np.isfinite(lp_g):
            assert lp[c] == lp_g, desc
            continue
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-11 * max(1.0, abs(lp_g)), (desc, text)
        assert grad_err(g[:, c], g_g) < 1e-10, (desc, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(40)))
def test_structured_random_model_cuda_matches_c_oracle(seed):
    from test_gpu_parity import _check
    text, data, desc = make_structured(seed)
    m = jbugs.compile(text, data)
    theta = random_thetas(np.zeros(m.plan.D), 37, seed=seed, scale=0.3)
    _check(m.cuda, COracle(m.plan), theta)


@pytest.mark.parametrize("seed", list(range(48)))
def test_nonlinear_random_model_plan_matches_node_loop(seed):
    """expression plates: random predictors that are not linear in the parameters, every observation family"""
    text, data, desc = make_nonlinear(seed)
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert len(plan.plates) == 1 and plan.plates[0].xops, desc
    assert plan.D == gm.D and lw.param_names() == gm.param_names(), desc
    theta = random_thetas(np.zeros(plan.D), 3, seed=seed, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        if not np.isfinite(lp_g):
            assert lp[c] == lp_g, desc
            continue
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-11 * max(1.0, abs(lp_g)), (desc, text)
        assert grad_err(g[:, c], g_g) < 1e-10, (desc, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(48)))
def test_nonlinear_random_model_cuda_matches_c_oracle(seed):
    from test_gpu_parity import _check
    text, data, desc = make_nonlinear(seed, n_max=90)
    m = jbugs.compile(text, data)
    assert "expression_tape" in m.cuda.describe()
    theta = random_thetas(np.zeros(m.plan.D), 70, seed=seed, scale=0.3)
    _check(m.cuda, COracle(m.plan), theta)
    _check(m.cuda, COracle(m.plan), theta, layout="param_fast")


def _truncate_a_prior(text, seed):
    """the random model with the prior of its intercept replaced by a truncated normal (one, two or upper-only bounds)"""
    import re
    bounds = ["T(-1.5, 2.0)", "T(0, )", "T(, 1.2)"][seed % 3]
    new, n = re.subn(r"(\n  b0 ~ )[a-z]+\([^)]*\)", rf"\1dnorm(0.3, 1.5) {bounds}", text, count=1)
    return new if n else None


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_model_with_a_truncated_prior(seed):
    text, data, desc = make(seed)
    text = _truncate_a_prior(text, seed)
    if text is None:
        pytest.skip("no intercept in this model")
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert lw.param_names() == gm.param_names(), desc
    theta = random_thetas(np.zeros(plan.D), 3, seed=seed, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        if not np.isfinite(lp_g):
            assert lp[c] == lp_g, desc
            continue
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-11 * max(1.0, abs(lp_g)), (desc, text)
        assert grad_err(g[:, c], g_g) < 1e-10, (desc, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(12)))
def test_random_model_with_a_truncated_prior_cuda(seed):
    from test_gpu_parity import _check
    text, data, desc = make(seed)
    text = _truncate_a_prior(text, seed)
    if text is None:
        pytest.skip("no intercept in this model")
    m = jbugs.compile(text, data)
    _check(m.cuda, COracle(m.plan), random_thetas(np.zeros(m.plan.D), 37, seed=seed, scale=0.3))
