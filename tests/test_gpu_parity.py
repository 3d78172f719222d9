"""GPU parity: libjbugs_cuda (through the C ABI) against the oracle on the same seeded inputs.
Tolerance: 1e-10 relative in fp64 on logp and gradient (BASELINE.json north_star)."""
import numpy as np
import pytest

import jbugs
from conftest import grad_err, rel_err
from helpers import ZOO, random_thetas, start_theta, synth_rats

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _check(cp, co, theta, layout="chain_fast", tol=TOL):
    lp, g, st = cp.eval_host(theta, layout=layout)
    lp0, g0, st0 = co.eval_batch(theta, layout=layout)
    assert np.array_equal(st, st0)
    assert rel_err(lp, lp0) < tol, (lp[:4], lp0[:4])
    ok = np.isfinite(lp0)  # a chain at -Inf (outside the support, not a DomainError) has no meaningful gradient
    if ok.any():
        assert grad_err(g[:, ok], g0[:, ok]) < tol
    return lp, g


@pytest.mark.parametrize("dynamic", [1, 0])
@pytest.mark.parametrize("name,text,fx,ini", ZOO, ids=[z[0] for z in ZOO])
def test_examples_match_oracle(fixtures, name, text, fx, ini, dynamic):
    import graph_oracle as go
    from c_oracle import COracle
    data = fixtures[fx]["data"]
    m = jbugs.compile(text, data, fixtures[fx][ini])
    gm = go.compile(text, data, fixtures[fx][ini]).use_generated_order()
    co = COracle(m.plan)
    cp = m.cuda
    cp.set_option("dynamic", dynamic)
    th0 = start_theta(name, m.plan.D, gm)
    for C in (1, 7, 130):
        theta = random_thetas(th0, C, seed=C)
        lp, g = _check(cp, co, theta)
        _check(cp, co, theta, layout="param_fast")
    # the per-node graph oracle (dual numbers through the node loop) on the first chain
    lp_g, g_g = gm.logdensity_and_gradient(theta[:, 0])
    assert abs(lp[0] - lp_g) <= 1e-10 * abs(lp_g)
    assert grad_err(g[:, 0], g_g) < 1e-10  # the independent per-node oracle, same 1e-10 as against the C restatement


@pytest.mark.parametrize("name,text,tol", [("rats", jbugs.examples.RATS, 1e-10), ("dogs", jbugs.examples.DOGS, 1e-10),
                                           ("blockers", jbugs.examples.BLOCKERS, 1e-6)])
@pytest.mark.parametrize("transformed", [True, False])
def test_golden_logjoint(fixtures, name, text, tol, transformed):
    """the reference's own known answers at the example inits (test/model/evaluation.jl:14-33,355-379, rtol 1e-6
    there; the Blockers constant is only good to 1.5e-8, see tests/test_oracle_goldens.py)."""
    fx = fixtures[name]
    m = jbugs.compile(text, fx["data"], fx["inits"])
    if not transformed:
        m = jbugs.settrans(m, False)
    lp = jbugs.LogDensityProblems.logdensity(m, jbugs.getparams(m))
    gold = fx["golden_logjoint"]["transformed" if transformed else "untransformed"]
    assert abs(lp - gold) <= tol * abs(gold), (lp, gold)


@pytest.mark.parametrize("N,C", [(1000, 64), (10000, 33)])
def test_rats_scaled(N, C):
    from c_oracle import COracle
    data, alpha, beta = synth_rats(N)
    m = jbugs.compile(jbugs.examples.RATS, data)
    co = COracle(m.plan)
    th0 = np.zeros(m.plan.D)
    names = None
    # theta layout of the generated order: 5 hypers then (beta_i, alpha_i) interleaved
    th0[:5] = [np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)]
    th0[5::2] = beta
    th0[6::2] = alpha
    theta = random_thetas(th0, C, seed=3, scale=0.05)
    _check(m.cuda, co, theta)

