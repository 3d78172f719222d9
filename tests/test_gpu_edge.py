"""GPU edge cases and size-independent properties (through the C ABI), mirroring what the reference tests:
DomainError mapping (J/test/model/domain_error_handling.jl), untransformed models (settrans), alternative theta
orders, scalar logical nodes, empty plates, re-upload of observations, invalid data, ragged tile sizes."""
import numpy as np
import pytest

import jbugs
from conftest import rel_err
from helpers import random_thetas, synth_rats, synth_seeds
from conftest import grad_err
from test_gpu_parity import TOL, _check

pytestmark = pytest.mark.gpu


def _oracle(plan):
    from c_oracle import COracle
    return COracle(plan)


@pytest.mark.parametrize("text,data,good,bad", [
    ("model { alpha ~ dnorm(0.0, 1.0)\n y ~ dpois(alpha) }", {"y": 1}, 1.0, -1.0),
    ("model { theta ~ dnorm(0.5, 1.0)\n y ~ dbin(theta, 10) }", {"y": 3}, 0.5, -0.5),
    ("model { tau ~ dnorm(1.0, 1.0)\n y ~ dnorm(0.0, tau) }", {"y": 1.0}, 1.0, -1.0),
])
def test_domain_error_maps_to_minus_inf_and_nan(text, data, good, bad):
    m = jbugs.compile(text, data)
    theta = np.array([[good, bad, good]])
    lp, g, st = m.cuda.eval_host(theta, layout="chain_fast")
    assert st.tolist() == [0, 1, 0]
    assert np.isfinite(lp[[0, 2]]).all() and lp[1] == -np.inf
    assert np.isfinite(g[:, [0, 2]]).all() and np.isnan(g[:, 1]).all()
    lp0, g0, st0 = _oracle(m.plan).eval_batch(theta)
    assert st0.tolist() == [0, 1, 0] and rel_err(lp[[0, 2]], lp0[[0, 2]]) < TOL
    # the scalar entry points of the reference API
    assert jbugs.LogDensityProblems.logdensity(m, [bad]) == -np.inf
    l, gr = jbugs.LogDensityProblems.logdensity_and_gradient(m, [bad])
    assert l == -np.inf and np.isnan(gr).all()


def test_untransformed_model(fixtures):
    fx = fixtures["rats"]
    m = jbugs.settrans(jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"]), False)
    th = jbugs.getparams(m)
    lp = jbugs.LogDensityProblems.logdensity(m, th)
    assert abs(lp - fx["golden_logjoint"]["untransformed"]) < 1e-10 * 174029.4
    theta = np.abs(random_thetas(th, 9, seed=2, scale=0.05))
    _check(m.cuda, _oracle(m.plan), theta)


@pytest.mark.parametrize("name", ["rats", "epil"])
@pytest.mark.parametrize("dynamic", [0, 1])
def test_graph_order_negative_strides(fixtures, name, dynamic):
    """UseGraph's reversed-interleaved theta layout passed as an explicit order."""
    import graph_oracle as go
    text = {"rats": jbugs.examples.RATS, "epil": jbugs.examples.EPIL}[name]
    fx = fixtures[name]
    gm = go.compile(text, fx["data"], fx["inits"])
    m = jbugs.compile(text, fx["data"], fx["inits"], theta_order=gm.param_names())
    m.cuda.set_option("dynamic", dynamic)
    theta = random_thetas(gm.getparams(), 40, seed=6, scale=0.05)
    lp, g = _check(m.cuda, _oracle(m.plan), theta)
    lp_g, g_g = gm.logdensity_and_gradient(theta[:, 3])
    assert abs(lp[3] - lp_g) <= TOL * abs(lp_g)
    assert grad_err(g[:, 3], g_g) < 1e-10


def test_index_array_slots(fixtures):
    """a scrambled theta order that no affine map can express: explicit int64 index arrays."""
    fx = fixtures["seeds"]
    base = jbugs.compile(jbugs.examples.SEEDS, fx["data"])
    names = base.lowering.param_names()
    rng = np.random.default_rng(3)
    order = [names[k] for k in rng.permutation(len(names))]
    m = jbugs.compile(jbugs.examples.SEEDS, fx["data"], theta_order=order)
    assert m.plan.plates[0].locals[0].idx_slot >= 0
    theta = random_thetas(np.zeros(m.plan.D), 33, seed=1, scale=0.3)
    _check(m.cuda, _oracle(m.plan), theta)


def test_scalar_logical_nodes_gtape():
    text = """
    model {
        for (i in 1:N) { y[i] ~ dnorm(mu + b * x[i], tau) }
        tau <- 1 / (sigma * sigma)
        sigma ~ dunif(0, 10)
        mu ~ dnorm(0, 0.001)
        b ~ dnorm(0, 0.001)
    }"""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1000)
    data = {"N": 1000, "x": x, "y": 1.0 + 2.0 * x + 0.3 * rng.standard_normal(1000)}
    m = jbugs.compile(text, data)
    theta = random_thetas(np.zeros(m.plan.D), 50, seed=2, scale=0.5)
    _check(m.cuda, _oracle(m.plan), theta)


@pytest.mark.parametrize("N,C", [(17, 5), (3001, 70)])
def test_sibling_observation_statements(N, C):
    """several observation statements sharing their random effects -> one N x K plate (Blockers pattern)."""
    from test_lowering import SIBLINGS, siblings_data
    m = jbugs.compile(SIBLINGS, siblings_data(N))
    assert len(m.plan.plates) == 1 and m.plan.plates[0].T == 3
    theta = random_thetas(np.zeros(m.plan.D), C, seed=3, scale=0.4)
    _check(m.cuda, _oracle(m.plan), theta)


def test_empty_plate_and_priors_only():
    data = {"x": np.array([8.0, 15, 22, 29, 36]), "xbar": 22.0, "N": 0, "T": 5, "Y": np.zeros((0, 5))}
    m = jbugs.compile(jbugs.examples.RATS, data)
    assert m.plan.D == 5 and not m.plan.plates
    theta = random_thetas(np.zeros(5), 4, seed=0, scale=0.5)
    _check(m.cuda, _oracle(m.plan), theta)


def test_set_observed_values_reuploads(fixtures):
    """abstractppl.jl:872-888: new observations, same structure -> data re-upload, same plan handle."""
    fx = fixtures["rats"]
    m = jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    th = jbugs.getparams(m)
    lp1 = jbugs.LogDensityProblems.logdensity(m, th)
    handle = m.cuda.h.value
    m.set_observed_values(Y=fx["data"]["Y"] + 1.0)
    assert m.cuda.h.value == handle
    lp2 = jbugs.LogDensityProblems.logdensity(m, th)
    lp2_ref = _oracle(m.plan).eval_batch(th.reshape(-1, 1))[0][0]
    assert lp1 != lp2 and abs(lp2 - lp2_ref) < TOL * abs(lp2_ref)


def test_invalid_counts_fall_back_to_interpreter():
    """k > n: Distributions' logpdf is -Inf (not a DomainError); the fused binomial path refuses such data."""
    data, b = synth_seeds(200, seed=5)
    data["r"][7] = data["n"][7] + 3.0
    m = jbugs.compile(jbugs.examples.SEEDS, data)
    theta = random_thetas(np.concatenate([[2.0, 0, 0, 0, 0], b]), 5, seed=1, scale=0.05)
    lp, g, st = m.cuda.eval_host(theta)
    assert "dynamic" in m.cuda.describe()
    lp0, _, st0 = _oracle(m.plan).eval_batch(theta)
    assert np.all(lp == -np.inf) and np.all(lp0 == -np.inf) and np.array_equal(st, st0)


@pytest.mark.parametrize("N,C", [(1003, 5), (33, 257), (4097, 64), (1, 1)])
def test_ragged_sizes(N, C):
    data, alpha, beta = synth_rats(N, seed=N)
    m = jbugs.compile(jbugs.examples.RATS, data)
    th0 = np.zeros(m.plan.D)
    th0[:5] = [np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)]
    th0[5::2], th0[6::2] = beta, alpha
    theta = random_thetas(th0, C, seed=C, scale=0.05)
    co = _oracle(m.plan)
    _check(m.cuda, co, theta)
    _check(m.cuda, co, theta, layout="param_fast")
    m.cuda.set_option("tile", 7)  # tiles that are not whole staged chunks
    _check(m.cuda, co, theta)


@pytest.mark.parametrize("N", [4096, 4098, 1001])
def test_tma_and_cp_async_staging_agree(N):
    """the data streams are staged either by the TMA bulk-copy path (aligned chunks) or by cp.async (unaligned
    chunks / odd N): both must give bit-identical results."""
    data, alpha, beta = synth_rats(N, seed=2)
    m = jbugs.compile(jbugs.examples.RATS, data)
    theta = random_thetas(np.concatenate([[1.0, 6, -5, 240, -3.5], np.column_stack([beta, alpha]).ravel()]), 70, seed=4, scale=0.05)
    m.cuda.set_option("tma", 1)
    a = m.cuda.eval_host(theta)
    _check(m.cuda, _oracle(m.plan), theta)
    m.cuda.set_option("tma", 0)
    b = m.cuda.eval_host(theta)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_bitwise_repeatability():
    data, alpha, beta = synth_rats(5000, seed=9)
    m = jbugs.compile(jbugs.examples.RATS, data)
    theta = random_thetas(np.concatenate([[1.0, 6, -5, 240, -3.5], np.column_stack([beta, alpha]).ravel()]), 200, seed=4, scale=0.05)
    a = m.cuda.eval_host(theta)
    b = m.cuda.eval_host(theta)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_full_size_plate_properties():
    """BASELINE.json size (10^6 rats x 5) with a few chains -- too big for the node oracle, so checked through
    size-independent properties: (1) the plate built from k copies of a block scores k times the block's
    likelihood + the hyper priors once; (2) the C port on two of the chains."""
    blk, k, C = 1000, 1000, 8
    d1, a1, b1 = synth_rats(blk, seed=3)
    dK = dict(d1, N=blk * k, Y=np.tile(d1["Y"], (k, 1)))
    m1 = jbugs.compile(jbugs.examples.RATS, d1)
    mK = jbugs.compile(jbugs.examples.RATS, dK)
    m0 = jbugs.compile(jbugs.examples.RATS, dict(d1, N=0, Y=np.zeros((0, 5))))
    hyp = np.array([np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)])
    loc = np.column_stack([b1, a1]).ravel()
    th1 = random_thetas(np.concatenate([hyp, loc]), C, seed=7, scale=0.02)
    thK = np.concatenate([th1[:5], np.tile(th1[5:], (k, 1))])
    lp1, g1, _ = m1.cuda.eval_host(th1)
    lpK, gK, _ = mK.cuda.eval_host(thK)
    lp0, g0, _ = m0.cuda.eval_host(th1[:5])
    assert rel_err(lpK, lp0 + k * (lp1 - lp0)) < 1e-11
    assert rel_err(gK[5:5 + 2 * blk], g1[5:]) < 1e-12
    assert rel_err(gK[-2 * blk:], g1[5:]) < 1e-12
    assert grad_err(gK[:5], g0 + k * (g1[:5] - g0)) < 1e-9
    lpo, go_, _ = _oracle(mK.plan).eval_batch(thK[:, :2])
    assert rel_err(lpK[:2], lpo) < TOL and grad_err(gK[:, :2], go_) < TOL


@pytest.mark.parametrize("N,P,C,binomial", [(300, 7, 5, False), (3000, 70, 150, False), (1000, 33, 257, True)])
def test_dense_logistic_gemm_path(N, P, C, binomial):
    """BASELINE.json config 4 shape: inprod(X[i,1:P], beta[1:P]) as FP64 tensor-core GEMMs fused with dbern / dbin."""
    rng = np.random.default_rng(N + P)
    X = np.asfortranarray(rng.standard_normal((N, P)))
    bt = rng.normal(0.0, 0.3, P)
    pr = 1.0 / (1.0 + np.exp(-X @ bt))
    data = {"N": N, "P": P, "X": X}
    text = jbugs.examples.DENSE_LOGISTIC
    if binomial:
        n = rng.integers(1, 30, N).astype(float)
        data.update(n=n, y=rng.binomial(n.astype(int), pr).astype(float))
        text = text.replace("y[i] ~ dbern(p[i])", "y[i] ~ dbin(p[i], n[i])")
    else:
        data["y"] = (rng.random(N) < pr).astype(float)
    m = jbugs.compile(text, data)
    assert len(m.plan.dense) == 1 and m.plan.dense[0].P == P
    th0 = np.concatenate([[0.5, 0.1], bt])
    theta = random_thetas(th0, C, seed=3, scale=0.1)
    co = _oracle(m.plan)
    _check(m.cuda, co, theta)
    _check(m.cuda, co, theta, layout="param_fast")


def test_abi_error_codes(fixtures):
    """nothing throws across the ABI: bad arguments come back as negative codes with a message, and the plan
    stays usable afterwards."""
    from jbugs.cuda import JBugsCudaError
    fx = fixtures["rats"]
    m = jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    cp = m.cuda
    D, C = m.plan.D, 8
    th = np.ascontiguousarray(random_thetas(jbugs.getparams(m), C, seed=1, scale=0.01))
    lp, g, st = np.empty(C), np.empty((D, C)), np.empty(C, dtype=np.int32)
    with pytest.raises(JBugsCudaError, match="unknown layout"):
        cp.eval_raw(th, C, C, 7, lp, g, st)
    with pytest.raises(JBugsCudaError, match="leading dimension"):
        cp.eval_raw(th, C - 1, C, 0, lp, g, st)
    with pytest.raises(JBugsCudaError, match="grad is null"):
        cp.eval_raw(th, C, C, 0, lp, None, st, want_grad=True)
    with pytest.raises(JBugsCudaError, match="outside"):
        cp.set_shard(0, 0, 31)
    with pytest.raises(JBugsCudaError, match="out of range"):
        cp.set_shard(3, 0, 1)
    with pytest.raises(JBugsCudaError, match="no dense predictor"):
        cp.set_dense_shard(0, 1)
    with pytest.raises(JBugsCudaError, match="unknown option"):
        cp.set_option("no_such_knob", 1)
    cp.eval_raw(th, C, 0, 0, lp, g, st)          # C = 0 is a no-op, not an error
    lp2, g2, st2 = cp.eval_host(th)              # and the handle still works
    assert np.all(st2 == 0) and np.all(np.isfinite(lp2))


@pytest.mark.parametrize("name", ["rats", "seeds", "epil", "bones", "dugongs", "stacks"])
def test_logprior_loglikelihood_and_temperature(fixtures, name):
    """the (logprior, loglikelihood, tempered_logjoint) triple of evaluate_with_values!! (evaluation.jl:116, 269-274;
    reference test test/model/evaluation.jl:112-165) and the gradient of the tempered log joint, against the per-node
    oracle's own split and dual-number gradient"""
    import graph_oracle as go
    from helpers import ZOO
    from numeric import Dual
    _, text, fx, ini = next(z for z in ZOO if z[0] == name)
    data, inits = fixtures[fx]["data"], fixtures[fx][ini]
    m = jbugs.compile(text, data, inits)
    gm = go.compile(text, data, inits).use_generated_order()
    th0 = gm.getparams()
    th0 = np.where(np.isfinite(th0), th0, 0.0)
    theta = random_thetas(th0, 5, seed=11, scale=0.05)
    T = 2.5
    lpr, ll, lp, grad, st = m.cuda.eval_tempered(theta, T)
    assert (st == 0).all()
    n = theta.shape[0]
    for c in range(5):
        _, r = gm.evaluate_with_values([Dual.seed(float(theta[i, c]), i, n) for i in range(n)], temperature=T)
        val = lambda x: x.v if isinstance(x, Dual) else float(x)  # noqa: E731
        assert abs(lpr[c] - val(r["logprior"])) <= 1e-10 * max(1.0, abs(val(r["logprior"])))
        assert abs(ll[c] - val(r["loglikelihood"])) <= 1e-10 * max(1.0, abs(val(r["loglikelihood"])), abs(val(r["logprior"])))
        tj = r["tempered_logjoint"]
        assert abs(lp[c] - tj.v) <= 1e-10 * max(1.0, abs(tj.v))
        assert grad_err(grad[:, c], np.asarray(tj.d, dtype=float)) < 1e-10
    # temperature 1 is the ordinary log density; a single theta goes through the host mirror of the reference's name
    lp1, g1, _ = m.cuda.eval_host(theta, layout="param_fast")
    _, _, lpt, gt, _ = m.cuda.eval_tempered(theta, 1.0)
    assert rel_err(lpt, lp1) < 1e-13 and grad_err(gt, g1) < 1e-12
    r1 = jbugs.evaluate_with_values(m, theta[:, 0], temperature=T)
    assert r1["tempered_logjoint"] == pytest.approx(r1["logprior"] + T * r1["loglikelihood"], rel=1e-14)


def test_asynchronous_host_calls_run_as_one_pipeline(fixtures):
    """JBUGS_EVAL_ASYNC with HOST buffers: consecutive calls only enqueue (copy-in / kernels / copy-out of one call overlap
    the next); after jbugs_plan_synchronize every call's results are in its own host buffers and equal the blocking call's"""
    import ctypes
    from helpers import synth_rats
    data, alpha, beta = synth_rats(3000)
    m = jbugs.compile(jbugs.examples.RATS, data)
    cp = m.cuda
    D = m.plan.D
    th0 = np.zeros(D)
    th0[:5] = [np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)]
    th0[5::2], th0[6::2] = beta, alpha
    C = 96
    thetas = [np.ascontiguousarray(random_thetas(th0, C, seed=s, scale=0.05)) for s in range(3)]
    want = [cp.eval_host(t, layout="chain_fast") for t in thetas]
    lps = [np.full(C, np.nan) for _ in thetas]
    grads = [np.full((D, C), np.nan) for _ in thetas]
    sts = [np.full(C, -1, dtype=np.int32) for _ in thetas]
    for t, lp, g, st in zip(thetas, lps, grads, sts):
        cp.eval_raw(t.ctypes.data, C, C, 0, lp.ctypes.data, g.ctypes.data, st.ctypes.data, True, 1, None)
    cp.synchronize()
    for (lp0, g0, st0), lp, g, st in zip(want, lps, grads, sts):
        assert np.array_equal(lp, lp0) and np.array_equal(g, g0) and np.array_equal(st, st0)
