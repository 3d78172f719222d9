"""The remaining univariate families of SURVEY 8f rank 3 (dlogis, ddexp, dweib, dnegbin, dchisqr;
J/src/BUGSPrimitives/distributions.jl:30-33,92-95,215-217,233-235,516-518) as priors and as observation families:
per-node oracle (dual numbers) == lowered plan through the C restatement == CUDA kernels."""
import numpy as np
import pytest

import graph_oracle as go
import jbugs
from c_oracle import COracle
from conftest import grad_err, rel_err
from helpers import random_thetas

rng = np.random.default_rng(12)
N = 37
x = rng.standard_normal(N)
MODELS = {
    "priors": ("""model {
        for (i in 1:N) { y[i] ~ dnorm(mu + b[i], 1.0)
                         b[i] ~ dlogis(m, t) }
        m ~ ddexp(0.3, 2.0)
        t ~ dchisqr(3)
        w ~ dweib(2.0, 1.5)
        mu ~ dnorm(w, 0.01) }""", {"N": N, "y": rng.standard_normal(N)}),
    "obs_logistic": ("""model {
        for (i in 1:N) { y[i] ~ dlogis(mu + b[i], tau)
                         b[i] ~ dnorm(0, 1) }
        tau ~ dweib(2.0, 1.5)
        mu ~ dnorm(0, 0.01) }""", {"N": N, "y": rng.standard_normal(N)}),
    "obs_laplace": ("""model {
        for (i in 1:N) { y[i] ~ ddexp(a + b * x[i], tau) }
        tau ~ dgamma(2, 2)
        a ~ dnorm(0, 0.01)
        b ~ dlogis(0, 1) }""", {"N": N, "x": x, "y": 0.5 + x + rng.laplace(0, 0.7, N)}),
    "obs_negbin": ("""model {
        for (i in 1:N) { k[i] ~ dnegbin(p[i], r)
                         logit(p[i]) <- a + b * x[i] }
        r ~ dgamma(2, 1)
        a ~ dnorm(0, 0.1)
        b ~ dnorm(0, 0.1) }""", {"N": N, "x": x, "k": rng.negative_binomial(3, 0.4, N).astype(float)}),
    "obs_geom": ("""model {
        for (i in 1:N) { k[i] ~ dgeom(p[i])
                         logit(p[i]) <- a + b * x[i] + u[i]
                         u[i] ~ dnorm(0, tau) }
        tau ~ dgamma(2, 2)
        a ~ dnorm(0, 0.1)
        b ~ dnorm(0, 0.1) }""", {"N": N, "x": x, "k": (rng.geometric(0.4, N) - 1).astype(float)}),
    "obs_betabin": ("""model {
        for (i in 1:N) { r[i] ~ dbetabin(al[i], be, 12)
                         log(al[i]) <- a + b * x[i] }
        be ~ dgamma(2, 1)
        a ~ dnorm(0, 0.1)
        b ~ dnorm(0, 0.1) }""", {"N": N, "x": x, "r": rng.binomial(12, rng.beta(2.0, 3.0, N)).astype(float)}),
    "obs_weibull": ("""model {
        for (i in 1:N) { w[i] ~ dweib(v, lam[i])
                         log(lam[i]) <- a + b * x[i] }
        v ~ dgamma(2, 2)
        a ~ dnorm(0, 0.1)
        b ~ dnorm(0, 0.1) }""", {"N": N, "x": x, "w": rng.weibull(1.5, N) + 0.05}),
    "obs_t": ("""model {
        for (i in 1:N) { y[i] ~ dt(mu[i], tau, 4)
                         mu[i] <- a + b * x[i] + u[i]
                         u[i] ~ dt(0, 2.0, 7) }
        tau ~ dgamma(2, 2)
        a ~ dt(0.5, 0.25, 3)
        b ~ dnorm(0, 0.1) }""", {"N": N, "x": x, "y": 0.5 + x + rng.standard_t(4, N)}),
    "obs_chisq": ("""model {
        for (i in 1:N) { z[i] ~ dchisqr(k) }
        k ~ dunif(1, 10) }""", {"N": N, "z": rng.chisquare(4, N)}),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_plan_matches_node_loop(name):
    text, data = MODELS[name]
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 4, seed=3, scale=0.4)
    lp, g, st = COracle(plan).eval_batch(theta)
    assert np.all(st == 0)
    for c in range(4):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g), name
        assert grad_err(g[:, c], g_g) < 1e-10, name


def test_new_families_against_mpmath():
    """50-digit evaluation of the same node loop: the float64 formulas carry no hidden cancellation."""
    for name in ("priors", "obs_negbin", "obs_geom", "obs_betabin", "obs_weibull"):
        text, data = MODELS[name]
        gm = go.compile(text, data)
        th = 0.3 * np.ones(gm.D)
        assert gm.logdensity(th) == pytest.approx(float(gm.logdensity_mp(th)), rel=1e-13)


@pytest.mark.parametrize("text,data,bad", [
    ("model { t ~ dnorm(1, 1)\n y ~ dlogis(0, t) }", {"y": 0.3}, -1.0),
    ("model { t ~ dnorm(1, 1)\n y ~ ddexp(0, t) }", {"y": 0.3}, -1.0),
    ("model { p ~ dnorm(0.5, 1)\n k ~ dnegbin(p, 3) }", {"k": 2}, 1.5),
    ("model { p ~ dnorm(0.5, 1)\n k ~ dgeom(p) }", {"k": 2}, 1.5),
    ("model { p ~ dnorm(0.5, 1)\n k ~ dgeom(p) }", {"k": 2}, 0.0),
    ("model { a ~ dnorm(0.5, 1)\n k ~ dbetabin(a, 2, 5) }", {"k": 2}, -0.5),
    ("model { v ~ dnorm(1, 1)\n w ~ dweib(v, 1) }", {"w": 0.7}, -0.5),
    ("model { k ~ dnorm(1, 1)\n z ~ dchisqr(k) }", {"z": 0.7}, -0.5),
])
def test_domain_errors_of_new_families(text, data, bad):
    gm = go.compile(text, data)
    assert gm.logdensity([bad]) == -np.inf
    plan, _ = jbugs.lower(text, data)
    lp, g, st = COracle(plan).eval_batch(np.array([[bad]]))
    assert st[0] == 1 and lp[0] == -np.inf and np.all(np.isnan(g))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MODELS))
def test_cuda_matches_c_oracle(name):
    from test_gpu_parity import _check
    text, data = MODELS[name]
    m = jbugs.compile(text, data)
    theta = random_thetas(np.zeros(m.plan.D), 70, seed=5, scale=0.4)
    _check(m.cuda, COracle(m.plan), theta)
    theta[0, 3] = 6.0   # one chain far out in the tail: finite on both sides, or flagged identically
    _check(m.cuda, COracle(m.plan), theta)


# ---- single-node models against an independent implementation (scipy.stats), in both spaces: the shape of the
# reference's own per-distribution tests (J/test/model/evaluation.jl:167-352: logpdf + logabsdetjac of the bijector)
def _scipy_cases():
    from scipy import stats as S
    return [
        # (BUGS call, scipy distribution, support (lb, ub) or None for the real line / discrete, test value)
        ("dnorm(1.5, 4.0)", S.norm(1.5, 0.5), None, 0.7),
        ("dgamma(0.001, 0.001)", S.gamma(0.001, scale=1000.0), (0.0, np.inf), np.exp(10.0)),
        ("dgamma(3.0, 2.0)", S.gamma(3.0, scale=0.5), (0.0, np.inf), 1.3),
        ("dexp(2.5)", S.expon(scale=0.4), (0.0, np.inf), 0.9),
        ("dunif(-1.0, 3.0)", S.uniform(-1.0, 4.0), (-1.0, 3.0), 0.25),
        ("dbeta(2.0, 5.0)", S.beta(2.0, 5.0), (0.0, 1.0), 0.3),
        ("dlnorm(0.5, 4.0)", S.lognorm(s=0.5, scale=np.exp(0.5)), (0.0, np.inf), 2.2),
        ("dlogis(0.5, 4.0)", S.logistic(0.5, 0.5), None, -0.4),
        ("ddexp(0.5, 4.0)", S.laplace(0.5, 0.5), None, 1.9),
        ("dweib(1.7, 0.8)", S.weibull_min(1.7, scale=1.25), (0.0, np.inf), 0.6),
        ("dchisqr(5)", S.chi2(5), (0.0, np.inf), 3.3),
        ("dbern(0.3)", S.bernoulli(0.3), "discrete", 1.0),
        ("dbin(0.1, 10)", S.binom(10, 0.1), "discrete", 2.0),
        ("dpois(3.5)", S.poisson(3.5), "discrete", 4.0),
        ("dnegbin(0.4, 3)", S.nbinom(3, 0.4), "discrete", 5.0),
        ("dt(1.5, 4.0, 5)", S.t(5, 1.5, 0.5), None, 0.7),
        ("dgeom(0.3)", S.geom(0.3, loc=-1), "discrete", 4.0),
        ("dbetabin(2.5, 1.5, 9)", S.betabinom(9, 2.5, 1.5), "discrete", 4.0),
    ]


def _reference_logp(dist, support, x):
    """(untransformed logp, transformed logp, theta in the transformed space)"""
    if support == "discrete":
        return dist.logpmf(x), dist.logpmf(x), x
    lp = dist.logpdf(x)
    if support is None:
        return lp, lp, x
    lb, ub = support
    if np.isfinite(ub):
        u = (x - lb) / (ub - lb)
        return lp, lp + np.log((x - lb) * (ub - x) / (ub - lb)), np.log(u) - np.log1p(-u)
    return lp, lp + np.log(x - lb), np.log(x - lb)


@pytest.mark.parametrize("k", range(18))
def test_single_node_models_against_scipy(k):
    call, dist, support, x = _scipy_cases()[k]
    text = "model { a ~ %s }" % call
    lp_u, lp_t, y = _reference_logp(dist, support, x)
    gm = go.compile(text, {}, {"a": x})
    assert gm.getparams()[0] == pytest.approx(y, rel=1e-13, abs=1e-13)
    assert gm.logdensity(gm.getparams()) == pytest.approx(lp_t, rel=1e-11)
    gu = go.compile(text, {}, {"a": x}, transformed=False)
    assert gu.logdensity(gu.getparams()) == pytest.approx(lp_u, rel=1e-11)
    for transformed, th, ref in ((True, y, lp_t), (False, x, lp_u)):
        plan, _ = jbugs.lower(text, {}, transformed=transformed)
        lp, _, st = COracle(plan).eval_batch(np.array([[th]]))
        assert st[0] == 0 and lp[0] == pytest.approx(ref, rel=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(18))
def test_single_node_models_against_scipy_cuda(k):
    call, dist, support, x = _scipy_cases()[k]
    text = "model { a ~ %s }" % call
    lp_u, lp_t, y = _reference_logp(dist, support, x)
    m = jbugs.compile(text, {}, {"a": x})
    assert jbugs.getparams(m)[0] == pytest.approx(y, rel=1e-13, abs=1e-13)
    assert jbugs.LogDensityProblems.logdensity(m, jbugs.getparams(m)) == pytest.approx(lp_t, rel=1e-11)
    mu = jbugs.settrans(m, False)
    assert jbugs.LogDensityProblems.logdensity(mu, jbugs.getparams(mu)) == pytest.approx(lp_u, rel=1e-11)


# ---- random effects picked by a data index (b[school[i]]): observations regrouped into a padded G x Tmax plate ----
NESTED = """model {
  for (i in 1:n) {
    y[i] ~ %s
    %s <- a + b[school[i]] + s[school[i]] * x[i] + c * x[i]
  }
  for (g in 1:G) {
    b[g] ~ dnorm(0, tau.b)
    s[g] ~ dnorm(m.s, 4)
  }
  a ~ dnorm(0, 0.01)
  c ~ dnorm(0, 0.01)
  m.s ~ dnorm(0, 1)
  tau ~ dgamma(2, 2)
  tau.b ~ dgamma(2, 2)
}"""
NESTED_KINDS = {"normal": ("dnorm(mu[i], tau)", "mu[i]"), "poisson": ("dpois(mu[i])", "log(mu[i])"),
                "binomial": ("dbin(mu[i], m[i])", "logit(mu[i])")}


def nested_data(kind, n, G, seed=0):
    r = np.random.default_rng(seed)
    school = np.concatenate([np.arange(1, G + 1), r.integers(1, G + 1, n - G)]).astype(float)
    r.shuffle(school)
    data = {"n": n, "G": G, "school": school, "x": r.standard_normal(n)}
    if kind == "normal":
        data["y"] = r.standard_normal(n)
    elif kind == "poisson":
        data["y"] = r.poisson(2.0, n).astype(float)
    else:
        data["m"] = r.integers(1, 9, n).astype(float)
        data["y"] = r.binomial(data["m"].astype(int), 0.4).astype(float)
    return data


@pytest.mark.parametrize("kind", list(NESTED_KINDS))
def test_nested_index_plate_matches_node_loop(kind):
    text = NESTED % NESTED_KINDS[kind]
    data = nested_data(kind, 41, 6)
    gm = go.compile(text, data).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    pl = plan.plates[0]
    assert (pl.N, pl.n_real) == (6, 41) and pl.weight is not None and plan.n_obs() == 41
    assert lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 3, seed=1, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g)
        assert grad_err(g[:, c], g_g) < 1e-10


def test_nested_index_with_an_empty_group_is_refused():
    data = nested_data("normal", 30, 5)
    data["school"][data["school"] == 3] = 2.0
    with pytest.raises(jbugs.LoweringError, match="at least one observation"):
        jbugs.lower(NESTED % NESTED_KINDS["normal"], data)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,G", [("normal", 41, 6), ("poisson", 5000, 300), ("binomial", 977, 40)])
def test_nested_index_plate_cuda(kind, n, G):
    from test_gpu_parity import _check
    m = jbugs.compile(NESTED % NESTED_KINDS[kind], nested_data(kind, n, G, seed=n))
    theta = random_thetas(np.zeros(m.plan.D), 70, seed=2, scale=0.2)
    _check(m.cuda, COracle(m.plan), theta)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["rats", "seeds"])
def test_forward_only_kernels_give_the_same_logp(which):
    """logdensity without the gradient runs forward-only instantiations of the large-plate kernels (128-chain CTAs):
    same log densities as the forward+reverse kernels, and as the C restatement."""
    from helpers import synth_rats, synth_seeds
    if which == "rats":
        data, alpha, beta = synth_rats(5000, seed=9)
        m = jbugs.compile(jbugs.examples.RATS, data)
        th0 = np.concatenate([[np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)], np.column_stack([beta, alpha]).ravel()])
    else:
        data, b = synth_seeds(20000, seed=9)
        m = jbugs.compile(jbugs.examples.SEEDS, data)
        th0 = np.concatenate([[np.log(1 / 0.09), -0.84, 1.36, 0.09, -0.55], b])
    theta = random_thetas(th0, 200, seed=4, scale=0.05)
    lp_g, g, st = m.cuda.eval_host(theta, want_grad=True)
    lp_f, _, st_f = m.cuda.eval_host(theta, want_grad=False)
    assert np.array_equal(st, st_f) and rel_err(lp_f, lp_g) < 1e-13
    lp0, _, _ = COracle(m.plan).eval_batch(theta, want_grad=False)
    assert rel_err(lp_f, lp0) < 1e-10


# ---- several plates in one model (two data sets sharing hyper-parameters; different loop extents and families) ----
TWO_PLATES = """model {
  for (i in 1:N1) {
    for (j in 1:T) { Y[i, j] ~ dnorm(mu[i, j], tau.c)
                     mu[i, j] <- alpha[i] + beta.c * (x[j] - xbar) }
    alpha[i] ~ dnorm(alpha.c, alpha.tau)
  }
  for (k in 1:N2) {
    r[k] ~ dbin(p[k], n[k])
    logit(p[k]) <- g0 + beta.c * z[k] + b[k]
    b[k] ~ dnorm(0, alpha.tau)
  }
  for (k in 1:N3) { c[k] ~ dpois(lam[k])
                    log(lam[k]) <- g0 + w[k] }
  alpha.c ~ dnorm(0, 1.0E-4)
  beta.c ~ dnorm(0, 1.0E-2)
  g0 ~ dnorm(0, 1.0E-2)
  tau.c ~ dgamma(1.0E-3, 1.0E-3)
  alpha.tau ~ dgamma(1.0E-1, 1.0E-1)
}"""


def two_plates_data(N1, N2, N3, seed=0):
    r = np.random.default_rng(seed)
    n = r.integers(1, 20, N2).astype(float)
    return {"N1": N1, "T": 4, "N2": N2, "N3": N3, "x": np.array([1.0, 2.0, 4.0, 7.0]), "xbar": 3.5,
            "Y": r.normal(3.0, 1.0, (N1, 4)), "z": r.standard_normal(N2), "n": n,
            "r": r.binomial(n.astype(int), 0.4).astype(float), "w": 0.3 * r.standard_normal(N3),
            "c": r.poisson(1.5, N3).astype(float)}


def test_three_plates_plan_matches_node_loop():
    data = two_plates_data(7, 9, 5)
    gm = go.compile(TWO_PLATES, data).use_generated_order()
    plan, lw = jbugs.lower(TWO_PLATES, data)
    assert len(plan.plates) == 3 and lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 3, seed=1, scale=0.3)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g) and grad_err(g[:, c], g_g) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,C", [((7, 9, 5), 5), ((700, 1300, 90), 150)])
def test_three_plates_cuda(sizes, C):
    from test_gpu_parity import _check
    m = jbugs.compile(TWO_PLATES, two_plates_data(*sizes, seed=3))
    theta = random_thetas(np.zeros(m.plan.D), C, seed=2, scale=0.2)
    _check(m.cuda, COracle(m.plan), theta)
    _check(m.cuda, COracle(m.plan), theta, layout="param_fast")
