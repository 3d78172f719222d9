"""UseAutoMarginalization (J/src/model/evaluation.jl:739-961, J/src/model/bugsmodel.jl:803-876):
finite discrete latents summed out.  The models and the closed-form expectations are the ones of the reference's own
test file (J/test/model/auto_marginalization.jl:216-461,654-745), written in BUGS vocabulary (dnorm takes a precision,
so `Normal(mu, sigma)` becomes `dnorm(mu, tau)` with `tau <- 1 / sigma^2`).

Three implementations must agree: the per-node oracle's recursion over the latents' supports (exponential, the
reference's algorithm without the memo table), the lowered plate plan evaluated by the C restatement, and the CUDA
tape kernel; all of them against the closed form (log-sum-exp over components per observation) from scipy."""
import numpy as np
import pytest
from scipy import stats as S
from scipy.special import logsumexp

import graph_oracle as go
import jbugs
from c_oracle import COracle
from conftest import grad_err

GMM2 = """model {
  w[1] <- %s
  w[2] <- %s
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  sigma[1] ~ dexp(1)
  sigma[2] ~ dexp(1)
  for (k in 1:2) { tau[k] <- 1 / (sigma[k] * sigma[k]) }
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], tau[z[i]])
  }
}"""

GMM3 = """model {
  w[1] <- 0.2
  w[2] <- 0.5
  w[3] <- 0.3
  mu[1] ~ dnorm(-3, 0.04)
  mu[2] ~ dnorm(0, 0.04)
  mu[3] ~ dnorm(3, 0.04)
  for (k in 1:3) {
    sigma[k] ~ dexp(1)
    tau[k] <- 1 / (sigma[k] * sigma[k])
  }
  for (i in 1:N) {
    z[i] ~ dcat(w[1:3])
    y[i] ~ dnorm(mu[z[i]], tau[z[i]])
  }
}"""

GMM_SHARED = """model {
  w[1] <- 0.3
  w[2] <- 0.7
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  sigma ~ dexp(1)
  tau <- 1 / (sigma * sigma)
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], tau)
  }
}"""

SCALAR = """model {
  mu ~ dnorm(0, 1)
  z ~ dcat(w[1:K])
  y ~ dnorm(mu + delta[z], tau)
}"""

# weights that are themselves a function of a parameter (not in the reference's file: the Dirichlet prior it uses
# there is multivariate and outside the plate class)
GMM_W = """model {
  p ~ dbeta(2, 3)
  w[1] <- p
  w[2] <- 1 - p
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], 1.5)
  }
}"""

# a Poisson mixture with a covariate: the latent selects the intercept of a log-linear predictor
POISMIX = """model {
  w[1] <- 0.6
  w[2] <- 0.4
  a[1] ~ dnorm(0, 1)
  a[2] ~ dnorm(1, 1)
  b ~ dnorm(0, 4)
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    log(lam[i]) <- a[z[i]] + b * x[i]
    y[i] ~ dpois(lam[i])
  }
}"""


def mixture_loglik(y, w, mus, sigmas):
    """`mixture_loglikelihood` of the reference's test (auto_marginalization.jl:218-233)"""
    return float(sum(logsumexp([np.log(w[j]) + S.norm(mus[j], sigmas[j]).logpdf(yi) for j in range(len(w))]) for yi in y))


def theta_by_name(names, values):
    return np.array([values[n] for n in names], dtype=float)


def lowered(text, data):
    plan, lw = jbugs.lower(text, data, marginalize=True)
    gm = go.compile(text, data, {}).use_generated_order()
    assert lw.param_names() == gm.marginal_param_names()
    return plan, gm


def check_c_port(plan, gm, theta, tol=1e-12):
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(theta.shape[1]):
        lp_g, g_g = gm.marginal_logdensity_and_gradient(theta[:, c])
        assert st[c] == 0
        assert abs(lp[c] - lp_g) <= tol * max(1.0, abs(lp_g))
        assert grad_err(g[:, c], g_g) < 1e-10
    return lp, g


def test_two_component_mixture_fixed_weights():
    """auto_marginalization.jl:235-288"""
    y = np.array([-1.5, 2.3, -2.1, 1.8])
    data = {"N": 4, "y": y}
    plan, gm = lowered(GMM2 % (0.3, 0.7), data)
    assert plan.D == 4  # dimension(model) == 4: the latents are not parameters
    th = theta_by_name(gm.marginal_param_names(), {"sigma[1]": 0.0, "sigma[2]": 0.0, "mu[2]": 2.0, "mu[1]": -2.0})
    expected = (mixture_loglik(y, [0.3, 0.7], [-2.0, 2.0], [1.0, 1.0]) + S.norm(-2, 5).logpdf(-2.0) + S.norm(2, 5).logpdf(2.0)
                + 2 * S.expon().logpdf(1.0))
    assert gm.evaluate_with_marginalization_values(list(th))["tempered_logjoint"] == pytest.approx(expected, abs=1e-10)
    lp, _ = check_c_port(plan, gm, th.reshape(-1, 1))
    assert lp[0] == pytest.approx(expected, abs=1e-10)


def test_three_component_mixture():
    """auto_marginalization.jl:290-349"""
    y = np.array([-2.5, 0.5, 3.2])
    data = {"N": 3, "y": y}
    plan, gm = lowered(GMM3, data)
    assert plan.D == 6
    vals = {"sigma[1]": 0.0, "sigma[2]": 0.0, "sigma[3]": 0.0, "mu[1]": -3.0, "mu[2]": 0.0, "mu[3]": 3.0}
    th = theta_by_name(gm.marginal_param_names(), vals)
    expected = (mixture_loglik(y, [0.2, 0.5, 0.3], [-3.0, 0.0, 3.0], [1.0] * 3)
                + S.norm(-3, 5).logpdf(-3.0) + S.norm(0, 5).logpdf(0.0) + S.norm(3, 5).logpdf(3.0) + 3 * S.expon().logpdf(1.0))
    lp, _ = check_c_port(plan, gm, th.reshape(-1, 1))
    assert lp[0] == pytest.approx(expected, abs=1e-10)


def test_label_invariance():
    """auto_marginalization.jl:351-389: equal weights, swapping the components' parameters leaves logp unchanged"""
    data = {"N": 4, "y": np.array([1.0, 2.0, -1.0, 3.0])}
    text = GMM2.replace("dnorm(-2, 0.04)", "dnorm(0, 0.01)").replace("dnorm(2, 0.04)", "dnorm(0, 0.01)") % (0.5, 0.5)
    plan, gm = lowered(text, data)
    names = gm.marginal_param_names()
    a = theta_by_name(names, {"sigma[1]": -0.5, "sigma[2]": 0.0, "mu[2]": 3.0, "mu[1]": 1.0})
    b = theta_by_name(names, {"sigma[1]": 0.0, "sigma[2]": -0.5, "mu[2]": 1.0, "mu[1]": 3.0})
    lp, _ = check_c_port(plan, gm, np.stack([a, b], axis=1))
    assert lp[0] == pytest.approx(lp[1], abs=1e-10)


def test_partially_observed_latent():
    """auto_marginalization.jl:391-462: z = [2, missing, 1, missing]; the observed cells take their own branch"""
    y = np.array([1.0, 2.0, -1.0, 3.0])
    data = {"N": 4, "y": y, "z": np.array([2.0, np.nan, 1.0, np.nan])}
    plan, gm = lowered(GMM_SHARED, data)
    assert plan.D == 3
    th = theta_by_name(gm.marginal_param_names(), {"sigma": 0.0, "mu[2]": 2.0, "mu[1]": -2.0})
    w, mu = [0.3, 0.7], [-2.0, 2.0]
    obs = np.log(w[1]) + S.norm(mu[1], 1).logpdf(1.0) + np.log(w[0]) + S.norm(mu[0], 1).logpdf(-1.0)
    marg = sum(logsumexp([np.log(w[k]) + S.norm(mu[k], 1).logpdf(v) for k in range(2)]) for v in (2.0, 3.0))
    expected = obs + marg + S.norm(-2, 5).logpdf(-2.0) + S.norm(2, 5).logpdf(2.0) + S.expon().logpdf(1.0)
    lp, _ = check_c_port(plan, gm, th.reshape(-1, 1))
    assert lp[0] == pytest.approx(expected, abs=1e-10)


def test_gradient_gmm():
    """auto_marginalization.jl:654-703: gradient at the initialised point (dual numbers through the recursion ==
    analytic posterior-weighted reverse sweep of the plan), and central differences as the reference does"""
    y = np.array([-1.8, -2.2, 1.9, 2.1, -1.5, 2.4])
    data = {"N": 6, "y": y}
    plan, gm = lowered(GMM2 % (0.3, 0.7), data)
    th = theta_by_name(gm.marginal_param_names(), {"mu[1]": -2.0, "mu[2]": 2.0, "sigma[1]": np.log(1.1), "sigma[2]": np.log(0.9)})
    lp, g = check_c_port(plan, gm, th.reshape(-1, 1))
    co = COracle(plan)
    eps = 1e-6
    fd = np.zeros_like(th)
    for i in range(th.size):
        e = np.zeros_like(th)
        e[i] = eps
        fd[i] = (co.eval_batch((th + e).reshape(-1, 1))[0][0] - co.eval_batch((th - e).reshape(-1, 1))[0][0]) / (2 * eps)
    assert np.max(np.abs(g[:, 0] - fd) / (np.abs(fd) + 1e-8)) < 5e-5


def test_prior_likelihood_split_and_tempering():
    """auto_marginalization.jl:705-745: a scalar latent picking a data offset; logprior / loglikelihood / temperature"""
    data = {"K": 2, "w": np.array([0.3, 0.7]), "delta": np.array([0.0, 2.0]), "tau": 1.0, "y": 1.5}
    plan, gm = lowered(SCALAR, data)
    assert plan.D == 1
    r = gm.evaluate_with_marginalization_values([0.0], temperature=0.4)
    e_prior = S.norm(0, 1).logpdf(0.0)
    e_lik = logsumexp([np.log(data["w"][i]) + S.norm(data["delta"][i], 1.0).logpdf(1.5) for i in range(2)])
    assert r["logprior"] == pytest.approx(e_prior, abs=1e-10)
    assert r["loglikelihood"] == pytest.approx(e_lik, abs=1e-10)
    assert r["tempered_logjoint"] == pytest.approx(e_prior + 0.4 * e_lik, abs=1e-10)
    co = COracle(plan)
    lp, _, _ = co.eval_batch(np.zeros((1, 1)))
    assert lp[0] == pytest.approx(e_prior + e_lik, abs=1e-10)


def _random_cases():
    rng = np.random.default_rng(11)
    N = 37
    z = rng.integers(0, 2, N)
    y = np.where(z == 0, rng.normal(-2, 1.0, N), rng.normal(2, 0.7, N))
    x = rng.standard_normal(N)
    yp = rng.poisson(np.exp(np.where(z == 0, 0.2, 1.3) + 0.3 * x)).astype(float)
    return {
        "gmm2": (GMM2 % (0.3, 0.7), {"N": N, "y": y}),
        "gmm3": (GMM3, {"N": N, "y": y}),
        "gmm_weights_from_a_parameter": (GMM_W, {"N": N, "y": y}),
        "poisson_mixture": (POISMIX, {"N": N, "y": yp, "x": x}),
    }


@pytest.mark.parametrize("name", ["gmm2", "gmm3", "gmm_weights_from_a_parameter", "poisson_mixture"])
def test_random_thetas_node_oracle_vs_plan(name):
    text, data = _random_cases()[name]
    data = dict(data)
    data["N"] = 5   # the recursion over the joint support is exponential in N
    for k in ("y", "x"):
        if k in data:
            data[k] = data[k][:5]
    plan, gm = lowered(text, data)
    rng = np.random.default_rng(3)
    theta = 0.4 * rng.standard_normal((plan.D, 4))
    check_c_port(plan, gm, theta)


def test_edge_cases_without_latents():
    """auto_marginalization.jl:592-652: a model without discrete latents, and one whose discrete nodes are all observed,
    evaluate under the marginalising mode exactly as without it"""
    text = "model { mu ~ dnorm(0, 0.01)\n sigma ~ dexp(1)\n tau <- 1 / (sigma * sigma)\n for (i in 1:N) { y[i] ~ dnorm(mu, tau) } }"
    data = {"N": 3, "y": np.array([1.0, 2.0, 3.0])}
    plan, lw = jbugs.lower(text, data, marginalize=True)
    plain, _ = jbugs.lower(text, data)
    assert plan.D == 2 and plan.to_bytes() == plain.to_bytes()
    text2 = "model { p ~ dbeta(1, 1)\n for (i in 1:N) { x[i] ~ dbern(p) } }"
    data2 = {"N": 5, "x": np.array([1.0, 0.0, 1.0, 1.0, 0.0])}
    plan2, _ = jbugs.lower(text2, data2, marginalize=True)
    assert plan2.D == 1
    lp, _, st = COracle(plan2).eval_batch(np.zeros((1, 1)))
    expected = S.beta(1, 1).logpdf(0.5) + 3 * np.log(0.5) + 2 * np.log(0.5) + np.log(0.25)
    assert st[0] == 0 and lp[0] == pytest.approx(expected, abs=1e-10)


def test_unsupported_shapes_raise():
    hmm = """model {
      z[1] ~ dcat(p0[1:2])
      for (t in 2:T) { z[t] ~ dcat(A[z[t-1], 1:2]) }
      for (t in 1:T) { y[t] ~ dnorm(mu[z[t]], 1) }
      mu[1] ~ dnorm(0, 1)
      mu[2] ~ dnorm(3, 1)
    }"""
    data = {"T": 3, "y": np.array([0.1, 2.9, 3.2]), "p0": np.array([0.5, 0.5]), "A": np.array([[0.9, 0.1], [0.2, 0.8]])}
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(hmm, data, marginalize=True)


def test_mode_object():
    data = {"N": 4, "y": np.array([-1.5, 2.3, -2.1, 1.8])}
    with pytest.raises(jbugs.LoweringError):
        jbugs.compile(GMM2 % (0.3, 0.7), data)  # z[i] as a discrete theta entry steering mu[z[i]]: not on the device
    mm = jbugs.compile(GMM2 % (0.3, 0.7), data, evaluation_mode=jbugs.UseAutoMarginalization())
    assert jbugs.LogDensityProblems.dimension(mm) == 4
    assert isinstance(mm.evaluation_mode, jbugs.UseCUDABatch)  # the marginalising mode is a CUDA batch mode


# ---- the CUDA tape kernel ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gmm2", "gmm3", "gmm_weights_from_a_parameter", "poisson_mixture"])
@pytest.mark.parametrize("layout", ["chain_fast", "param_fast"])
def test_cuda_matches_c_port(name, layout):
    text, data = _random_cases()[name]
    m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
    rng = np.random.default_rng(5)
    theta = 0.4 * rng.standard_normal((m.plan.D, 70))
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    assert (st == 0).all() and (st0 == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10


@pytest.mark.gpu
def test_cuda_reference_cases():
    """the reference's closed-form expectations through the C ABI"""
    y = np.array([-1.5, 2.3, -2.1, 1.8])
    m = jbugs.compile(GMM2 % (0.3, 0.7), {"N": 4, "y": y}, evaluation_mode=jbugs.UseAutoMarginalization())
    assert jbugs.LogDensityProblems.dimension(m) == 4
    th = theta_by_name(m.lowering.param_names(), {"sigma[1]": 0.0, "sigma[2]": 0.0, "mu[2]": 2.0, "mu[1]": -2.0})
    expected = (mixture_loglik(y, [0.3, 0.7], [-2.0, 2.0], [1.0, 1.0]) + S.norm(-2, 5).logpdf(-2.0) + S.norm(2, 5).logpdf(2.0)
                + 2 * S.expon().logpdf(1.0))
    assert jbugs.LogDensityProblems.logdensity(m, th) == pytest.approx(expected, abs=1e-10)
    data = {"K": 2, "w": np.array([0.3, 0.7]), "delta": np.array([0.0, 2.0]), "tau": 1.0, "y": 1.5}
    ms = jbugs.compile(SCALAR, data, evaluation_mode=jbugs.UseAutoMarginalization())
    r = jbugs.evaluate_with_values(ms, np.zeros(1), temperature=0.4)
    e_prior = S.norm(0, 1).logpdf(0.0)
    e_lik = logsumexp([np.log(data["w"][i]) + S.norm(data["delta"][i], 1.0).logpdf(1.5) for i in range(2)])
    assert r["logprior"] == pytest.approx(e_prior, abs=1e-10)
    assert r["loglikelihood"] == pytest.approx(e_lik, abs=1e-10)
    assert r["tempered_logjoint"] == pytest.approx(e_prior + 0.4 * e_lik, abs=1e-10)
    # partially observed latent
    yp = np.array([1.0, 2.0, -1.0, 3.0])
    mp = jbugs.compile(GMM_SHARED, {"N": 4, "y": yp, "z": np.array([2.0, np.nan, 1.0, np.nan])},
                       evaluation_mode=jbugs.UseAutoMarginalization())
    th = theta_by_name(mp.lowering.param_names(), {"sigma": 0.0, "mu[2]": 2.0, "mu[1]": -2.0})
    lp0, g0, _ = COracle(mp.plan).eval_batch(th.reshape(-1, 1))
    lp, g = jbugs.LogDensityProblems.logdensity_and_gradient(mp, th)
    assert lp == pytest.approx(lp0[0], rel=1e-12)
    assert grad_err(g, g0[:, 0]) < 1e-10


@pytest.mark.gpu
def test_cuda_large_mixture_against_closed_form():
    """10^5 observations x 64 chains: the plate's log-sum-exp against numpy's closed form, gradient by the C port on 4 chains"""
    rng = np.random.default_rng(2)
    N = 100_000
    z = rng.random(N) < 0.3
    y = np.where(z, rng.normal(-2, 1.0, N), rng.normal(2, 0.7, N))
    m = jbugs.compile(GMM2 % (0.3, 0.7), {"N": N, "y": y}, evaluation_mode=jbugs.UseAutoMarginalization())
    names = m.lowering.param_names()
    base = theta_by_name(names, {"sigma[1]": 0.0, "sigma[2]": np.log(0.7), "mu[2]": 2.0, "mu[1]": -2.0})
    theta = base[:, None] + 0.02 * rng.standard_normal((4, 64))
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    assert (st == 0).all()
    ix = {n: i for i, n in enumerate(names)}
    for c in (0, 17, 63):
        mu = [theta[ix["mu[1]"], c], theta[ix["mu[2]"], c]]
        sg = [np.exp(theta[ix["sigma[1]"], c]), np.exp(theta[ix["sigma[2]"], c])]
        ll = logsumexp(np.stack([np.log(w) + S.norm(mu[k], sg[k]).logpdf(y) for k, w in enumerate((0.3, 0.7))]), axis=0).sum()
        pr = (S.norm(-2, 5).logpdf(mu[0]) + S.norm(2, 5).logpdf(mu[1]) + sum(S.expon().logpdf(s) + np.log(s) for s in sg))
        assert lp[c] == pytest.approx(ll + pr, rel=1e-11)
    lp0, g0, _ = COracle(m.plan).eval_batch(theta[:, :4])
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g[:, :4] - g0) / scale) < 1e-10


# ---- Bernoulli latents: the support {0, 1} summed out like a two-state dcat (evaluation.jl:860-890 enumerates the
# support of any finite discrete node).  The latent may enter arithmetic directly, or through a derived subscript
# (`s[i] <- z[i] + 1; mu[s[i]]`, the idiom of the Cervix example, J/src/BUGSExamples/Volume_2/08_Cervix.jl:5-10)
BERN_ARITH = """model {
  p ~ dbeta(2, 3)
  mu0 ~ dnorm(-2, 0.04)
  delta ~ dnorm(4, 0.04)
  for (i in 1:N) {
    z[i] ~ dbern(p)
    y[i] ~ dnorm(mu0 + delta * z[i], 1.5)
  }
}"""

BERN_SUBSCRIPT = """model {
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  for (i in 1:N) {
    z[i] ~ dbern(q[i])
    s[i] <- z[i] + 1
    y[i] ~ dnorm(mu[s[i]], 1.5)
  }
}"""


def _bern_cases(N=5):
    rng = np.random.default_rng(23)
    y = 2.0 * rng.standard_normal(N)
    return {"bern_arith": (BERN_ARITH, {"N": N, "y": y}),
            "bern_subscript": (BERN_SUBSCRIPT, {"N": N, "y": y, "q": 0.1 + 0.8 * rng.random(N)})}


@pytest.mark.parametrize("name", ["bern_arith", "bern_subscript"])
def test_bernoulli_latents_node_oracle_vs_plan(name):
    text, data = _bern_cases()[name]
    plan, gm = lowered(text, data)
    rng = np.random.default_rng(4)
    theta = 0.4 * rng.standard_normal((plan.D, 4))
    lp, _ = check_c_port(plan, gm, theta)
    if name == "bern_subscript":   # closed form: a two-component mixture with per-observation weights
        for c in range(4):
            mu = theta[:, c]       # generated order: mu[1], mu[2]
            names = gm.marginal_param_names()
            m1, m2 = mu[names.index("mu[1]")], mu[names.index("mu[2]")]
            sd = 1 / np.sqrt(1.5)
            ll = sum(logsumexp([np.log(1 - q) + S.norm(m1, sd).logpdf(yi), np.log(q) + S.norm(m2, sd).logpdf(yi)])
                     for yi, q in zip(data["y"], data["q"]))
            pr = S.norm(-2, 5).logpdf(m1) + S.norm(2, 5).logpdf(m2)
            assert lp[c] == pytest.approx(ll + pr, abs=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bern_arith", "bern_subscript"])
def test_cuda_bernoulli_latents(name):
    text, data = _bern_cases(N=41)[name]
    m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
    rng = np.random.default_rng(6)
    theta = 0.4 * rng.standard_normal((m.plan.D, 70))
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    assert (st == 0).all() and (st0 == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10


# ---- Cervix (vol. 2): the true exposure x[i] is observed for 115 of 2044 subjects; the missing ones are Bernoulli
# latents with TWO observed children each (d[i] through the logistic model, w[i] through phi[x[i] + 1, d[i] + 1]).
# The second child is folded into the state's log weight, so a cell is still one sum over the latent's states.
def _cervix_data(sel=None):
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "examples_vol2.json")))["cervix"]["data"]
    full = {k: (np.array(v, dtype=float) if isinstance(v, list) else v) for k, v in fx.items()}
    if sel is None:
        return full
    return {"N": len(sel), **{k: full[k][sel] for k in ("x", "d", "w")}}


def _cervix_closed_form(data, names, theta):
    """log joint in unconstrained space, written from the model text: per subject sum over x in {0, 1} (or the observed x)"""
    ix = {n: i for i, n in enumerate(names)}
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))   # noqa: E731
    q = sig(theta[ix["q"]])
    phi = np.array([[sig(theta[ix[f"phi[{j},{k}]"]]) for k in (1, 2)] for j in (1, 2)])
    b0, b = theta[ix["beta0C"]], theta[ix["beta"]]
    x, d, w = data["x"], data["d"].astype(int), data["w"].astype(int)
    terms = []
    for xv in (0, 1):
        p = sig(b0 + b * xv)
        lt = (np.log(q) if xv else np.log1p(-q)) + np.where(d == 1, np.log(p), np.log1p(-p)) \
            + np.where(w == 1, np.log(phi[xv, d]), np.log1p(-phi[xv, d]))
        terms.append(np.where(np.isnan(x) | (x == xv), lt, -np.inf))
    lik = logsumexp(np.stack(terms), axis=0).sum()
    prior = S.norm(0, np.sqrt(1e5)).logpdf(b0) + S.norm(0, np.sqrt(1e5)).logpdf(b)   # the five dunif(0, 1) priors are 0
    logjac = sum(np.log(v) + np.log1p(-v) for v in [q] + list(phi.ravel()))
    return lik + prior + logjac


def test_cervix_subset_against_the_node_oracle():
    from jbugs import examples as ex
    sel = np.r_[0:3, 40:43, 200:203, 2040:2043]   # 6 observed exposures, 6 latent ones
    data = _cervix_data(sel)
    plan, gm = lowered(ex.CERVIX, data)
    assert plan.D == 7 and len(plan.plates) == 1
    rng = np.random.default_rng(0)
    theta = 0.4 * rng.standard_normal((plan.D, 3))
    lp, _ = check_c_port(plan, gm, theta)
    names = gm.marginal_param_names()
    for c in range(3):
        assert lp[c] == pytest.approx(_cervix_closed_form(data, names, theta[:, c]), abs=1e-10)


def test_cervix_full_size_closed_form():
    from jbugs import examples as ex
    data = _cervix_data()
    plan, lw = jbugs.lower(ex.CERVIX, data, marginalize=True)
    names = lw.param_names()
    assert sorted(names) == sorted(["q", "beta0C", "beta", "phi[1,1]", "phi[1,2]", "phi[2,1]", "phi[2,2]"])
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(ex.CERVIX, data)   # without marginalisation the 1929 missing x[i] are discrete entries of theta
    rng = np.random.default_rng(1)
    theta = 0.5 * rng.standard_normal((plan.D, 4))
    lp, g, st = COracle(plan).eval_batch(theta)
    assert (st == 0).all()
    for c in range(4):
        assert lp[c] == pytest.approx(_cervix_closed_form(data, names, theta[:, c]), rel=1e-12)
    # gradient against central differences of the closed form (auto_marginalization.jl:654-703 checks the same way)
    h = 1e-6
    for k in range(plan.D):
        e = np.zeros(plan.D)
        e[k] = h
        fd = (_cervix_closed_form(data, names, theta[:, 0] + e) - _cervix_closed_form(data, names, theta[:, 0] - e)) / (2 * h)
        assert g[k, 0] == pytest.approx(fd, rel=1e-5, abs=1e-5)


def test_partially_observed_bernoulli_latent():
    rng = np.random.default_rng(9)
    y = 2.0 * rng.standard_normal(5)
    z = np.array([1.0, np.nan, 0.0, np.nan, np.nan])
    for text, data in ((BERN_ARITH, {"N": 5, "y": y, "z": z}),
                       (BERN_SUBSCRIPT, {"N": 5, "y": y, "q": 0.1 + 0.8 * rng.random(5), "z": z})):
        plan, gm = lowered(text, data)
        check_c_port(plan, gm, 0.4 * rng.standard_normal((plan.D, 3)))


@pytest.mark.gpu
def test_cuda_cervix():
    from jbugs import examples as ex
    data = _cervix_data()
    m = jbugs.compile(ex.CERVIX, data, evaluation_mode=jbugs.UseAutoMarginalization())
    rng = np.random.default_rng(8)
    theta = 0.5 * rng.standard_normal((m.plan.D, 200))
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    assert (st == 0).all() and (st0 == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10
    names = m.lowering.param_names()
    assert lp[0] == pytest.approx(_cervix_closed_form(data, names, theta[:, 0]), rel=1e-12)


# ---- wider mixtures: an expression plate keeps up to 24 global values per chain (MAX_GREF), so eight components with
# their own means and standard deviations (16 referenced globals) fit; linear-predictor plates keep 12
def _gmm_k(K, N=120, seed=0):
    text = ("model {\n" + "".join(f"  w[{k}] <- {1.0 / K}\n" for k in range(1, K + 1)) +
            f"  for (k in 1:{K}) {{ mu[k] ~ dnorm(0, 0.04)\n    sigma[k] ~ dexp(1)\n    tau[k] <- 1 / (sigma[k] * sigma[k]) }}\n"
            f"  for (i in 1:N) {{ z[i] ~ dcat(w[1:{K}])\n    y[i] ~ dnorm(mu[z[i]], tau[z[i]]) }}\n}}")
    rng = np.random.default_rng(seed)
    y = rng.normal(rng.integers(0, K, N) * 2.0 - K + 1, 0.5)
    return text, {"N": N, "y": y}


def test_eight_component_mixture_plan_against_closed_form():
    text, data = _gmm_k(8)
    plan, lw = jbugs.lower(text, data, marginalize=True)
    names = lw.param_names()
    assert plan.D == 16
    rng = np.random.default_rng(1)
    theta = 0.3 * rng.standard_normal((16, 3))
    lp, _, st = COracle(plan).eval_batch(theta)
    ix = {n: i for i, n in enumerate(names)}
    for c in range(3):
        mus = [theta[ix[f"mu[{k}]"], c] for k in range(1, 9)]
        sgs = [np.exp(theta[ix[f"sigma[{k}]"], c]) for k in range(1, 9)]
        e = mixture_loglik(data["y"], [1 / 8] * 8, mus, sgs)
        e += sum(S.norm(0, 5).logpdf(m) for m in mus) + sum(S.expon().logpdf(s) + np.log(s) for s in sgs)
        assert st[c] == 0 and lp[c] == pytest.approx(e, rel=1e-12)


@pytest.mark.gpu
def test_cuda_eight_component_mixture():
    text, data = _gmm_k(8)
    m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
    theta = 0.3 * np.random.default_rng(2).standard_normal((m.plan.D, 70))
    lp0, g0, _ = COracle(m.plan).eval_batch(theta)
    for compiled in (False, True):
        if compiled:
            m.cuda.specialize()
            assert "expression_tape_specialised_at_run_time" in m.cuda.describe()
        lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
        assert (st == 0).all()
        assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
        scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
        assert np.max(np.abs(g - g0) / scale) < 1e-10


# ---- generated quantities of a marginalised model: the summed-out latents are recovered from p(z | theta, y) before
# quantities that depend on them are forward-sampled (J/test/model/auto_marginalization.jl:975-1041, "case 3"; the
# reference pins the empirical distribution of the recovered latents and E[g] against the analytic posterior)
GMM_GQ = """model {
  w[1] <- 0.3
  w[2] <- 0.7
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], 1.0)
  }
  g ~ dnorm(mu[z[1]], 1)
}"""


def test_recovered_latents_follow_their_posterior():
    from jbugs import gq
    data = {"N": 3, "y": np.array([-0.3, 0.6, 0.1])}
    plan, lw = jbugs.lower(GMM_GQ, data, marginalize=True)
    assert sorted(lw.param_names()) == ["mu[1]", "mu[2]"]        # g is a generated quantity, z is summed out
    n, mu = 40000, [-0.5, 1.0]
    out = gq.generated_quantities(lw, {"mu[1]": np.full(n, mu[0]), "mu[2]": np.full(n, mu[1])}, seed=1)
    assert out["z"].shape == (3, n) and out["g"].shape == (n,)
    post = []
    for i, y in enumerate(data["y"]):
        w = np.array([0.3 * S.norm(mu[0], 1).pdf(y), 0.7 * S.norm(mu[1], 1).pdf(y)])
        post.append(w / w.sum())
        assert abs((out["z"][i] == 2).mean() - post[-1][1]) < 0.02     # the reference's tolerance
    assert set(np.unique(out["z"])) == {1.0, 2.0}
    assert out["g"].mean() == pytest.approx(post[0] @ np.array(mu), abs=0.05)


def test_recovered_bernoulli_latents_with_two_children():
    """Cervix subset: x[i] | theta, d[i], w[i]; observed exposures keep their values"""
    from jbugs import examples as ex
    from jbugs import gq
    sel = np.r_[0:3, 40:43, 200:203, 2040:2043]
    data = _cervix_data(sel)
    plan, lw = jbugs.lower(ex.CERVIX, data, marginalize=True)
    n = 30000
    val = {"q": 0.4, "beta0C": -0.3, "beta": 0.9, "phi[1,1]": 0.2, "phi[1,2]": 0.35, "phi[2,1]": 0.6, "phi[2,2]": 0.8}
    out = gq.generated_quantities(lw, {k: np.full(n, v) for k, v in val.items()}, seed=3)
    x = out["x"]
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))   # noqa: E731
    for i in range(len(sel)):
        if not np.isnan(data["x"][i]):
            assert (x[i] == data["x"][i]).all()
            continue
        d, w = int(data["d"][i]), int(data["w"][i])
        wt = []
        for xv in (0, 1):
            p = sig(val["beta0C"] + val["beta"] * xv)
            ph = val[f"phi[{xv + 1},{d + 1}]"]
            wt.append((val["q"] if xv else 1 - val["q"]) * (p if d else 1 - p) * (ph if w else 1 - ph))
        assert abs((x[i] == 1).mean() - wt[1] / sum(wt)) < 0.02
    assert "gamma1" in out and out["gamma1"].shape == (n,)


@pytest.mark.gpu
def test_nuts_on_a_marginalised_mixture_and_latent_recovery():
    """auto_marginalization.jl:747-813 (AutoMarg + NUTS runs, fewer dimensions than the graph model, finite draws) on the
    device sampler, plus what the CPU reference cannot afford in a smoke test: the posterior means, and the recovered
    latents classifying the observations"""
    import jbugs.hmc as hmc
    rng = np.random.default_rng(42)
    y = np.concatenate([rng.normal(-2, 1, 50), rng.normal(2, 1, 50)])
    m = jbugs.compile(GMM2 % (0.3, 0.7), {"N": 100, "y": y}, evaluation_mode=jbugs.UseAutoMarginalization())
    assert jbugs.LogDensityProblems.dimension(m) == 4          # the graph model has 104
    names = m.lowering.param_names()
    init = theta_by_name(names, {"mu[1]": -2.0, "mu[2]": 2.0, "sigma[1]": 0.0, "sigma[2]": 0.0})
    ch = hmc.sample(m, hmc.NUTS(max_depth=6), n_samples=200, n_adapts=200, n_chains=64, seed=7, init=init)
    assert np.isfinite(ch.samples).all() and np.abs(ch.samples).max() < 20
    assert ch.mean("mu[1]") == pytest.approx(y[:50].mean(), abs=0.5)
    assert ch.mean("mu[2]") == pytest.approx(y[50:].mean(), abs=0.5)
    gq = ch.generated_quantities()
    z = gq["z"].reshape(100, -1)
    assert (z[:50] == 1).mean() > 0.85 and (z[50:] == 2).mean() > 0.85
    with pytest.raises(ValueError):
        hmc.sample(m, hmc.NUTS(max_depth=4), n_samples=5, n_adapts=5, n_chains=32, init=init).generated_quantities(device=True)


# ---- every family a latent's further observed child may have (folded into the state's log weight as tape arithmetic)
def _two_children(kind):
    second = {"dnorm": "c[i] ~ dnorm(a[z[i]], 2.0)",
              "dpois": "c[i] ~ dpois(lam[i])\n    log(lam[i]) <- a[z[i]]",
              "dbin": "c[i] ~ dbin(pr[i], n[i])\n    logit(pr[i]) <- a[z[i]]",
              "dbern": "c[i] ~ dbern(pr[i])\n    logit(pr[i]) <- a[z[i]]"}[kind]
    text = f"""model {{
  w[1] <- 0.4
  w[2] <- 0.6
  mu[1] ~ dnorm(-1, 0.25)
  mu[2] ~ dnorm(1, 0.25)
  a[1] ~ dnorm(0, 1)
  a[2] ~ dnorm(0.5, 1)
  for (i in 1:N) {{
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], 1.5)
    {second}
  }}
}}"""
    rng = np.random.default_rng(5)
    N = 5
    data = {"N": N, "y": rng.standard_normal(N)}
    if kind == "dnorm":
        data["c"] = rng.standard_normal(N)
    elif kind == "dpois":
        data["c"] = rng.poisson(1.5, N).astype(float)
    elif kind == "dbin":
        data["n"] = rng.integers(3, 9, N).astype(float)
        data["c"] = rng.binomial(data["n"].astype(int), 0.5).astype(float)
    else:
        data["c"] = rng.integers(0, 2, N).astype(float)
    return text, data


@pytest.mark.parametrize("kind", ["dnorm", "dpois", "dbin", "dbern"])
def test_second_child_families_against_the_node_oracle(kind):
    text, data = _two_children(kind)
    plan, gm = lowered(text, data)
    assert len(plan.plates) == 1 and plan.D == 4
    theta = 0.5 * np.random.default_rng(8).standard_normal((plan.D, 4))
    check_c_port(plan, gm, theta)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dnorm", "dpois", "dbin", "dbern"])
def test_cuda_second_child_families(kind):
    text, data = _two_children(kind)
    m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
    theta = 0.5 * np.random.default_rng(9).standard_normal((m.plan.D, 70))
    lp0, g0, _ = COracle(m.plan).eval_batch(theta)
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    assert (st == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10


# ---- randomly assembled marginalised models: latent kind x number of states x observation family x optional second child
# x partial observation, node-oracle recursion == plan through the C restatement (and CUDA below)
def _random_marginalised(seed):
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(3, 6))
    bern = bool(rng.integers(0, 2))
    K = 2 if bern else int(rng.integers(2, 4))
    fam = ["dnorm", "dpois", "dbern"][int(rng.integers(0, 3))]
    child = [None, "dnorm", "dpois", "dbin"][int(rng.integers(0, 4))]
    partial = bool(rng.integers(0, 2))
    weights_param = bool(rng.integers(0, 2)) and K == 2
    lines = ["model {"]
    if bern:
        lines.append("  p ~ dbeta(2, 2)" if weights_param else "  p <- 0.35")
        zdef, zidx = "z[i] ~ dbern(p)", "s[i]"
    else:
        if weights_param:
            lines += ["  p ~ dbeta(2, 2)", "  w[1] <- p", "  w[2] <- 1 - p"]
        else:
            cuts = np.sort(rng.choice(np.arange(100, 900), K - 1, replace=False))
            w = np.diff(np.r_[0, cuts, 1000])          # integers summing to 1000: three-digit weights summing to 1 exactly
            lines += [f"  w[{k + 1}] <- 0.{w[k]:03d}" for k in range(K)]
        zdef, zidx = f"z[i] ~ dcat(w[1:{K}])", "z[i]"
    lines += [f"  a[{k + 1}] ~ dnorm({k - 0.5:.1f}, 0.5)" for k in range(K)]
    lines.append("  b ~ dnorm(0, 1)")
    lines.append("  for (i in 1:N) {")
    lines.append("    " + zdef)
    if bern:
        lines.append("    s[i] <- z[i] + 1")
    if fam == "dnorm":
        lines.append(f"    y[i] ~ dnorm(a[{zidx}] + b * x[i], 1.3)")
    elif fam == "dpois":
        lines += [f"    y[i] ~ dpois(lam[i])", f"    log(lam[i]) <- a[{zidx}] + b * x[i]"]
    else:
        lines += [f"    y[i] ~ dbern(q[i])", f"    logit(q[i]) <- a[{zidx}] + b * x[i]"]
    if child == "dnorm":
        lines.append(f"    c[i] ~ dnorm(b * a[{zidx}], 2.0)")
    elif child == "dpois":
        lines += ["    c[i] ~ dpois(nu[i])", f"    log(nu[i]) <- 0.5 * a[{zidx}]"]
    elif child == "dbin":
        lines += ["    c[i] ~ dbin(r[i], m[i])", f"    logit(r[i]) <- a[{zidx}] - b"]
    lines += ["  }", "}"]
    data = {"N": N, "x": rng.standard_normal(N)}
    data["y"] = {"dnorm": rng.standard_normal(N), "dpois": rng.poisson(1.5, N).astype(float),
                 "dbern": rng.integers(0, 2, N).astype(float)}[fam]
    if child == "dnorm":
        data["c"] = rng.standard_normal(N)
    elif child == "dpois":
        data["c"] = rng.poisson(1.0, N).astype(float)
    elif child == "dbin":
        data["m"] = rng.integers(2, 7, N).astype(float)
        data["c"] = rng.binomial(data["m"].astype(int), 0.4).astype(float)
    if partial:
        z = np.full(N, np.nan)
        z[0] = float(rng.integers(0, 2)) if bern else float(rng.integers(1, K + 1))
        data["z"] = z
    return "\n".join(lines), data, f"bern={bern} K={K} fam={fam} child={child} partial={partial} wparam={weights_param}"


@pytest.mark.parametrize("seed", list(range(32)))
def test_random_marginalised_model_against_the_node_oracle(seed):
    text, data, desc = _random_marginalised(seed)
    plan, gm = lowered(text, data)
    theta = 0.5 * np.random.default_rng(seed).standard_normal((plan.D, 3))
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        l0, g0 = gm.marginal_logdensity_and_gradient(theta[:, c])
        assert st[c] == 0 and abs(lp[c] - l0) <= 1e-11 * max(1.0, abs(l0)), (desc, text)
        assert grad_err(g[:, c], g0) < 1e-10, (desc, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(16)))
def test_random_marginalised_model_cuda(seed):
    text, data, desc = _random_marginalised(seed)
    m = jbugs.compile(text, data, evaluation_mode=jbugs.UseAutoMarginalization())
    theta = 0.5 * np.random.default_rng(seed).standard_normal((m.plan.D, 40))
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    assert (st == 0).all() and (st0 == 0).all(), desc
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12, desc
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10, desc
