"""Truncated priors, the BUGS suffix `T(lower, upper)` (`truncated(dist, lower, upper)`: J/src/parser/bugs_parser.jl:470-500;
Magnesium's `tau.sqrd[6] ~ dnorm(0, p0) T(0, )`, J/src/BUGSExamples/Volume_1/06_Magnesium.jl:63): half-normal and
truncated-normal priors of scalar parameters with constant arguments.  The bounds are the prior's support (bijector),
the normaliser log(cdf(upper) - cdf(lower)) a constant: scipy.stats.truncnorm == per-node oracle == lowered plan through
the C restatement == CUDA."""
import numpy as np
import pytest
from scipy import stats as S

import graph_oracle as go
import jbugs
from c_oracle import COracle
from conftest import grad_err

TEXT = """model {
  s2 ~ dnorm(0, 0.25) T(0, )
  m ~ dnorm(1, 4) T(-1, 2.5)
  u ~ dnorm(0, 1) T(, 0.5)
  for (k in 1:2) { c[k] ~ dnorm(3, 0.5) T(2, ) }
  for (i in 1:N) { y[i] ~ dnorm(m + u + c[1] * x[i] + c[2], s2) }
}"""
rng = np.random.default_rng(31)
DATA = {"N": 6, "y": rng.standard_normal(6) + 4.0, "x": rng.standard_normal(6)}


def closed_form(names, theta):
    """log joint in unconstrained space from scipy's truncated normal + the bijectors' log-Jacobians"""
    ix = {n: i for i, n in enumerate(names)}
    t = {n: theta[i] for n, i in ix.items()}
    s2 = np.exp(t["s2"])
    sg = 1.0 / (1.0 + np.exp(-t["m"]))
    m = -1.0 + 3.5 * sg
    u = 0.5 - np.exp(t["u"])
    c = [2.0 + np.exp(t[f"c[{k}]"]) for k in (1, 2)]
    lp = S.truncnorm(0, np.inf, loc=0, scale=2.0).logpdf(s2) + t["s2"]
    lp += S.truncnorm((-1 - 1) / 0.5, (2.5 - 1) / 0.5, loc=1, scale=0.5).logpdf(m) + np.log((m + 1) * (2.5 - m) / 3.5)
    lp += S.truncnorm(-np.inf, 0.5, loc=0, scale=1.0).logpdf(u) + t["u"]
    sd = 1 / np.sqrt(0.5)
    for k in (0, 1):
        lp += S.truncnorm((2 - 3) / sd, np.inf, loc=3, scale=sd).logpdf(c[k]) + t[f"c[{k + 1}]"]
    lp += S.norm(m + u + c[0] * DATA["x"] + c[1], 1 / np.sqrt(s2)).logpdf(DATA["y"]).sum()
    return lp


def test_oracle_and_plan_against_scipy():
    gm = go.compile(TEXT, DATA, {}).use_generated_order()
    plan, lw = jbugs.lower(TEXT, DATA)
    names = lw.param_names()
    assert names == gm.param_names() and plan.D == 5
    theta = 0.7 * np.random.default_rng(2).standard_normal((plan.D, 5))
    lp, g, st = COracle(plan).eval_batch(theta)
    assert (st == 0).all()
    for c in range(theta.shape[1]):
        l0, g0 = gm.logdensity_and_gradient(theta[:, c])
        assert l0 == pytest.approx(closed_form(names, theta[:, c]), abs=1e-11)
        assert lp[c] == pytest.approx(l0, abs=1e-11)
        assert grad_err(g[:, c], g0) < 1e-10


def test_support_bounds_drive_the_bijector():
    plan, lw = jbugs.lower(TEXT, DATA)
    by = {g.name: g for g in plan.globals}
    assert (by["s2"].transform, by["s2"].lb, by["s2"].ub) == (jbugs.plan.TR_LOG, 0.0, np.inf)
    assert (by["m"].transform, by["m"].lb, by["m"].ub) == (jbugs.plan.TR_LOGIT, -1.0, 2.5)
    assert (by["u"].transform, by["u"].lb, by["u"].ub) == (jbugs.plan.TR_UPPER, -np.inf, 0.5)
    assert by["c[2]"].transform == jbugs.plan.TR_LOG and by["c[2]"].lb == 2.0


@pytest.mark.parametrize("text", [
    "model { a ~ dpois(2) T(1, )\n for (i in 1:N) { y[i] ~ dnorm(a, 1) } }",                     # a discrete prior
    "model { t ~ dgamma(1, 1)\n a ~ dnorm(0, t) T(0, )\n for (i in 1:N) { y[i] ~ dnorm(a, 1) } }",   # argument is a parameter
    "model { a ~ dnorm(0, 1)\n for (i in 1:N) { y[i] ~ dnorm(a, 1) T(0, ) } }",                  # truncated observation
    "model { for (i in 1:N) { b[i] ~ dnorm(0, 1) T(0, )\n y[i] ~ dnorm(b[i], 1) } }",            # parameter of a plate
])
def test_what_is_not_lowered_raises(text):
    # (N = 20: a short vector b[1:6] would be written out element by element as six truncated scalar parameters by the
    # second reading of the lowering, which is fine)
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(text, {"N": 20, "y": np.linspace(-1.0, 1.0, 20)})


def test_untransformed_space_is_refused():
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(TEXT, DATA, transformed=False)


@pytest.mark.gpu
def test_cuda_matches_c_port_and_scipy():
    m = jbugs.compile(TEXT, DATA)
    names = m.lowering.param_names()
    theta = 0.7 * np.random.default_rng(3).standard_normal((m.plan.D, 70))
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    assert (st == 0).all() and (st0 == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    scale = np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max(axis=0, keepdims=True))
    assert np.max(np.abs(g - g0) / scale) < 1e-10
    assert lp[0] == pytest.approx(closed_form(names, theta[:, 0]), abs=1e-10)


# ---- every family whose truncation is provided (include/jbugs_trunc.h), against scipy's own truncation
FAMILIES = {
    # BUGS prior, bounds, scipy distribution, (bijector of the truncated support: value and log-Jacobian from theta)
    "dgamma": ("dgamma(2.5, 1.5) T(0.4, 3.0)", S.gamma(2.5, scale=1 / 1.5), (0.4, 3.0)),
    "dgamma_lower": ("dgamma(0.7, 0.5) T(1.0, )", S.gamma(0.7, scale=2.0), (1.0, np.inf)),
    "dexp": ("dexp(0.8) T(, 2.5)", S.expon(scale=1 / 0.8), (0.0, 2.5)),
    "dlnorm": ("dlnorm(0.2, 2.0) T(0.5, 4.0)", S.lognorm(s=1 / np.sqrt(2.0), scale=np.exp(0.2)), (0.5, 4.0)),
    "dlogis": ("dlogis(0.5, 1.5) T(-1.0, )", S.logistic(0.5, 1 / np.sqrt(1.5)), (-1.0, np.inf)),
    "ddexp": ("ddexp(0.0, 2.0) T(-0.5, 1.5)", S.laplace(0.0, 1 / np.sqrt(2.0)), (-0.5, 1.5)),
    "dweib": ("dweib(1.7, 0.6) T(0.3, )", S.weibull_min(1.7, scale=1 / 0.6), (0.3, np.inf)),
    "dchisqr": ("dchisqr(3) T(, 6.0)", S.chi2(3), (0.0, 6.0)),
    "dunif": ("dunif(-2, 5) T(0, 3)", S.uniform(-2, 7), (0.0, 3.0)),
    # (regularised incomplete beta function: jbugs_beta_ij)
    "dbeta": ("dbeta(2.5, 4.0) T(0.2, 0.9)", S.beta(2.5, 4.0), (0.2, 0.9)),
    "dbeta_lower": ("dbeta(0.6, 0.8) T(0.7, )", S.beta(0.6, 0.8), (0.7, 1.0)),
    "dbeta_far_tail": ("dbeta(30, 2) T(, 0.3)", S.beta(30, 2), (0.0, 0.3)),
    "dt": ("dt(0.5, 2.0, 4) T(-1.0, 3.0)", S.t(4, loc=0.5, scale=1 / np.sqrt(2.0)), (-1.0, 3.0)),
    "dt_half": ("dt(0, 1.0, 3) T(0, )", S.t(3), (0.0, np.inf)),                       # the half-t prior of scale parameters
    "dt_far_tail": ("dt(0, 1.0, 2.5) T(, -40.0)", S.t(2.5), (-np.inf, -40.0)),
    "dnorm_far_tail": ("dnorm(0, 1) T(6.0, )", S.norm(0, 1), (6.0, np.inf)),
}


def _single(prior):
    return f"model {{ a ~ {prior}\n for (i in 1:N) {{ y[i] ~ dnorm(a, 2.0) }} }}"


def _value_and_logjac(theta, lb, ub):
    if np.isfinite(lb) and np.isfinite(ub):
        sg = 1.0 / (1.0 + np.exp(-theta))
        x = lb + (ub - lb) * sg
        return x, np.log((x - lb) * (ub - x) / (ub - lb))
    if np.isfinite(lb):
        return lb + np.exp(theta), theta
    return ub - np.exp(theta), theta


@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_truncated_families_against_scipy(name):
    prior, dist, (lb, ub) = FAMILIES[name]
    data = {"N": 3, "y": np.array([0.8, 1.4, 1.1]) + (6.0 if name == "dnorm_far_tail" else 0.0)}
    text = _single(prior)
    gm = go.compile(text, data, {}).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    g = plan.globals[0]
    assert (g.lb, g.ub) == (lb, ub)
    theta = np.array([[-0.7, 0.1, 0.9]])
    lp, gr, st = COracle(plan).eval_batch(theta)
    mass = (dist.sf(lb) - dist.sf(ub)) if dist.cdf(lb) > 0.5 else (dist.cdf(ub) - dist.cdf(lb))
    for c in range(3):
        x, lj = _value_and_logjac(theta[0, c], lb, ub)
        expect = dist.logpdf(x) - np.log(mass) + lj + S.norm(x, 1 / np.sqrt(2.0)).logpdf(data["y"]).sum()
        l0, g0 = gm.logdensity_and_gradient(theta[:, c])
        assert l0 == pytest.approx(expect, abs=1e-10)
        assert st[c] == 0 and lp[c] == pytest.approx(l0, abs=1e-10)
        assert grad_err(gr[:, c], g0) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_truncated_families_cuda(name):
    prior, dist, (lb, ub) = FAMILIES[name]
    data = {"N": 3, "y": np.array([0.8, 1.4, 1.1]) + (6.0 if name == "dnorm_far_tail" else 0.0)}
    m = jbugs.compile(_single(prior), data)
    theta = np.linspace(-1.5, 1.5, 64)[None, :]
    lp, g, st = m.cuda.eval_host(theta, layout="param_fast")
    lp0, g0, st0 = COracle(m.plan).eval_batch(theta)
    assert (st == 0).all() and (st0 == 0).all()
    assert np.max(np.abs(lp - lp0) / np.maximum(1.0, np.abs(lp0))) < 1e-12
    assert np.max(np.abs(g - g0) / np.maximum(np.abs(g0), 1e-4 * np.abs(g0).max())) < 1e-10


@pytest.mark.parametrize("seed", range(24))
def test_random_beta_and_t_truncations_against_scipy(seed):
    """the regularised incomplete beta function of include/jbugs_trunc.h over random shapes, scales and bounds (both
    entry sides of the continued fraction, shapes below and above 1, bounds in either tail): C port == node oracle ==
    scipy's cdf / sf, value and gradient"""
    rng = np.random.default_rng(1000 + seed)
    if seed % 2 == 0:
        a, b = np.round(rng.uniform(0.3, 12.0, 2), 3)
        dist = S.beta(a, b)
        q = np.sort(rng.uniform(0.02, 0.98, 2))
        lo, hi = np.round(dist.ppf(q), 4)
        name = f"dbeta({a}, {b})"
        nat = (0.0, 1.0)
    else:
        mu, tau, nu = np.round(rng.uniform(-2, 2), 3), np.round(rng.uniform(0.2, 5.0), 3), np.round(rng.uniform(0.8, 30.0), 2)
        dist = S.t(nu, loc=mu, scale=1 / np.sqrt(tau))
        q = np.sort(rng.uniform(0.001, 0.999, 2))
        lo, hi = np.round(dist.ppf(q), 3)
        name = f"dt({mu}, {tau}, {nu})"
        nat = (-np.inf, np.inf)
    if hi <= lo:
        hi = lo + 0.05
    which = seed % 3    # two-sided, lower only, upper only
    prior, lb, ub = [(f"{name} T({lo}, {hi})", lo, hi), (f"{name} T({lo}, )", lo, nat[1]), (f"{name} T(, {hi})", nat[0], hi)][which]
    data = {"N": 3, "y": np.array([0.3, -0.2, 0.6]) + 0.5 * (lo + hi)}
    text = _single(prior)
    gm = go.compile(text, data, {}).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    g = plan.globals[0]
    assert (g.lb, g.ub) == (lb, ub)
    theta = np.array([[-1.1, 0.2, 1.3]])
    lp, gr, st = COracle(plan).eval_batch(theta)
    mass = (dist.sf(lb) - dist.sf(ub)) if dist.cdf(lb) > 0.5 else (dist.cdf(ub) - dist.cdf(lb))
    for c in range(3):
        x, lj = _value_and_logjac(theta[0, c], lb, ub)
        expect = dist.logpdf(x) - np.log(mass) + lj + S.norm(x, 1 / np.sqrt(2.0)).logpdf(data["y"]).sum()
        l0, g0 = gm.logdensity_and_gradient(theta[:, c])
        assert l0 == pytest.approx(expect, abs=1e-9, rel=1e-11)
        assert st[c] == 0 and lp[c] == pytest.approx(l0, abs=1e-9, rel=1e-11)
        assert grad_err(gr[:, c], g0) < 1e-10
