"""The plate lowering (product) against the per-node oracle: theta layout bit-exact, logp/gradient of the lowered
plan (evaluated by the C restatement) equal to the node-loop oracle."""
import numpy as np
import pytest

import graph_oracle as go
import jbugs
from c_oracle import COracle
from conftest import grad_err, rel_err
from helpers import ZOO, random_thetas, start_theta, synth_rats
from jbugs import plan as P


@pytest.mark.parametrize("name,text,fx,ini", ZOO, ids=[z[0] for z in ZOO])
def test_plan_matches_graph_oracle(fixtures, name, text, fx, ini):
    data, inits = fixtures[fx]["data"], fixtures[fx][ini]
    plan, lw = jbugs.lower(text, data)
    gm = go.compile(text, data, inits).use_generated_order()
    # plate/index lowering must be bit-exact: same dimension, same slot for every parameter
    assert plan.D == gm.D
    assert lw.param_names() == gm.param_names()
    co = COracle(plan)
    th0 = start_theta(name, plan.D, gm)
    theta = random_thetas(th0, 3, seed=5)
    lp, g, st = co.eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g)
        assert grad_err(g[:, c], g_g) < 1e-10


def test_graph_order_via_theta_order(fixtures):
    """UseGraph's reversed-interleaved layout fed as an explicit order: negative strides, same numbers."""
    fx = fixtures["rats"]
    gm = go.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    plan, lw = jbugs.lower(jbugs.examples.RATS, fx["data"], theta_order=gm.param_names())
    loc = {l.name: l for l in plan.plates[0].locals}
    assert (loc["alpha"].theta_base, loc["alpha"].stride_i) == (5 + 2 * 29 + 1, -2)
    assert (loc["beta"].theta_base, loc["beta"].stride_i) == (5 + 2 * 29, -2)
    th = gm.getparams() + 0.01 * np.arange(gm.D) / gm.D
    lp, g, _ = COracle(plan).eval_batch(th.reshape(-1, 1))
    lp_g, g_g = gm.logdensity_and_gradient(th)
    assert abs(lp[0] - lp_g) <= 1e-12 * abs(lp_g)
    assert grad_err(g[:, 0], g_g) < 1e-10


def test_epil_graph_order_index_maps(fixtures):
    fx = fixtures["epil"]
    gm = go.compile(jbugs.examples.EPIL, fx["data"], fx["inits"])
    plan, _ = jbugs.lower(jbugs.examples.EPIL, fx["data"], theta_order=gm.param_names())
    loc = {l.name: l for l in plan.plates[0].locals}
    # for j = N..1: b1[j], b[j,T], ..., b[j,1]
    assert (loc["b1"].theta_base, loc["b1"].stride_i) == (8 + 58 * 5, -5)
    assert (loc["b"].theta_base, loc["b"].stride_i, loc["b"].stride_j) == (8 + 58 * 5 + 4, -5, -1)


def test_untransformed_model(fixtures):
    """settrans(model, false): theta in constrained space, no log-Jacobian (bugsmodel.jl:697-707)."""
    fx = fixtures["rats"]
    plan, _ = jbugs.lower(jbugs.examples.RATS, fx["data"], transformed=False)
    assert all(g.transform == P.TR_IDENTITY for g in plan.globals)
    gm = go.compile(jbugs.examples.RATS, fx["data"], fx["inits"], transformed=False).use_generated_order()
    th = gm.getparams()
    lp, g, _ = COracle(plan).eval_batch(th.reshape(-1, 1))
    assert abs(lp[0] - fx["golden_logjoint"]["untransformed"]) < 1e-10 * 174029.4


def test_gtape_scalar_logicals():
    """BUGS idiom `tau <- 1 / (sigma * sigma)` with a uniform prior on sigma: scalar logical nodes of globals."""
    text = """
    model {
        for (i in 1:N) { y[i] ~ dnorm(mu + b * x[i], tau) }
        tau <- 1 / (sigma * sigma)
        sigma ~ dunif(0, 10)
        mu ~ dnorm(0, 0.001)
        b ~ dnorm(0, 0.001)
    }"""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(40)
    data = {"N": 40, "x": x, "y": 1.0 + 2.0 * x + 0.3 * rng.standard_normal(40)}
    plan, lw = jbugs.lower(text, data)
    assert len(plan.gtape) >= 2 and plan.globals[0].transform in (P.TR_LOGIT, P.TR_IDENTITY)
    gm = go.compile(text, data).use_generated_order()
    assert lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 4, seed=2, scale=0.5)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(4):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g)
        assert grad_err(g[:, c], g_g) < 1e-10


SIBLINGS = """
model {
    for (i in 1:N) {
        y1[i] ~ dnorm(m1[i], tau)
        y2[i] ~ dnorm(m2[i], tau)
        y3[i] ~ dnorm(b[i], tau)
        m1[i] <- a + b[i] * x1[i] + c[i]
        m2[i] <- a * 2 + b[i] * x2[i] - off[i]
        b[i] ~ dnorm(0, 1)
        c[i] ~ dnorm(mu.c, 4)
    }
    a ~ dnorm(0, 0.01)
    mu.c ~ dnorm(0, 0.01)
    tau ~ dgamma(2, 2)
}"""


def siblings_data(N, seed=4):
    rng = np.random.default_rng(seed)
    return {"N": N, "x1": rng.standard_normal(N), "x2": rng.standard_normal(N), "off": rng.standard_normal(N),
            "y1": rng.standard_normal(N), "y2": rng.standard_normal(N), "y3": rng.standard_normal(N)}


def test_sibling_observations_merge_into_one_plate():
    """three observation statements of one loop sharing b[i] (the Blockers pattern, with per-statement covariates
    and offsets): one N x 3 plate whose columns are the statements, same numbers as the node loop."""
    data = siblings_data(17)
    plan, lw = jbugs.lower(SIBLINGS, data)
    assert len(plan.plates) == 1 and (plan.plates[0].N, plan.plates[0].T) == (17, 3)
    assert sorted(l.name for l in plan.plates[0].locals) == ["b", "c"]
    gm = go.compile(SIBLINGS, data).use_generated_order()
    assert lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 4, seed=2, scale=0.5)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(4):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g)
        assert grad_err(g[:, c], g_g) < 1e-10


def test_large_plate_is_lowered_at_statement_level():
    """10^5 rats lower in well under a second: no per-node unrolling (the reference's compile would create
    1.2e6 vertices here)."""
    import time
    data, _, _ = synth_rats(100_000)
    t = time.time()
    plan, _ = jbugs.lower(jbugs.examples.RATS, data)
    assert time.time() - t < 5.0
    assert plan.D == 200_005 and plan.plates[0].N == 100_000 and plan.n_obs() == 500_000
    a = plan.plates[0].locals[0]
    assert (a.name, a.theta_base, a.stride_i) == ("alpha", 6, 2)


def test_tiled_data_property():
    """size-independent property: a plate made of k copies of a block, with theta tiled the same way, has
    log-likelihood contributions that add up (checked through the C oracle at a size the node oracle would
    take minutes for)."""
    data1, a1, b1 = synth_rats(50, seed=3)
    k = 40
    dataK = dict(data1, N=50 * k, Y=np.tile(data1["Y"], (k, 1)))
    p1, _ = jbugs.lower(jbugs.examples.RATS, data1)
    pK, _ = jbugs.lower(jbugs.examples.RATS, dataK)
    hyp = np.array([np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)])
    loc1 = np.empty(100)
    loc1[0::2], loc1[1::2] = b1, a1
    th1 = np.concatenate([hyp, loc1])
    thK = np.concatenate([hyp, np.tile(loc1, k)])
    lp1, g1, _ = COracle(p1).eval_batch(th1.reshape(-1, 1))
    lpK, gK, _ = COracle(pK).eval_batch(thK.reshape(-1, 1))
    # hyper-prior part counted once, plate part k times
    p0, _ = jbugs.lower(jbugs.examples.RATS, dict(data1, N=0, Y=np.zeros((0, 5))))
    lp0, _, _ = COracle(p0).eval_batch(hyp.reshape(-1, 1))
    assert lpK[0] == pytest.approx(lp0[0] + k * (lp1[0] - lp0[0]), rel=1e-12)
    assert np.allclose(gK[5:105, 0], g1[5:, 0], rtol=1e-13, atol=0)


NONLINEAR = [
    "model { for (i in 1:N) { y[i] ~ dnorm(mu[i], 1)\n mu[i] <- a * b * x[i] }\n a ~ dnorm(0,1)\n b ~ dnorm(0,1) }",
    "model { for (i in 1:N) { y[i] ~ dnorm(exp(a) * x[i], 1) }\n a ~ dnorm(0,1) }",
    "model { for (i in 1:N) { y[i] ~ dgamma(exp(a + u[i]), b * x[i]) \n u[i] ~ dnorm(0, t) }\n a ~ dnorm(0,1)\n b ~ dgamma(1,1)\n t ~ dgamma(2,2) }",
]


@pytest.mark.parametrize("text", NONLINEAR)
def test_nonlinear_predictors_become_expression_plates(text):
    """round 1 refused these ("not linear in the parameters"); they now lower to an expression tape and agree with the
    per-node oracle (dual numbers through the node loop) at 1e-10"""
    rng = np.random.default_rng(4)
    data = {"N": 7, "x": 0.5 + rng.random(7), "y": 0.5 + rng.random(7)}
    plan, lw = jbugs.lower(text, data)
    assert len(plan.plates) == 1 and plan.plates[0].xops and not plan.plates[0].terms
    gm = go.compile(text, data).use_generated_order()
    assert lw.param_names() == gm.param_names()
    theta = random_thetas(np.zeros(plan.D), 3, seed=1, scale=0.4)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-12 * max(1.0, abs(lp_g))
        assert grad_err(g[:, c], g_g) < 1e-10


@pytest.mark.parametrize("text,msg", [
    ("model { for (i in 1:N) { y[i] ~ dnorm(abs(a) * x[i], 1) }\n a ~ dnorm(0,1) }", "unsupported expression"),
    ("model { for (i in 1:N) { y[i] ~ dmnorm(a, 1) }\n a ~ dnorm(0,1) }", "outside the supported"),
    ("model { for (i in 1:N) { y[i] ~ dnorm(u[i], 1)\n u[i] ~ dnorm(v[i], 1)\n v[i] ~ dnorm(a, 1) }\n a ~ dnorm(0,1) }", "must be a constant, data, or a scalar parameter"),
])
def test_unsupported_models_fail_loudly(text, msg):
    with pytest.raises(jbugs.LoweringError, match=msg):
        jbugs.lower(text, {"N": 3, "x": np.ones(3), "y": np.ones(3)})


def test_partially_observed_plate_carries_a_weight(fixtures):
    """Bones: 20 of the 442 grades are missing.  Nothing depends on them, so they are generated quantities and leave the
    target (the reference's golden constant is reproduced only that way); the plate marks them with weight 0."""
    fx = fixtures["bones"]
    plan, _ = jbugs.lower(jbugs.examples.BONES, fx["data"])
    pl = plan.plates[0]
    assert pl.family == P.FAM_CATEGORICAL and pl.weight is not None and pl.n_real == 422 and plan.n_obs() == 422
    w = plan.data[pl.weight.id].values
    assert w.sum() == 422 and set(np.unique(w)) == {0.0, 1.0}
    th = np.zeros(plan.D)
    th[:] = np.asarray(fx["inits"]["theta"], dtype=float)
    lp, _, _ = COracle(plan).eval_batch(th.reshape(-1, 1))
    assert abs(lp[0] - fx["golden_logjoint"]["transformed"]) <= 1e-12 * 139.0


def test_partially_observed_with_children_is_refused():
    text = "model { for (i in 1:N) { x[i] ~ dnorm(0, 1)\n y[i] ~ dnorm(x[i], 1) } }"
    with pytest.raises(jbugs.LoweringError, match="partially observed"):
        jbugs.lower(text, {"N": 3, "x": np.array([0.1, np.nan, 0.3]), "y": np.ones(3)})


def test_kidney_drops_the_rows_without_observations(fixtures):
    """Kidney: b[14], b[19], b[36] have no observed descendants (both of the patient's times are censored = missing), so
    the reference takes exactly those ELEMENTS out of theta: 43 parameters, not 46.  The plate and the statement of b
    are compacted to the remaining 35 rows."""
    fx = fixtures["kidney"]
    gm = go.compile(jbugs.examples.KIDNEY, fx["data"], fx["inits"]).use_generated_order()
    assert gm.D == 43 and "b[14]" not in gm.param_names()
    plan, lw = jbugs.lower(jbugs.examples.KIDNEY, fx["data"])
    assert plan.D == 43 and lw.param_names() == gm.param_names()
    assert (plan.plates[0].N, plan.plates[0].T, plan.plates[0].n_real) == (35, 2, 58)
    # a row's parameter that also feeds something else cannot be dropped
    text = jbugs.examples.KIDNEY.replace("sigma <- 1 / sqrt(tau)", "sigma <- 1 / sqrt(tau)\n    for (i in 1:N) { z[i] ~ dnorm(b[i], 1) }")
    data = dict(fx["data"], z=np.zeros(38))
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(text, data)


def test_factor_levels_picked_by_a_data_index():
    """the Kidney idiom on complete data: beta.dis[disease[i]] with a corner-point constraint given as data"""
    rng = np.random.default_rng(4)
    N, M = 30, 2
    text = jbugs.examples.KIDNEY
    dis = rng.integers(1, 5, N).astype(float)
    dis[:4] = [1, 2, 3, 4]
    data = {"N": N, "M": M, "t": rng.weibull(1.2, (N, M)) * 50 + 1.0, "t.cen": np.zeros((N, M)), "age": rng.normal(45, 10, (N, M)),
            "sex": rng.integers(0, 2, N).astype(float), "disease": dis, "beta.dis": np.array([0.0, np.nan, np.nan, np.nan])}
    data["t"][3, 1] = np.nan
    data["t.cen"][3, 1] = 60.0
    gm = go.compile(text, data, {}).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert plan.D == gm.D == 8 + N and lw.param_names() == gm.param_names()
    th0 = np.zeros(plan.D)
    names = lw.param_names()
    th0[names.index("alpha")] = -4.0
    th0[names.index("beta.age")] = 0.002
    theta = random_thetas(th0, 3, seed=2, scale=0.05)
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(3):
        lp_g, g_g = gm.logdensity_and_gradient(theta[:, c])
        assert st[c] == 0 and abs(lp[c] - lp_g) <= 1e-12 * abs(lp_g)
        assert grad_err(g[:, c], g_g) < 1e-10
    # an observation ON its censoring bound is refused (it would need the inner distribution's cdf)
    data["t.cen"][0, 0] = data["t"][0, 0]
    with pytest.raises(jbugs.LoweringError, match="censoring bounds"):
        jbugs.lower(text, data)


# ---- third reading: short outer loops written out, sibling statements with non-linear predictors merged ---------------
OUTER_LINEAR = """model {
  prec[1] ~ dgamma(2, 1)
  s[2] ~ dunif(0.5, 3)
  prec[2] <- 1 / (s[2] * s[2])
  prec[3] <- 0.7
  for (j in 1:3) {
    m[j] ~ dnorm(0, prec[j])
    for (k in 1:K) {
      th[j, k] ~ dnorm(m[j], 2.0)
      y[j, k] ~ dpois(lam[j, k])
      log(lam[j, k]) <- th[j, k] + b * x[k]
    }
  }
  b ~ dnorm(0, 1)
}"""

OUTER_SIBLINGS = """model {
  for (j in 1:2) {
    mu[j] ~ dnorm(0, 0.5)
    for (k in 1:K) {
      d[j, k] ~ dnorm(mu[j], 1.5)
      p0[j, k] ~ dbeta(2, 3)
      r0[j, k] ~ dbin(p0[j, k], n[k])
      r1[j, k] ~ dbin(p1[j, k], n[k])
      logit(p1[j, k]) <- d[j, k] + logit(p0[j, k])
    }
  }
}"""


@pytest.mark.parametrize("which", ["linear", "siblings"])
def test_short_outer_loops_are_written_out(which):
    rng = np.random.default_rng(17)
    K = 5
    if which == "linear":
        text = OUTER_LINEAR
        data = {"K": K, "x": rng.standard_normal(K), "y": rng.poisson(2.0, (3, K)).astype(float)}
    else:
        text = OUTER_SIBLINGS
        n = rng.integers(5, 20, K).astype(float)
        data = {"K": K, "n": n, "r0": rng.binomial(n.astype(int), 0.3, (2, K)).astype(float),
                "r1": rng.binomial(n.astype(int), 0.5, (2, K)).astype(float)}
    gm = go.compile(text, data, {}).use_generated_order()
    plan, lw = jbugs.lower(text, data)
    assert lw.param_names() == gm.param_names()          # theta keeps the order of the program as written
    assert not any("<" in n for n in lw.param_names())   # public labels carry the dropped subscript again
    theta = 0.3 * rng.standard_normal((plan.D, 4))
    lp, g, st = COracle(plan).eval_batch(theta)
    for c in range(4):
        l0, g0 = gm.logdensity_and_gradient(theta[:, c])
        assert st[c] == 0 and abs(lp[c] - l0) <= 1e-11 * max(1.0, abs(l0))
        assert grad_err(g[:, c], g0) < 1e-10
    if which == "siblings":
        assert len(plan.plates) == 2 and all(p.T == 2 and p.xops for p in plan.plates)   # one merged plate per iteration


def test_arrays_used_outside_their_outer_loop_are_not_split():
    text = """model {
      for (j in 1:2) { m[j] ~ dnorm(0, 1)
        for (k in 1:K) { th[j, k] ~ dnorm(m[j], 1)
                         y[j, k] ~ dnorm(th[j, k], 1) } }
      z ~ dnorm(th[1, 2], 1)
    }"""
    with pytest.raises(jbugs.LoweringError):
        jbugs.lower(text, {"K": 3, "y": np.zeros((2, 3)), "z": 0.5})


def test_unrolled_model_keeps_its_public_names(fixtures):
    """Magnesium is lowered with its outer loop written out (arrays split per iteration inside the plan); what a user sees
    -- parameter names, the sampler's monitor table, generated quantities -- carries the names of the model text"""
    import jbugs.hmc as hmc
    from jbugs import examples as ex, gq
    data = fixtures["magnesium"]["data"]
    plan, lw = jbugs.lower(ex.MAGNESIUM, data)
    pn = lw.param_names()
    assert "theta[3,5]" in pn and not any("<" in n for n in pn)

    class M:   # (the table needs the plan and the lowering only: no device)
        pass
    m = M()
    m.plan, m.lowering = plan, lw
    names, by = hmc.monitor_table(m)
    assert {"mu[1]", "mu[6]", "B0", "D0", "tau[3]"} <= set(names) and not any("<" in n for n in names)
    names, by = hmc.monitor_table(m, ["mu[3]", "theta[3,5]", "pc[6,8]"])
    assert [pn[by[n][0]] for n in names] == ["mu[3]", "theta[3,5]", "pc[6,8]"]
    assert by["pc[6,8]"][1] == P.TR_LOGIT and by["theta[3,5]"][1] == P.TR_IDENTITY
    with pytest.raises(KeyError):
        hmc.monitor_table(m, ["theta[7,1]"])
    rng = np.random.default_rng(0)
    S = 5
    draws = {n: (rng.uniform(0.1, 0.9, S) if n.startswith(("pc", "B0", "D0")) else
                 rng.uniform(0.5, 2.0, S) if n.startswith(("tau", "inv.tau")) else rng.standard_normal(S)) for n in pn}
    out = gq.generated_quantities(lw, draws)
    assert out["odds.ratio"].shape == (6, S)
    assert np.allclose(out["odds.ratio"][2], np.exp(draws["mu[3]"]))


def test_monitor_table_of_the_sampler_test_models(fixtures):
    """the sampler's monitor table (names -> theta slot, bijector, bounds) for the models tests/test_gpu_hmc.py samples,
    written out independently: globals under their plan names in plan order; elements of plate-level parameters from the
    theta map and the plate's local record.  (No device: the table needs the plan and the lowering only.)"""
    import jbugs.hmc as hmc
    from jbugs import examples as ex

    def expected(m, monitor):
        by = {g.name: (g.theta_idx, g.transform, g.lb, g.ub) for g in m.plan.globals}
        names = [g.name for g in m.plan.globals] if monitor is None else list(monitor)
        slot = {lab: k for k, lab in enumerate(m.lowering.param_names())}
        tr = {l.name: (l.transform, l.lb, l.ub) for pl in m.plan.plates for l in pl.locals}
        for n in names:
            if n not in by:
                by[n] = (slot[n],) + tr[n.partition("[")[0]]
        return names, {n: by[n] for n in names}

    cases = [(ex.SEEDS, "seeds", None), (ex.RATS, "rats", None), (ex.STACKS, "stacks", None),
             (ex.PUMPS, "pumps", ["alpha", "beta", "theta[1]", "theta[10]"]), (ex.RATS, "rats", ["alpha[17]", "beta[2]", "tau.c"])]
    for text, name, monitor in cases:
        m = jbugs.compile(text, fixtures[name]["data"], fixtures[name]["inits"])
        names, by = hmc.monitor_table(m, monitor)
        names0, by0 = expected(m, monitor)
        assert names == names0 and {n: by[n] for n in names} == by0
    m = jbugs.compile(ex.BEETLES % "probit", ex.BEETLES_DATA, {"alpha.star": 0.0, "beta": 0.0})
    assert hmc.monitor_table(m)[0] == [g.name for g in m.plan.globals]
