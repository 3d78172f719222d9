"""Generated quantities computed for all draws at once (jbugs/gq.py) against the per-node oracle's environment and
against the defining formulas (reference: J/src/model/abstractmcmc.jl:92-110, evaluation.jl:354-378)."""
import numpy as np
import pytest

import graph_oracle as go
import jbugs
from jbugs import examples as ex
from jbugs.gq import generated_quantities


def _draws_from_thetas(gm, thetas):
    """constrained parameter values per label for S points theta (columns), through the oracle's own env."""
    names = gm.param_names()
    vals = {n: [] for n in names}
    envs = []
    for c in range(thetas.shape[1]):
        env, _ = gm.evaluate_with_values(thetas[:, c])
        envs.append(env)
        for n in names:
            base, _, rest = n.partition("[")
            v = env[base]
            if rest:
                v = v[tuple(int(t) - 1 for t in rest.rstrip("]").split(","))]
            vals[n].append(float(v))
    return {n: np.asarray(v) for n, v in vals.items()}, envs


@pytest.mark.parametrize("name,text,nodes", [
    ("rats", ex.RATS, ["mu"]),
    ("seeds", ex.SEEDS, ["p"]),
    ("epil", ex.EPIL, ["mu"]),
    ("stacks", ex.STACKS, ["mu"]),
])
def test_logical_nodes_match_the_node_loop(fixtures, name, text, nodes):
    fx = fixtures[name]
    gm = go.compile(text, fx["data"], fx["inits"]).use_generated_order()
    lw = jbugs.Lowering(text, fx["data"])
    th0 = gm.getparams()
    th0 = np.where(np.isfinite(th0), th0, 0.0)
    rng = np.random.default_rng(0)
    thetas = th0[:, None] + 0.05 * rng.standard_normal((gm.D, 3))
    draws, envs = _draws_from_thetas(gm, thetas)
    out = generated_quantities(lw, draws, only_gq=False)
    for nd in nodes:
        for c, env in enumerate(envs):
            ref = np.asarray(env[nd], dtype=np.float64)
            got = out[nd][..., c]
            mask = ~np.isnan(ref)
            assert np.allclose(got[mask], ref[mask], rtol=1e-13, atol=0), (name, nd)


def test_rats_generated_quantities_follow_their_definitions(fixtures):
    fx = fixtures["rats"]
    lw = jbugs.Lowering(ex.RATS, fx["data"])
    rng = np.random.default_rng(1)
    S = 50
    draws = {"tau.c": rng.gamma(2.0, 0.02, S), "alpha.c": rng.normal(240, 3, S), "beta.c": rng.normal(6, 0.1, S)}
    out = generated_quantities(lw, draws)
    assert set(out) == {"sigma", "alpha0"}          # exactly the nodes the reference excludes from the target
    assert np.allclose(out["sigma"], 1.0 / np.sqrt(draws["tau.c"]), rtol=1e-15)
    assert np.allclose(out["alpha0"], draws["alpha.c"] - fx["data"]["xbar"] * draws["beta.c"], rtol=1e-15)


def test_blockers_forward_samples_the_unobserved_leaf(fixtures):
    """delta.new ~ dnorm(d, tau) has no observed descendant: it is drawn forward, once per posterior draw."""
    fx = fixtures["blockers"]
    lw = jbugs.Lowering(ex.BLOCKERS, fx["data"])
    S = 200_000
    draws = {"d": np.full(S, -0.25), "tau": np.full(S, 16.0)}
    out = generated_quantities(lw, draws, seed=3)
    assert np.allclose(out["sigma"], 0.25)
    assert out["delta.new"].mean() == pytest.approx(-0.25, abs=3e-3)
    assert out["delta.new"].std() == pytest.approx(0.25, rel=1e-2)


def test_stacks_indexed_generated_quantities(fixtures):
    fx = fixtures["stacks"]
    lw = jbugs.Lowering(ex.STACKS, fx["data"])
    rng = np.random.default_rng(2)
    S = 20
    draws = {"beta0": rng.normal(17, 1, S), "tau": rng.gamma(5, 0.02, S),
             "beta[1]": rng.normal(6, 1, S), "beta[2]": rng.normal(3, 1, S), "beta[3]": rng.normal(0, 1, S)}
    out = generated_quantities(lw, draws)
    x = np.asarray(fx["data"]["x"], dtype=float)
    Y = np.asarray(fx["data"]["Y"], dtype=float)
    sd, mean = x.std(axis=0, ddof=1), x.mean(axis=0)
    b = np.stack([draws[f"beta[{j + 1}]"] / sd[j] for j in range(3)])
    assert np.allclose(out["b"], b, rtol=1e-14)
    assert np.allclose(out["b0"], draws["beta0"] - (b * mean[:, None]).sum(axis=0), rtol=1e-13)
    z = (x - mean) / sd
    mu = draws["beta0"][None, :] + z @ np.stack([draws[f"beta[{j + 1}]"] for j in range(3)])
    stres = (Y[:, None] - mu) * np.sqrt(draws["tau"])[None, :]
    assert np.allclose(out["stres"], stres, rtol=1e-12)
    assert np.array_equal(out["outlier"], (stres - 2.5 >= 0).astype(float) + (-(stres + 2.5) >= 0).astype(float))


def test_condition_and_decondition(fixtures):
    """AbstractPPL.condition / decondition (J/src/model/abstractppl.jl:700-799; contract pinned by
    J/test/model/abstractppl.jl): conditioned variables leave theta and are scored as observations; deconditioning
    restores the original plan byte for byte."""
    from c_oracle import COracle
    fx = fixtures["rats"]
    m = jbugs.compile(ex.RATS, fx["data"], fx["inits"])
    mc = jbugs.condition(m, {"tau.c": 0.03, "alpha.c": 240.0})
    assert m.plan.D == 65 and mc.plan.D == 63
    assert "tau.c" not in mc.lowering.param_names() and "alpha.c" not in mc.lowering.param_names()
    gm = go.compile(ex.RATS, dict(fx["data"], **{"tau.c": 0.03, "alpha.c": 240.0}), fx["inits"]).use_generated_order()
    assert mc.lowering.param_names() == gm.param_names()
    th = np.where(np.isfinite(gm.getparams()), gm.getparams(), 0.0) + 0.01
    lp, g, _ = COracle(mc.plan).eval_batch(th.reshape(-1, 1))
    lp_g, g_g = gm.logdensity_and_gradient(th)
    assert abs(lp[0] - lp_g) <= 1e-12 * abs(lp_g)
    assert np.max(np.abs(g[:, 0] - g_g) / np.maximum(np.abs(g_g), 1e-6 * np.abs(g_g).max())) < 1e-10
    md = jbugs.decondition(mc, "tau.c")
    assert md.plan.D == 64 and "tau.c" in md.lowering.param_names()
    assert jbugs.decondition(md).plan.to_bytes() == m.plan.to_bytes()
    with pytest.raises(KeyError):
        jbugs.decondition(m, "tau.c")


def test_chain_diagnostics_on_known_processes():
    """split-R-hat ~ 1 and ESS ~ N (1 - rho) / (1 + rho) for stationary AR(1) chains; R-hat >> 1 when chains disagree."""
    from jbugs.hmc import Chains, effective_sample_size, split_rhat
    rng = np.random.default_rng(0)
    n, m, rho = 4000, 8, 0.6
    x = np.empty((n, m))
    x[0] = rng.standard_normal(m)
    for t in range(1, n):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal(m)
    assert abs(split_rhat(x) - 1.0) < 0.01
    assert effective_sample_size(x) == pytest.approx(n * m * (1 - rho) / (1 + rho), rel=0.12)
    iid = rng.standard_normal((n, m))
    assert effective_sample_size(iid) == pytest.approx(n * m, rel=0.1)
    assert split_rhat(iid + np.arange(m)[None, :]) > 1.5
    ch = Chains(["a"], iid[:, None, :], 0.8, 0.1)
    s = ch.summary()["a"]
    assert abs(s["mean"]) < 0.03 and abs(s["std"] - 1.0) < 0.03 and abs(s["rhat"] - 1.0) < 0.01


@pytest.mark.parametrize("fn,args,dist", [
    ("dlogis", (0.5, 4.0), lambda S: S.logistic(0.5, 0.5)),
    ("ddexp", (0.5, 4.0), lambda S: S.laplace(0.5, 0.5)),
    ("dweib", (1.7, 0.8), lambda S: S.weibull_min(1.7, scale=1.25)),
    ("dchisqr", (5.0,), lambda S: S.chi2(5)),
    ("dt", (1.5, 4.0, 7.0), lambda S: S.t(7, 1.5, 0.5)),
    ("dnegbin", (0.4, 3.0), lambda S: S.nbinom(3, 0.4)),
    ("dgeom", (0.3,), lambda S: S.geom(0.3, loc=-1)),
    ("dbetabin", (2.5, 1.5, 9.0), lambda S: S.betabinom(9, 2.5, 1.5)),
])
def test_forward_sampler_parameterisations(fn, args, dist):
    """The BUGS -> numpy parameterisation of each forward sampler (precision -> scale, rate -> scale, dgeom counting
    failures: J/src/BUGSPrimitives/distributions.jl) against the moments of the scipy distribution the density tests use."""
    from scipy import stats
    from jbugs.gq import _sample
    d = dist(stats)
    x = _sample(np.random.default_rng(7), fn, args, (400_000,))
    assert x.shape == (400_000,)
    assert x.mean() == pytest.approx(d.mean(), abs=6 * d.std() / np.sqrt(x.size))
    assert x.std() == pytest.approx(d.std(), rel=2e-2)
