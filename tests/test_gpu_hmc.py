"""On-device batched HMC.  Statistical checks in the style of the reference's sampler tests
(J/test/ext/JuliaBUGSAdvancedHMCExt.jl:30-63: posterior mean and std vs the WinBUGS `reference_results` at rtol 0.3)."""
import numpy as np
import pytest

import jbugs
from helpers import synth_seeds
from jbugs import hmc

pytestmark = pytest.mark.gpu


def test_conjugate_normal_mean():
    """y_i ~ N(mu, 1), flat-ish prior: posterior N(ybar, 1/n) -- exact answer."""
    rng = np.random.default_rng(0)
    y = rng.normal(1.7, 1.0, 400)
    text = "model { for (i in 1:N) { y[i] ~ dnorm(mu, 1.0) }\n mu ~ dnorm(0.0, 1.0E-6) }"
    m = jbugs.compile(text, {"N": 400, "y": y}, {"mu": 0.0})
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=8, step_size=0.05), n_samples=300, n_adapts=200, n_chains=512, seed=3)
    mu = ch["mu"]
    assert abs(mu.mean() - y.mean()) < 5e-3
    assert abs(mu.std() - 1.0 / np.sqrt(400)) < 0.1 / np.sqrt(400)
    assert 0.5 < ch.accept_rate <= 1.0


def test_seeds_posterior_matches_winbugs_reference(fixtures):
    """Seeds (src/BUGSExamples/Volume_1/04_Seeds.jl:51-57), the second model of the reference's NUTS test."""
    fx = fixtures["seeds"]
    m = jbugs.compile(jbugs.examples.SEEDS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=32, step_size=0.01, jitter=0.05), n_samples=1000, n_adapts=1500, n_chains=256, seed=7)
    ref = {"alpha0": (-0.5499, 0.1965), "alpha1": (0.08902, 0.3124), "alpha12": (-0.841, 0.4372), "alpha2": (1.356, 0.2772)}
    for k, (mean, std) in ref.items():
        assert abs(ch.mean(k) - mean) < 0.15 * std, k     # the reference accepts rtol = 0.3 on the mean itself
        assert ch.std(k) == pytest.approx(std, rel=0.1), k
    sigma = 1.0 / np.sqrt(ch["tau"])
    assert sigma.mean() == pytest.approx(0.2922, rel=0.1) and sigma.std() == pytest.approx(0.1467, rel=0.15)


def test_rats_posterior_matches_winbugs_reference(fixtures):
    """Rats: alpha0 = alpha.c - xbar * beta.c, beta.c, sigma = 1/sqrt(tau.c) against the WinBUGS reference results the
    reference ships (src/BUGSExamples/Volume_1/01_Rats.jl:97-101), same tolerance as the reference's NUTS test."""
    fx = fixtures["rats"]
    m = jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=32, step_size=0.005, jitter=0.05), n_samples=1000, n_adapts=1500, n_chains=256, seed=1234)
    alpha0 = ch["alpha.c"] - 22.0 * ch["beta.c"]
    sigma = 1.0 / np.sqrt(ch["tau.c"])
    ref = {"alpha0": (106.6, 3.66), "beta.c": (6.186, 0.1086), "sigma": (6.093, 0.4643)}
    got = {"alpha0": alpha0, "beta.c": ch["beta.c"], "sigma": sigma}
    for k, (mean, std) in ref.items():
        # the reference's own NUTS test accepts rtol = 0.3; 256 chains x 1000 draws land within a few per cent
        assert got[k].mean() == pytest.approx(mean, rel=0.02), k
        assert got[k].std() == pytest.approx(std, rel=0.1), k
    # between-chain spread of the per-chain means: all chains found the same posterior
    assert alpha0.mean(axis=0).std() < 0.6
    assert 0.6 < ch.accept_rate < 0.95


def test_stacks_posterior_matches_winbugs_reference(fixtures):
    """Stacks, the third model of the reference's NUTS test (src/BUGSExamples/Volume_1/10_Stacks.jl:107-109):
    b0 and outlier[21] are generated quantities, filled in for every draw by jbugs.gq."""
    fx = fixtures["stacks"]
    m = jbugs.compile(jbugs.examples.STACKS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=16, step_size=0.05, jitter=0.05), n_samples=1000, n_adapts=1000, n_chains=256, seed=11)
    b0, outlier = ch["b0"], ch["outlier[21]"]
    assert b0.shape == (1000, 256)
    assert b0.mean() == pytest.approx(-39.64, rel=0.05) and b0.std() == pytest.approx(12.63, rel=0.1)
    assert outlier.mean() == pytest.approx(0.3324, rel=0.15) and outlier.std() == pytest.approx(0.4711, rel=0.1)
    # the same quantity recomputed by hand from the monitored parameters
    x = np.asarray(fx["data"]["x"], dtype=float)
    sd, mean = x.std(axis=0, ddof=1), x.mean(axis=0)
    by_hand = ch["beta0"] - sum(ch[f"beta[{j + 1}]"] / sd[j] * mean[j] for j in range(3))
    assert np.allclose(b0, by_hand, rtol=1e-12)


def test_monitoring_plate_level_parameters(fixtures):
    """labels of plate-level parameters are monitored through the theta map and the local's bijector."""
    fx = fixtures["pumps"]
    m = jbugs.compile(jbugs.examples.PUMPS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=16, step_size=0.05), n_samples=500, n_adapts=500, n_chains=128, seed=5,
                    monitor=["alpha", "beta", "theta[1]", "theta[10]"])
    # posterior means of the classic BUGS manual for Pumps (the reference ships no reference_results for it):
    # theta[1] 0.0598 +- 0.0253, theta[10] 1.99 +- 0.43, alpha 0.70 +- 0.27, beta 0.93 +- 0.54
    assert ch["theta[1]"].min() > 0.0
    assert ch.mean("theta[1]") == pytest.approx(0.0598, rel=0.1)
    assert ch.mean("theta[10]") == pytest.approx(1.99, rel=0.1)
    assert ch.mean("alpha") == pytest.approx(0.70, rel=0.15) and ch.mean("beta") == pytest.approx(0.93, rel=0.2)


def test_nuts_conjugate_normal():
    """NUTS on an exactly known posterior: y_i ~ N(mu, 1) -> N(ybar, 1/n)."""
    rng = np.random.default_rng(0)
    y = rng.normal(1.7, 1.0, 400)
    text = "model { for (i in 1:N) { y[i] ~ dnorm(mu, 1.0) }\n mu ~ dnorm(0.0, 1.0E-6) }"
    m = jbugs.compile(text, {"N": 400, "y": y}, {"mu": 0.0})
    ch = hmc.sample(m, hmc.NUTS(max_depth=8, step_size=0.05), n_samples=400, n_adapts=300, n_chains=512, seed=3)
    mu = ch["mu"]
    assert abs(mu.mean() - y.mean()) < 5e-3
    assert abs(mu.std() - 1.0 / np.sqrt(400)) < 0.08 / np.sqrt(400)
    assert 0.6 < ch.accept_rate <= 1.0 and ch.n_divergent == 0 and 1.0 <= ch.tree_depth <= 4.0


def test_nuts_rats_and_seeds_match_winbugs_reference(fixtures):
    """the reference's own sampler test (NUTS(0.8), 1000 + 1000, rtol 0.3 on mean and std) for Rats and Seeds."""
    fx = fixtures["rats"]
    m = jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.NUTS(max_depth=8, step_size=0.005, jitter=0.05), n_samples=500, n_adapts=1000, n_chains=128, seed=21)
    alpha0 = ch["alpha0"]        # generated quantity: alpha.c - xbar * beta.c
    sigma = ch["sigma"]
    ref = {"alpha0": (106.6, 3.66), "beta.c": (6.186, 0.1086), "sigma": (6.093, 0.4643)}
    got = {"alpha0": alpha0, "beta.c": ch["beta.c"], "sigma": sigma}
    for k, (mean, std) in ref.items():
        assert got[k].mean() == pytest.approx(mean, rel=0.02), k
        assert got[k].std() == pytest.approx(std, rel=0.12), k
    assert 0.6 < ch.accept_rate < 0.95
    diag = ch.summary(["beta.c", "alpha.c", "tau.c"])
    assert all(d["rhat"] < 1.05 and d["ess"] > 2000 for d in diag.values()), diag   # 128 chains x 500 draws
    fx = fixtures["seeds"]
    m = jbugs.compile(jbugs.examples.SEEDS, fx["data"], fx["inits"])
    ch = hmc.sample(m, hmc.NUTS(max_depth=8, step_size=0.01, jitter=0.05), n_samples=500, n_adapts=1000, n_chains=128, seed=22)
    ref = {"alpha0": (-0.5499, 0.1965), "alpha1": (0.08902, 0.3124), "alpha12": (-0.841, 0.4372), "alpha2": (1.356, 0.2772)}
    for k, (mean, std) in ref.items():
        assert abs(ch.mean(k) - mean) < 0.2 * std, k
        assert ch.std(k) == pytest.approx(std, rel=0.12), k


def test_chains_starting_outside_the_support_are_redrawn():
    """alpha is a real-line parameter used as a Poisson rate: some of the jittered starts are negative (DomainError,
    log density -Inf).  The sampler redraws those chains closer to the initial value instead of leaving them frozen at
    a point where every proposal is compared with -Inf."""
    y = np.array([3.0, 5.0, 4.0, 6.0, 2.0])
    m = jbugs.compile("model { for (i in 1:N) { y[i] ~ dpois(alpha) }\n alpha ~ dnorm(4.0, 0.01) }", {"N": 5, "y": y}, {"alpha": 0.5})
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=8, step_size=0.05, jitter=3.0), n_samples=300, n_adapts=300, n_chains=256, seed=2)
    a = ch["alpha"]
    assert np.all(np.isfinite(a)) and np.all(a > 0.0)          # ~43 % of the starts were negative
    assert a.std(axis=0).min() > 0.3                            # every chain moves: a start right at the boundary gets
    #                                                              its own small step size from the per-chain adaptation
    # the likelihood dominates the vague prior: posterior close to Gamma(sum y + 1, n)
    assert a.mean() == pytest.approx((y.sum() + 1) / 5.0, rel=0.05)


@pytest.mark.parametrize("link,ref", [
    # WinBUGS reference results shipped with the example (J/src/BUGSExamples/Volume_2/13_Beetles.jl:42-76)
    ("logit", {"alpha": (-60.78, 5.168), "beta": (34.3, 2.904), "rhat[1]": (3.571, 0.957), "rhat[4]": (33.9, 1.768), "rhat[8]": (58.68, 0.4284)}),
    ("probit", {"alpha": (-35.08, 2.64), "beta": (19.81, 1.485), "rhat[1]": (3.429, 1.015), "rhat[6]": (53.28, 1.153)}),
    ("cloglog", {"alpha": (-39.7, 3.197), "beta": (22.11, 1.775), "rhat[1]": (5.646, 1.113), "rhat[5]": (47.74, 1.712)}),
])
def test_beetles_three_links_match_winbugs_reference(link, ref):
    """Beetles with the logit, probit and complementary log-log links: NUTS posterior of the parameters and of the
    generated quantities alpha and rhat[i] = n[i] p[i] against the WinBUGS values, well inside the reference's rtol 0.3."""
    m = jbugs.compile(jbugs.examples.BEETLES % link, jbugs.examples.BEETLES_DATA, {"alpha.star": 0.0, "beta": 0.0})
    ch = hmc.sample(m, hmc.NUTS(max_depth=8, step_size=0.05, jitter=0.1), n_samples=500, n_adapts=700, n_chains=128, seed=31)
    for k, (mean, std) in ref.items():
        assert ch[k].mean() == pytest.approx(mean, rel=0.03), (link, k)
        assert ch[k].std() == pytest.approx(std, rel=0.12), (link, k)
    assert ch.summary(["beta"])["beta"]["rhat"] < 1.05


# ---- advisor findings of round 1 -------------------------------------------------------------------
def _hmc_handle(model, C, seed=7):
    import ctypes

    from jbugs.cuda import _check, load
    from jbugs.hmc import _bind
    L = load()
    _bind(L)
    h = ctypes.c_void_p()
    _check(L.jbugs_hmc_create(model.cuda.h, C, seed, ctypes.byref(h)))
    return L, h


@pytest.mark.gpu
@pytest.mark.parametrize("N,C", [(4995, 256), (995, 1024), (2995, 512)])  # D = 5000 / 1000 / 3000: odd row tiles before the fix
def test_momentum_draws_are_independent_across_row_tiles(N, C):
    """hmc_refresh_kernel takes the two normals of one Philox call for rows (d, d+1) counted from the tile start; with
    an odd number of rows per tile the last row of a tile and the first of the next shared a counter (identical
    draws).  Every pair of adjacent rows must be uncorrelated and no two rows may be equal."""
    import ctypes

    from jbugs.cuda import _check
    data, b = synth_seeds(N)
    m = jbugs.compile(jbugs.examples.SEEDS, data)
    D = m.plan.D
    L, h = _hmc_handle(m, C)
    try:
        th0 = np.zeros(D)
        th0[0] = np.log(1 / 0.09)
        _check(L.jbugs_hmc_init(h, th0.ctypes.data, 0.01, 0.01))
        p = np.empty((D, C))
        _check(L.jbugs_hmc_draw_momentum(h, 3, p.ctypes.data))
    finally:
        L.jbugs_hmc_destroy(h)
    assert np.all(np.isfinite(p))
    same = np.all(p[1:] == p[:-1], axis=1)
    assert not same.any(), f"rows {np.nonzero(same)[0][:5]} repeat their predecessor"
    z = (p - p.mean(axis=1, keepdims=True)) / p.std(axis=1, keepdims=True)
    corr = (z[1:] * z[:-1]).mean(axis=1)
    assert np.abs(corr).max() < 6.0 / np.sqrt(C), np.abs(corr).max()
    assert abs(p.var() - 1.0) < 0.01  # unit metric at initialisation


@pytest.mark.gpu
def test_monitor_buffers_grow_independently(fixtures):
    """a second run on the same handle that monitors MORE rows for FEWER iterations must not overflow the row-index buffer."""
    import ctypes

    from jbugs.cuda import _check
    fx = fixtures["rats"]
    m = jbugs.compile(jbugs.examples.RATS, fx["data"], fx["inits"])
    C = 64
    L, h = _hmc_handle(m, C)
    try:
        th0 = np.ascontiguousarray(jbugs.getparams(m))
        _check(L.jbugs_hmc_init(h, th0.ctypes.data, 0.01, 0.01))
        acc = ctypes.c_double()
        rows = np.array([0, 1], dtype=np.int64)
        out = np.empty((200, 2, C))
        _check(L.jbugs_hmc_run(h, 200, 4, 0, rows.ctypes.data, 2, out.ctypes.data, ctypes.byref(acc)))
        rows2 = np.arange(60, dtype=np.int64)
        out2 = np.full((5, 60, C), np.nan)
        _check(L.jbugs_hmc_run(h, 5, 4, 0, rows2.ctypes.data, 60, out2.ctypes.data, ctypes.byref(acc)))
        th = np.empty((m.plan.D, C))
        _check(L.jbugs_hmc_get_state(h, th.ctypes.data, None, None, None))
    finally:
        L.jbugs_hmc_destroy(h)
    assert np.all(np.isfinite(out2))
    np.testing.assert_array_equal(out2[-1], th[:60])  # the last monitored draw is the final state


@pytest.mark.gpu
def test_init_refuses_chains_that_stay_outside_the_support():
    """a chain whose log density is not finite can never accept: jbugs_hmc_init must not return OK with frozen chains."""
    from jbugs.cuda import JBugsCudaError, _check
    # x ~ dunif(0, 1) in the constrained space (settrans false): theta0 = 5 is outside the support for every re-draw
    m = jbugs.compile("model { x ~ dunif(0, 1)\n y ~ dnorm(x, 1) }", {"y": 0.3})
    m = jbugs.settrans(m, False)
    L, h = _hmc_handle(m, 32)
    try:
        th0 = np.array([5.0])
        with pytest.raises(JBugsCudaError, match="outside the support"):
            _check(L.jbugs_hmc_init(h, th0.ctypes.data, 0.01, 0.01))
    finally:
        L.jbugs_hmc_destroy(h)
