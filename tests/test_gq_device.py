"""Generated quantities on the device (include/jbugs_gq.h, jbugs/gq_device.py): the lowered program against the
per-node oracle's evaluation environment (reference: J/src/model/abstractmcmc.jl:92-110, evaluation.jl:354-378).
CPU: the program read by a numpy interpreter (tests/gq_interp.py).  GPU: the kernel through the C ABI, and the
forward samplers against the moments of the scipy distributions of the density tests."""
import numpy as np
import pytest

import graph_oracle as go
import jbugs
from gq_interp import run_program
from jbugs import examples as ex
from jbugs.gq import generated_quantities
from jbugs.gq_device import DeviceGQ, GQProgram

CASES = [
    ("rats", ex.RATS, "rats", ["mu"]),
    ("seeds", ex.SEEDS, "seeds", ["p"]),
    ("epil", ex.EPIL, "epil", ["mu"]),
    ("stacks", ex.STACKS, "stacks", ["mu"]),
    ("dogs", ex.DOGS, "dogs", ["p"]),
    ("dugongs", ex.DUGONGS, "dugongs", ["mu"]),
]


def _thetas(gm, S, seed=0):
    th0 = gm.getparams()
    th0 = np.where(np.isfinite(th0), th0, 0.0)
    return th0[:, None] + 0.05 * np.random.default_rng(seed).standard_normal((gm.D, S))


def _check_against_env(out, gm, thetas, nodes, name, rtol=1e-12):
    for c in range(thetas.shape[1]):
        env, _ = gm.evaluate_with_values(thetas[:, c])
        for nd in nodes:
            ref = np.asarray(env[nd], dtype=np.float64)
            got = np.asarray(out[nd])[..., c]
            if ref.shape != got.shape:   # the oracle's arrays may be larger than the statement's loop domain
                ref = ref[tuple(slice(0, n) for n in got.shape)]
            mask = ~np.isnan(ref) & ~np.isnan(got)
            assert mask.any(), (name, nd)
            assert np.allclose(got[mask], ref[mask], rtol=rtol, atol=0), (name, nd)


@pytest.mark.parametrize("name,text,fx,nodes", CASES, ids=[c[0] for c in CASES])
def test_program_matches_the_node_loop(fixtures, name, text, fx, nodes):
    f = fixtures[fx]
    gm = go.compile(text, f["data"], f["inits"]).use_generated_order()
    lw = jbugs.Lowering(text, f["data"])
    prog = GQProgram(lw)
    thetas = _thetas(gm, 3)
    Q = run_program(prog, thetas)
    out = {k: np.where(np.asarray(r)[..., None] >= 0, Q[np.maximum(np.asarray(r), 0)], np.nan) for k, (r, _) in prog.outputs.items()}
    _check_against_env(out, gm, thetas, nodes, name)


def test_program_outputs_are_the_nodes_outside_the_target(fixtures):
    lw = jbugs.Lowering(ex.RATS, fixtures["rats"]["data"])
    prog = GQProgram(lw)
    assert {k for k, (_, gq) in prog.outputs.items() if gq} == {"sigma", "alpha0"}   # what the reference excludes from the target
    assert not prog.skipped
    need = prog.theta_rows_needed()
    names = lw.param_names()
    assert {"tau.c", "alpha.c", "beta.c"} <= {names[r] for r in need}
    # ... and they follow their definitions
    gm = go.compile(ex.RATS, fixtures["rats"]["data"], fixtures["rats"]["inits"]).use_generated_order()
    th = _thetas(gm, 5)
    Q = run_program(prog, th)
    tau_c, alpha_c, beta_c = (th[names.index(n)] for n in ("tau.c", "alpha.c", "beta.c"))
    assert np.allclose(Q[int(prog.outputs["sigma"][0])], 1.0 / np.sqrt(np.exp(tau_c)), rtol=1e-14)
    assert np.allclose(Q[int(prog.outputs["alpha0"][0])], alpha_c - fixtures["rats"]["data"]["xbar"] * beta_c, rtol=1e-14)
    # restricted to what the quantities outside the target need: mu[i, j] (and with it alpha[i], beta[i]) drops out
    small = GQProgram(lw, only_gq=True)
    assert set(small.outputs) == {"sigma", "alpha0"}
    assert {names[r] for r in small.theta_rows_needed()} == {"tau.c", "alpha.c", "beta.c"}
    # the host twin names the same quantities
    S = 4
    rng = np.random.default_rng(1)
    draws = {"tau.c": rng.gamma(2.0, 0.02, S), "alpha.c": rng.normal(240, 3, S), "beta.c": rng.normal(6, 0.1, S)}
    assert set(generated_quantities(lw, draws)) == {"sigma", "alpha0"}


def test_reductions_and_function_vocabulary():
    text = """model {
      for (i in 1:N) { y[i] ~ dnorm(b[g[i]], tau) }
      for (k in 1:K) { b[k] ~ dnorm(m, 1) }
      m ~ dnorm(0, 0.1)
      tau ~ dgamma(1, 1)
      bbar <- mean(b[1:K])
      bsum <- sum(b[1:K])
      bsd <- sd(b[1:K])
      ip <- inprod(b[1:K], w[1:K])
      for (k in 1:K) {
        big[k] <- step(b[k] - bbar) * max(b[k], 0) + equals(k, 2) * abs(b[k]) + min(b[k], m) + pow(b[k], 2) / exp(tau)
      }
      ynew ~ dnorm(bbar, tau)
    }"""
    rng = np.random.default_rng(0)
    K, N = 5, 30
    data = {"N": N, "K": K, "g": rng.integers(1, K + 1, N).astype(float), "y": rng.standard_normal(N), "w": np.arange(1.0, K + 1)}
    lw = jbugs.Lowering(text, data)
    prog = GQProgram(lw)
    assert not prog.skipped, prog.skipped
    names = lw.param_names()
    theta = 0.5 * rng.standard_normal((lw.D, 6))
    Q = run_program(prog, theta)
    row = {n: k for k, n in enumerate(names)}
    b = np.stack([theta[row[f"b[{k}]"]] for k in range(1, K + 1)])
    m, tau = theta[row["m"]], np.exp(theta[row["tau"]])
    get = lambda nm: Q[np.asarray(prog.outputs[nm][0])]  # noqa: E731
    assert np.allclose(get("bbar"), b.mean(0), rtol=1e-14)
    assert np.allclose(get("bsum"), b.sum(0), rtol=1e-14)
    assert np.allclose(get("bsd"), b.std(0, ddof=1), rtol=1e-13)
    assert np.allclose(get("ip"), (b * data["w"][:, None]).sum(0), rtol=1e-14)
    big = ((b - b.mean(0) >= 0) * np.maximum(b, 0) + (np.arange(1, K + 1) == 2)[:, None] * np.abs(b) + np.minimum(b, m)
           + b ** 2 / np.exp(tau))
    assert np.allclose(get("big"), big, rtol=1e-13)
    assert prog.outputs["ynew"][1] and np.isnan(get("ynew")).all()   # stochastic: left to the device


# ---- the kernel ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,text,fx,nodes", CASES, ids=[c[0] for c in CASES])
def test_device_matches_the_node_loop(fixtures, name, text, fx, nodes):
    f = fixtures[fx]
    gm = go.compile(text, f["data"], f["inits"]).use_generated_order()
    lw = jbugs.Lowering(text, f["data"])
    dq = DeviceGQ(lw)
    thetas = _thetas(gm, 37)
    out = dq.run(thetas, only_gq=False)
    _check_against_env(out, gm, thetas[:, :4], nodes, name)
    # and bit-for-bit what the numpy reading of the same program gives, up to the device's libm
    Q = run_program(dq.program, thetas)
    for k, (r, _) in dq.program.outputs.items():
        ref = Q[np.maximum(np.asarray(r), 0)]
        ok = ~np.isnan(ref) & (np.asarray(r)[..., None] >= 0)
        assert np.allclose(out[k][ok], ref[ok], rtol=1e-11, atol=0), (name, k)   # (device pow / exp vs numpy: ulps, amplified by a - b)
    assert dq.launch_count() == len(dq.program.stmts)


@pytest.mark.gpu
def test_device_row_subset_and_device_buffers(fixtures):
    """only the monitored rows are supplied (what a sampler returns); device-resident buffers are used in place"""
    import torch
    f = fixtures["rats"]
    lw = jbugs.Lowering(ex.RATS, f["data"])
    gm = go.compile(ex.RATS, f["data"], f["inits"]).use_generated_order()
    dq = DeviceGQ(lw)
    thetas = _thetas(gm, 200)
    names = lw.param_names()
    rows = np.array([names.index(n) for n in ("tau.c", "alpha.c", "beta.c")])
    with pytest.raises(jbugs.JBugsCudaError):
        dq.run(thetas[rows], rows=rows, only_gq=False)      # mu[i, j] needs alpha[i], beta[i]
    need = dq.program.theta_rows_needed()
    full = dq.run(thetas)
    sub = dq.run(thetas[need], rows=need)
    for k in full:
        assert np.array_equal(full[k], sub[k])
    assert np.allclose(full["sigma"], 1.0 / np.sqrt(np.exp(thetas[names.index("tau.c")])), rtol=1e-14)
    th = torch.from_numpy(thetas).cuda()
    out = torch.full((dq.program.n_rows, 256), float("nan"), dtype=torch.float64, device="cuda")
    dq.run_raw(th.data_ptr(), th.shape[1], th.shape[0], None, th.shape[1], 0, out.data_ptr(), 256)
    torch.cuda.synchronize()
    got = dq.collect(out.cpu().numpy()[:, :200])
    assert np.array_equal(got["alpha0"], full["alpha0"])


@pytest.mark.gpu
def test_forward_samplers_have_the_right_moments():
    """every proper family, BUGS parameterisation -> moments of the scipy distribution of tests/test_families.py"""
    from scipy import stats as S
    fams = [
        ("dnorm(1.5, 4.0)", S.norm(1.5, 0.5)), ("dgamma(3.0, 2.0)", S.gamma(3.0, scale=0.5)), ("dgamma(0.4, 2.0)", S.gamma(0.4, scale=0.5)),
        ("dexp(2.5)", S.expon(scale=0.4)), ("dunif(-1.0, 3.0)", S.uniform(-1.0, 4.0)), ("dbeta(2.0, 5.0)", S.beta(2.0, 5.0)),
        ("dlnorm(0.5, 4.0)", S.lognorm(s=0.5, scale=np.exp(0.5))), ("dlogis(0.5, 4.0)", S.logistic(0.5, 0.5)),
        ("ddexp(0.5, 4.0)", S.laplace(0.5, 0.5)), ("dweib(1.7, 0.8)", S.weibull_min(1.7, scale=1.25)), ("dchisqr(5)", S.chi2(5)),
        ("dbern(0.3)", S.bernoulli(0.3)), ("dbin(0.1, 10)", S.binom(10, 0.1)), ("dbin(0.7, 500)", S.binom(500, 0.7)),
        ("dpois(3.5)", S.poisson(3.5)), ("dpois(140.0)", S.poisson(140.0)), ("dnegbin(0.4, 3)", S.nbinom(3, 0.4)),
        ("dt(1.5, 4.0, 7)", S.t(7, 1.5, 0.5)), ("dgeom(0.3)", S.geom(0.3, loc=-1)), ("dbetabin(2.5, 1.5, 9)", S.betabinom(9, 2.5, 1.5)),
    ]
    body = "\n".join(f"  q{k} ~ {call}" for k, (call, _) in enumerate(fams))
    text = "model {\n  a ~ dnorm(0, 1)\n  y ~ dnorm(a, 1)\n" + body + "\n}"
    lw = jbugs.Lowering(text, {"y": 0.3})
    dq = DeviceGQ(lw)
    assert not dq.program.skipped, dq.program.skipped
    n = 400_000
    out = dq.run(np.zeros((1, n)), seed=11)
    again = dq.run(np.zeros((1, n)), seed=11)
    other = dq.run(np.zeros((1, n)), seed=12)
    for k, (call, dist) in enumerate(fams):
        x = out[f"q{k}"]
        assert np.array_equal(x, again[f"q{k}"]) and not np.array_equal(x, other[f"q{k}"]), call   # counter-based: reproducible
        m, v = dist.mean(), dist.var()
        se = np.sqrt(v / n)
        assert abs(x.mean() - m) < 6 * se, (call, x.mean(), m)
        assert abs(x.var() - v) < 0.03 * v, (call, x.var(), v)
        if hasattr(dist, "ppf") and not hasattr(dist.dist, "pmf"):
            qs = np.array([0.1, 0.5, 0.9])
            assert np.allclose(np.quantile(x, qs), dist.ppf(qs), atol=0.02 * np.sqrt(v)), call
    # two statements draw independent streams
    assert abs(np.corrcoef(out["q0"], out["q6"])[0, 1]) < 0.01


@pytest.mark.gpu
def test_chains_generated_quantities_on_the_device(fixtures):
    """Chains.generated_quantities(device=True) == the host twin for the deterministic quantities"""
    from jbugs.hmc import HMC, sample
    f = fixtures["rats"]
    m = jbugs.compile(ex.RATS, f["data"], f["inits"])
    ch = sample(m, HMC(n_leapfrog=8, step_size=0.02), n_samples=20, n_adapts=50, n_chains=32, seed=3)
    host = ch.generated_quantities()
    ch._gq = None
    dev = ch.generated_quantities(device=True)
    assert set(dev) == set(host) == {"sigma", "alpha0"}
    for k in host:
        assert np.allclose(np.asarray(dev[k]).ravel(), np.asarray(host[k]).ravel(), rtol=1e-12)
