"""Seeded generator of random BUGS models inside the plate class (used by tests/test_random_models.py).

Every model is one observation plate (one or two loops) with a random observation family and link, a linear predictor
made of global coefficients x data covariates, group-level and observation-level random effects and a data offset,
random priors for the random effects (constant, data or global arguments) and for the globals (real-line, positive
and bounded supports, scalar logical nodes).  The point is coverage of combinations nobody wrote a fixture for."""
import numpy as np

OBS = [
    # (name, statement template, link options, aux globals [(name, prior)], y generator)
    ("normal", "y[{ix}] ~ dnorm(mu[{ix}], tau.y)", ["identity"], [("tau.y", "dgamma(2, 2)")],
     lambda rng, shp: rng.normal(0.3, 1.0, shp)),
    ("logistic", "y[{ix}] ~ dlogis(mu[{ix}], tau.y)", ["identity"], [("tau.y", "dgamma(2, 2)")],
     lambda rng, shp: rng.logistic(0.3, 1.0, shp)),
    ("laplace", "y[{ix}] ~ ddexp(mu[{ix}], tau.y)", ["identity"], [("tau.y", "dexp(1.0)")],
     lambda rng, shp: rng.laplace(0.3, 1.0, shp)),
    ("student", "y[{ix}] ~ dt(mu[{ix}], tau.y, 4)", ["identity"], [("tau.y", "dgamma(2, 2)")],
     lambda rng, shp: rng.standard_t(4, shp)),
    ("lognormal", "y[{ix}] ~ dlnorm(mu[{ix}], tau.y)", ["identity"], [("tau.y", "dgamma(2, 2)")],
     lambda rng, shp: rng.lognormal(0.1, 0.7, shp)),
    ("bernoulli", "y[{ix}] ~ dbern(mu[{ix}])", ["logit", "probit", "cloglog"], [],
     lambda rng, shp: (rng.random(shp) < 0.4).astype(float)),
    ("binomial", "y[{ix}] ~ dbin(mu[{ix}], n[{ix}])", ["logit", "probit", "cloglog"], [],
     lambda rng, shp: rng.integers(0, 6, shp).astype(float)),
    ("poisson", "y[{ix}] ~ dpois(mu[{ix}])", ["log"], [],
     lambda rng, shp: rng.poisson(2.0, shp).astype(float)),
    ("negbin", "y[{ix}] ~ dnegbin(mu[{ix}], r.y)", ["logit"], [("r.y", "dgamma(3, 1)")],
     lambda rng, shp: rng.negative_binomial(3, 0.5, shp).astype(float)),
    ("betabin", "y[{ix}] ~ dbetabin(mu[{ix}], b.y, 6)", ["log"], [("b.y", "dgamma(3, 1)")],
     lambda rng, shp: rng.integers(0, 7, shp).astype(float)),
    ("geometric", "y[{ix}] ~ dgeom(mu[{ix}])", ["logit", "probit", "cloglog"], [],
     lambda rng, shp: (rng.geometric(0.4, shp) - 1).astype(float)),
    ("gamma", "y[{ix}] ~ dgamma(a.y, mu[{ix}])", ["log"], [("a.y", "dgamma(3, 1)")],
     lambda rng, shp: rng.gamma(2.0, 0.8, shp) + 0.01),
    ("weibull", "y[{ix}] ~ dweib(a.y, mu[{ix}])", ["log"], [("a.y", "dgamma(3, 2)")],
     lambda rng, shp: rng.weibull(1.4, shp) + 0.02),
    ("exponential", "y[{ix}] ~ dexp(mu[{ix}])", ["log"], [],
     lambda rng, shp: rng.exponential(0.8, shp) + 0.01),
]
LINK_LHS = {"identity": "mu[{ix}]", "logit": "logit(mu[{ix}])", "probit": "probit(mu[{ix}])",
            "cloglog": "cloglog(mu[{ix}])", "log": "log(mu[{ix}])"}
GLOBAL_PRIORS = ["dnorm(0.0, 0.5)", "dnorm(0.2, 2.0)", "dlogis(0.0, 1.0)", "ddexp(0.0, 1.0)", "dunif(-3.0, 3.0)"]
POS_PRIORS = ["dgamma(2, 2)", "dexp(1.0)", "dlnorm(0.0, 1.0)", "dweib(1.5, 1.0)", "dchisqr(3)", "dunif(0.1, 4.0)"]


def make(seed, n_max=12):
    """-> (model text, data dict, description)"""
    rng = np.random.default_rng(1000 + seed)
    two = bool(rng.integers(0, 2))
    N = int(rng.integers(1, n_max))
    T = int(rng.integers(1, 4)) if two else 1
    shp = (N, T) if two else (N,)
    ix = "i, j" if two else "i"
    name, obs_stmt, links, aux, ygen = OBS[int(rng.integers(0, len(OBS)))]
    link = links[int(rng.integers(0, len(links)))]
    data = {"N": N, "y": ygen(rng, shp)}
    if two:
        data["T"] = T
    if name == "binomial":
        data["n"] = data["y"] + rng.integers(0, 5, shp).astype(float)
    globals_, body, tail, terms = [], [], [], []
    # global coefficients
    ng = int(rng.integers(1, 4))
    for k in range(ng):
        g = f"b{k}"
        globals_.append((g, GLOBAL_PRIORS[int(rng.integers(0, len(GLOBAL_PRIORS)))]))
        kind = int(rng.integers(0, 4)) if k else 0
        if kind == 0:
            terms.append(g)
        elif kind == 1:
            data[f"x{k}"] = rng.standard_normal(N)
            terms.append(f"{g} * x{k}[i]")
        elif kind == 2 and two:
            data[f"z{k}"] = rng.standard_normal(T)
            terms.append(f"{g} * (z{k}[j] - zbar{k})")
            data[f"zbar{k}"] = float(rng.standard_normal())
        else:
            data[f"w{k}"] = rng.standard_normal(shp)
            terms.append(f"{g} * w{k}[{ix}] / 2")
    # group-level random effects
    for k in range(int(rng.integers(0, 3))):
        l = f"u{k}"
        choice = int(rng.integers(0, 4))
        if choice == 0:
            prior = "dnorm(0.0, 1.0)"
        elif choice == 1:
            globals_.append((f"m.{l}", "dnorm(0.0, 1.0)"))
            globals_.append((f"t.{l}", POS_PRIORS[int(rng.integers(0, len(POS_PRIORS)))]))
            prior = f"dnorm(m.{l}, t.{l})"
        elif choice == 2:
            data[f"pm.{l}"] = 0.3 * rng.standard_normal(N)
            prior = f"dlogis(pm.{l}[i], 2.0)"
        else:
            globals_.append((f"s.{l}", "dunif(0.2, 3.0)"))
            tail.append(f"t.{l} <- 1 / (s.{l} * s.{l})")
            prior = f"ddexp(0.0, t.{l})"
        body.append(f"{l}[i] ~ {prior}")
        if two and rng.integers(0, 2):
            data[f"c.{l}"] = rng.standard_normal(T)
            terms.append(f"{l}[i] * c.{l}[j]")
        else:
            terms.append(f"{l}[i]")
    # observation-level random effect
    inner = []
    if two and rng.integers(0, 2):
        globals_.append(("t.e", "dgamma(2, 2)"))
        inner.append("e[i, j] ~ dnorm(0.0, t.e)")
        terms.append("e[i, j]")
    if rng.integers(0, 2):
        data["off"] = 0.2 * rng.standard_normal(shp)
        terms.append(f"off[{ix}]")
    globals_ += aux
    pred = " + ".join(terms)
    stmts = [obs_stmt.format(ix=ix), f"{LINK_LHS[link].format(ix=ix)} <- {pred}"] + inner
    lines = ["model {", "  for (i in 1:N) {"]
    if two:
        lines.append("    for (j in 1:T) {")
        lines += ["      " + s for s in stmts]
        lines.append("    }")
    else:
        lines += ["    " + s for s in stmts]
    lines += ["    " + s for s in body]
    lines.append("  }")
    lines += [f"  {g} ~ {p}" for g, p in globals_] + ["  " + t for t in tail]
    lines.append("}")
    return "\n".join(lines), data, f"{name}/{link} N={N} T={T}"


def make_structured(seed):
    """random models exercising the structural lowering paths: random effects picked by a data index (regrouped, padded,
    weighted plate) or sibling observation statements sharing a random effect (merged N x K plate).
    -> (model text, data dict, description)"""
    rng = np.random.default_rng(5000 + seed)
    name, obs_stmt, links, aux, ygen = OBS[int(rng.integers(0, len(OBS)))]
    link = links[int(rng.integers(0, len(links)))]
    globals_ = [("b0", GLOBAL_PRIORS[int(rng.integers(0, len(GLOBAL_PRIORS)))]), ("b1", "dnorm(0.0, 1.0)")] + list(aux)
    if rng.integers(0, 2) and name not in ("student", "betabin"):   # (the third argument and the padding weight share a slot)
        # --- gathered: y[i] ~ f(b0 + b1 x[i] + u[g[i]] (+ v[g[i]] z[i]))
        n, G = int(rng.integers(8, 60)), int(rng.integers(2, 7))
        grp = np.concatenate([np.arange(1, G + 1), rng.integers(1, G + 1, n - G)]).astype(float)
        rng.shuffle(grp)
        data = {"n": n, "G": G, "g": grp, "x": rng.standard_normal(n), "y": ygen(rng, (n,))}
        if name == "binomial":
            data["n.tr"] = data["y"] + rng.integers(0, 5, n).astype(float)
        terms = ["b0", "b1 * x[i]", "u[g[i]]"]
        body = ["u[k] ~ dnorm(0.0, t.u)"]
        globals_.append(("t.u", POS_PRIORS[int(rng.integers(0, len(POS_PRIORS)))]))
        if rng.integers(0, 2):
            data["z"] = rng.standard_normal(n)
            terms.append("v[g[i]] * z[i]")
            body.append("v[k] ~ dlogis(0.0, 2.0)")
        stmt = obs_stmt.format(ix="i").replace("n[i]", "n.tr[i]")
        lines = ["model {", "  for (i in 1:n) {", "    " + stmt, f"    {LINK_LHS[link].format(ix='i')} <- " + " + ".join(terms), "  }",
                 "  for (k in 1:G) {"] + ["    " + b for b in body] + ["  }"]
        desc = f"gathered {name}/{link} n={n} G={G}"
    else:
        # --- siblings: K statements of one loop sharing u[i], each with its own covariate and offset
        N, K = int(rng.integers(2, 25)), int(rng.integers(2, 4))
        data = {"N": N}
        lines = ["model {", "  for (i in 1:N) {"]
        for k in range(K):
            data[f"y{k}"] = ygen(rng, (N,))
            data[f"x{k}"] = rng.standard_normal(N)
            if name == "binomial":
                data[f"n{k}"] = data[f"y{k}"] + rng.integers(0, 5, N).astype(float)
            stmt = obs_stmt.format(ix="i").replace("y[i]", f"y{k}[i]").replace("mu[i]", f"mu{k}[i]").replace("n[i]", f"n{k}[i]")
            lhs = LINK_LHS[link].format(ix="i").replace("mu[i]", f"mu{k}[i]")
            lines += ["    " + stmt, f"    {lhs} <- b0 + b1 * x{k}[i] + u[i]" + (f" - {0.1 * k}" if k else "")]
        lines += ["    u[i] ~ dnorm(0.0, t.u)", "  }"]
        globals_.append(("t.u", POS_PRIORS[int(rng.integers(0, len(POS_PRIORS)))]))
        desc = f"siblings {name}/{link} N={N} K={K}"
    lines += [f"  {g} ~ {p}" for g, p in globals_] + ["}"]
    return "\n".join(lines), data, desc


def make_nonlinear(seed, n_max=12):
    """the same random plates with a predictor that is NOT linear in the parameters (products of parameters, parameters
    inside exp / pow): they lower to expression plates (tape interpreter).  -> (model text, data dict, description)"""
    text, data, desc = make(seed, n_max=n_max)
    rng = np.random.default_rng(9000 + seed)
    lines = text.split("\n")
    k = next(q for q, l in enumerate(lines) if "<-" in l and ("mu[" in l.split("<-")[0]))
    lhs, pred = lines[k].split("<-", 1)
    pred = pred.strip()
    kind = int(rng.integers(0, 4))
    extra = ["  sc ~ dnorm(0.0, 1.0)"]
    if kind == 0:
        new = f"({pred}) * exp(0.3 * sc)"
    elif kind == 1:
        new = f"{pred} + 0.1 * b0 * sc"
    elif kind == 2:
        new = f"pow(1.0 + exp(sc), -0.5) * ({pred})"
    else:
        new = f"{pred} + 0.2 * sc * u0[i]" if "u0[i]" in pred else f"{pred} - 0.1 * sc * sc"
    lines[k] = f"{lhs}<- {new}"
    end = max(q for q, l in enumerate(lines) if l.strip() == "}")
    lines[end:end] = extra
    return "\n".join(lines), data, desc + f" nonlinear#{kind}"
