"""shared model zoo for the tests: (name, model text, fixture key, inits key)."""
import numpy as np

from jbugs import examples as ex

ZOO = [
    ("rats", ex.RATS, "rats", "inits"),
    ("seeds", ex.SEEDS, "seeds", "inits"),
    ("surgical_realistic", ex.SURGICAL_REALISTIC, "surgical", "inits_realistic"),
    ("surgical_simple", ex.SURGICAL_SIMPLE, "surgical", "inits_simplistic"),
    ("pumps", ex.PUMPS, "pumps", "inits"),
    ("epil", ex.EPIL, "epil", "inits"),
    ("dogs", ex.DOGS, "dogs", "inits"),
    ("blockers", ex.BLOCKERS, "blockers", "inits"),   # two observation statements sharing mu[i] -> one N x 2 plate
    ("oxford", ex.OXFORD, "oxford", "inits"),         # same, with per-statement global covariates
    ("salm", ex.SALM, "salm", "inits"),
    ("equiv", ex.EQUIV, "equiv", "inits"),            # random effect indexed by the inner loop variable
    ("dyes", ex.DYES, "dyes", "inits"),
    ("stacks", ex.STACKS, "stacks", "inits"),         # beta[1], beta[2], beta[3] referenced element by element
    # predictors that are not linear in the parameters: expression plates (tape interpreter)
    ("bones", ex.BONES, "bones", "inits"),            # dcat of cumulative-logit probabilities, 20 missing cells
    ("dugongs", ex.DUGONGS, "dugongs", "inits"),      # alpha - beta * pow(gamma, x[i])
    ("lsat", ex.LSAT, "lsat", "inits"),               # logistic(beta * theta[j] - alpha[k]), ragged data transformation
    ("air", ex.AIR, "air", "inits"),                  # a latent covariate times a coefficient
    ("leuk", ex.LEUK, "leuk", "inits"),               # Y[i,j] * exp(beta * Z[i]) * dL0[j] as a Poisson mean
    ("mice", ex.MICE, "mice", "inits"),               # censored Weibull: the censored cells are missing, hence outside the target
    ("orange", ex.ORANGE, "orange", "inits"),         # vol. 2: a coefficient vector per group, theta[i, 1:3], column by column
    ("kidney", ex.KIDNEY, "kidney", "inits"),
    # six priors enumerated by a short outer loop (written out iteration by iteration), two observation statements per
    # study sharing pc[j, k] with a non-linear predictor (one merged expression plate per prior), a half-normal T(0, )
    ("magnesium", ex.MAGNESIUM, "magnesium", "inits"),
    # a frailty per matched pair, b[pair[i]], inside the Cox model's non-linear predictor: a parameter vector picked by a
    # data index on the tape (OP_GVEC_AT), 23 global values referenced by one expression plate
    ("leukfr", ex.LEUKFR, "leukfr", "inits"),         # + rows without any observed cell dropped (their b[i] leave theta), factor levels by data index
]


def start_theta(name, D, graph_model=None):
    """a finite starting point in unconstrained space for each example."""
    if graph_model is not None:
        try:
            th = graph_model.getparams()
        except ZeroDivisionError:   # partial inits (Magnesium): a derived precision is 1 / 0 at the zero-filled start
            th = np.full(D, np.nan)
        if np.all(np.isfinite(th)):
            return th
    return np.zeros(D)


def random_thetas(theta0, C, seed=0, scale=0.1):
    rng = np.random.default_rng(seed)
    th0 = np.where(np.isfinite(theta0), theta0, 0.0)
    return th0[:, None] + scale * rng.standard_normal((th0.size, C))


def synth_rats(N, T=5, seed=1234):
    """BASELINE.json config 2: Y_ij = alpha_i + beta_i (x_j - 22) + N(0, 6^2)."""
    rng = np.random.default_rng(seed)
    x = np.array([8.0, 15.0, 22.0, 29.0, 36.0])[:T]
    alpha = rng.normal(242.0, 14.0, N)
    beta = rng.normal(6.2, 0.5, N)
    Y = alpha[:, None] + beta[:, None] * (x[None, :] - 22.0) + rng.normal(0.0, 6.0, (N, T))
    return {"x": x, "xbar": 22.0, "N": N, "T": T, "Y": Y}, alpha, beta


def synth_seeds(N, seed=1234):
    """BASELINE.json config 3: random-effects logistic regression."""
    rng = np.random.default_rng(seed)
    x1 = rng.integers(0, 2, N).astype(float)
    x2 = rng.integers(0, 2, N).astype(float)
    n = rng.integers(10, 81, N).astype(float)
    b = rng.normal(0.0, 0.3, N)
    eta = -0.55 + 0.09 * x1 + 1.36 * x2 - 0.84 * x1 * x2 + b
    r = rng.binomial(n.astype(int), 1.0 / (1.0 + np.exp(-eta))).astype(float)
    return {"r": r, "n": n, "x1": x1, "x2": x2, "N": N}, b
