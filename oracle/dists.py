"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Restatement of the arithmetic the reference delegates to third-party Julia
packages that are NOT vendored under /root/reference (no Manifest; compat
ranges in `JuliaBUGS/Project.toml:53-84`):

* BUGS -> Distributions.jl parameterisation: `JuliaBUGS/src/BUGSPrimitives/distributions.jl`
  (`dnorm :11-18`, `dexp :186-188`, `dgamma :200-203`, `dunif :343-345`, `dbeta :357-359`,
  `dbern :442-444`, `dbin :458-460`, `dcat :472-474`, `dpois :486-488`, `dlnorm`, `dweib`, ...).
* `Distributions.logpdf` (Distributions 0.25 / StatsFuns 1.x published formulas:
  `normlogpdf`, `gammalogpdf`, `betalogpdf`, `binomlogpdf`, `poislogpdf`).
* `Bijectors.bijector` / `with_logabsdet_jacobian` (Bijectors 0.13-0.16): support-based
  `TruncatedBijector(lb, ub)` for continuous univariates, identity for discrete; call site
  `JuliaBUGS/src/model/evaluation.jl:243-257`.

Every distribution is a small object with `logpdf(M, x)`, `support()` -> (lb, ub, discrete)
and parameter checks that raise `DomainError` exactly where Distributions' `@check_args`
(or the explicit throw in `dnorm`) would.
"""
from __future__ import annotations

import math

from numeric import DomainError, _val


def _check(cond, msg):
    if not cond:
        raise DomainError(msg)


class Dist:
    discrete = False

    def support(self):
        return (-math.inf, math.inf)


class Normal(Dist):
    def __init__(self, M, mu, tau):
        # distributions.jl:11-18: throws for tau < 0, sigma = sqrt(1/tau)
        _check(not (_val(tau) < 0), "dnorm requires tau > 0")
        self.mu = mu
        self.sigma = M.sqrt(1.0 / tau)

    def logpdf(self, M, x):
        z = (x - self.mu) / self.sigma
        return -(z * z + M.log2pi) / 2 - M.log(self.sigma)


class Gamma(Dist):
    def __init__(self, M, a, b):
        # distributions.jl:200-203: Gamma(a, 1/b); Distributions checks shape>0, scale>0
        self.k = a
        self.theta = 1.0 / b
        _check(_val(a) > 0 and _val(self.theta) > 0, "Gamma requires shape > 0 and scale > 0")

    def support(self):
        return (0.0, math.inf)

    def logpdf(self, M, x):
        if _val(x) < 0:
            return -M.inf
        xt = x / self.theta
        # StatsFuns.gammalogpdf: -loggamma(k) + xlogy(k-1, x/theta) - log(theta) - x/theta
        km1 = self.k - 1
        xl = 0.0 if (_val(km1) == 0 and not (_val(xt) != _val(xt))) else km1 * M.log(xt)
        return -M.lgamma(self.k) + xl - M.log(self.theta) - xt


class Exponential(Dist):
    def __init__(self, M, lam):
        # distributions.jl:186-188: Exponential(1/lam); scale must be > 0
        self.theta = 1.0 / lam
        _check(_val(self.theta) > 0, "Exponential requires scale > 0")

    def support(self):
        return (0.0, math.inf)

    def logpdf(self, M, x):
        if _val(x) < 0:
            return -M.inf
        rate = 1.0 / self.theta
        return M.log(rate) - rate * x


class Uniform(Dist):
    def __init__(self, M, a, b):
        _check(_val(a) < _val(b), "Uniform requires a < b")
        self.a, self.b = a, b

    def support(self):
        return (self.a, self.b)

    def logpdf(self, M, x):
        if _val(self.a) <= _val(x) <= _val(self.b):
            return -M.log(self.b - self.a)
        return -M.inf


def _logbeta(M, a, b):
    return M.lgamma(a) + M.lgamma(b) - M.lgamma(a + b)


def _xlogy(M, x, y):
    return 0.0 if _val(x) == 0 else x * M.log(y)


def _xlog1py(M, x, y):
    return 0.0 if _val(x) == 0 else x * M.log1p(y)


def _betalogpdf(M, a, b, x):
    # StatsFuns.betalogpdf
    return _xlogy(M, a - 1, x) + _xlog1py(M, b - 1, -x) - _logbeta(M, a, b)


class Beta(Dist):
    def __init__(self, M, a, b):
        _check(_val(a) > 0 and _val(b) > 0, "Beta requires a, b > 0")
        self.a, self.b = a, b

    def support(self):
        return (0.0, 1.0)

    def logpdf(self, M, x):
        if not (0 <= _val(x) <= 1):
            return -M.inf
        return _betalogpdf(M, self.a, self.b, x)


class LogNormal(Dist):
    def __init__(self, M, mu, tau):
        self.mu = mu
        self.sigma = M.sqrt(1.0 / tau)
        _check(_val(self.sigma) > 0, "LogNormal requires sigma > 0")

    def support(self):
        return (0.0, math.inf)

    def logpdf(self, M, x):
        if _val(x) <= 0:
            return -M.inf
        lx = M.log(x)
        z = (lx - self.mu) / self.sigma
        return -(z * z + M.log2pi) / 2 - M.log(self.sigma) - lx


class Bernoulli(Dist):
    discrete = True

    def __init__(self, M, p):
        _check(0 <= _val(p) <= 1, "Bernoulli requires 0 <= p <= 1")
        self.p = p

    def logpdf(self, M, x):
        xv = _val(x)
        if xv == 0:
            return M.log(1.0 - self.p)
        if xv == 1:
            return M.log(self.p)
        return -M.inf


class Binomial(Dist):
    discrete = True

    def __init__(self, M, p, n):
        # distributions.jl:458-460 swaps the arguments: dbin(p, n) = Binomial(n, p)
        _check(_val(n) >= 0, "Binomial requires n >= 0")
        _check(0 <= _val(p) <= 1, "Binomial requires 0 <= p <= 1")
        self.p, self.n = p, n

    def logpdf(self, M, k):
        kv, nv = _val(k), _val(self.n)
        if not (0 <= kv <= nv) or not M.isint(k):
            return -M.inf
        # StatsFuns.binomlogpdf: betalogpdf(k+1, n-k+1, p) - log(n+1)
        return _betalogpdf(M, k + 1, self.n - k + 1, self.p) - M.log(self.n + 1)


class Poisson(Dist):
    discrete = True

    def __init__(self, M, lam):
        _check(_val(lam) >= 0, "Poisson requires lambda >= 0")
        self.lam = lam

    def logpdf(self, M, k):
        if _val(k) < 0 or not M.isint(k):
            return -M.inf
        # StatsFuns.poislogpdf: xlogy(x, lambda) - lambda - loggamma(x+1)
        return _xlogy(M, k, self.lam) - self.lam - M.lgamma(k + 1)


class Categorical(Dist):
    discrete = True

    def __init__(self, M, p):
        s = 0.0
        for pi in p:
            _check(_val(pi) >= 0, "Categorical requires p >= 0")
            s = s + pi
        _check(abs(_val(s) - 1.0) < 1e-8, "Categorical requires sum(p) == 1")
        self.p = list(p)

    def logpdf(self, M, k):
        kv = int(_val(k))
        if not (1 <= kv <= len(self.p)):
            return -M.inf
        return M.log(self.p[kv - 1])


class Logistic(Dist):
    def __init__(self, M, mu, tau):
        # distributions.jl:30-33: Logistic(mu, 1/sqrt(tau)); sqrt throws for tau < 0, Distributions checks scale > 0
        self.mu = mu
        self.s = 1.0 / M.sqrt(tau)
        _check(_val(self.s) > 0, "Logistic requires scale > 0")

    def logpdf(self, M, x):
        # Distributions: u = -abs(z); u - 2 * log1pexp(u) - log(s)
        z = (x - self.mu) / self.s
        u = -M.abs(z)
        return u - 2.0 * M.log1p(M.exp(u)) - M.log(self.s)


class Laplace(Dist):
    def __init__(self, M, mu, tau):
        # distributions.jl:92-95: Laplace(mu, 1/sqrt(tau))
        self.mu = mu
        self.b = 1.0 / M.sqrt(tau)
        _check(_val(self.b) > 0, "Laplace requires scale > 0")

    def logpdf(self, M, x):
        # Distributions: -(abs(x - mu) / b + log(2 b))
        return -(M.abs(x - self.mu) / self.b + M.log(2.0 * self.b))


class Weibull(Dist):
    def __init__(self, M, a, b):
        # distributions.jl:233-235: Weibull(a, 1/b) (shape a, scale 1/b)
        self.a = a
        self.theta = 1.0 / b
        _check(_val(a) > 0 and _val(self.theta) > 0, "Weibull requires shape > 0 and scale > 0")

    def support(self):
        return (0.0, math.inf)

    def logpdf(self, M, x):
        if _val(x) < 0:
            return -M.inf
        # Distributions: z = x / theta; log(a / theta) + xlogy(a - 1, z) - z^a
        z = x / self.theta
        return M.log(self.a / self.theta) + _xlogy(M, self.a - 1, z) - M.exp(self.a * M.log(z))


class NegativeBinomial(Dist):
    discrete = True

    def __init__(self, M, p, r):
        # distributions.jl:516-518 swaps the arguments: dnegbin(p, r) = NegativeBinomial(r, p)
        _check(_val(r) > 0, "NegativeBinomial requires r > 0")
        _check(0 < _val(p) <= 1, "NegativeBinomial requires 0 < p <= 1")
        self.p, self.r = p, r

    def logpdf(self, M, k):
        if _val(k) < 0 or not M.isint(k):
            return -M.inf
        # Distributions: xlogy(r, p) + xlog1py(k, -p) - log(k + r) - logbeta(r, k + 1)
        return (_xlogy(M, self.r, self.p) + _xlog1py(M, k, -self.p) - M.log(k + self.r)
                - _logbeta(M, self.r, k + 1))


class Geometric(Dist):
    discrete = True

    def __init__(self, M, p):
        # distributions.jl:500-502: dgeom(p) = Geometric(p), the number of failures before the first success
        _check(0 < _val(p) <= 1, "Geometric requires 0 < p <= 1")
        self.p = p

    def logpdf(self, M, k):
        if _val(k) < 0 or not M.isint(k):
            return -M.inf
        # Distributions: xlog1py(k, -p) + log(p)
        return _xlog1py(M, k, -self.p) + M.log(self.p)


class BetaBinomial(Dist):
    discrete = True

    def __init__(self, M, a, b, n):
        # distributions.jl:530-532: dbetabin(a, b, n) = BetaBinomial(n, a, b); Distributions checks n >= 0, a >= 0, b >= 0
        _check(_val(n) >= 0, "BetaBinomial requires n >= 0")
        _check(_val(a) >= 0 and _val(b) >= 0, "BetaBinomial requires a >= 0 and b >= 0")
        self.a, self.b, self.n = a, b, n

    def logpdf(self, M, k):
        kv, nv = _val(k), _val(self.n)
        if not (0 <= kv <= nv) or not M.isint(k):
            return -M.inf
        # Distributions: -log1p(n) - logbeta(k + 1, n - k + 1) + logbeta(k + a, n - k + b) - logbeta(a, b)
        n = self.n
        return (-M.log1p(n + 0.0 * self.a) - _logbeta(M, k + 1, n - k + 1) + _logbeta(M, k + self.a, n - k + self.b)
                - _logbeta(M, self.a, self.b))


class Chisq(Dist):
    def __init__(self, M, k):
        # distributions.jl:215-217: Chisq(k) = Gamma(k/2, 2)
        _check(_val(k) > 0, "Chisq requires k > 0")
        self.k = k

    def support(self):
        return (0.0, math.inf)

    def logpdf(self, M, x):
        if _val(x) < 0:
            return -M.inf
        h = self.k / 2.0
        return -M.lgamma(h) - h * math.log(2.0) + _xlogy(M, h - 1, x) - x / 2.0


class TDistShiftedScaled(Dist):
    def __init__(self, M, mu, tau, nu):
        # distributions.jl:73-80: dt(mu, tau, nu) = TDistShiftedScaled(nu, mu, sqrt(1/tau)) (TDist(nu) when mu = 0, sigma = 1:
        # the same density); logpdf = tdistlogpdf(nu, (x - mu) / sigma) - log(sigma)   (:58-60)
        _check(not (_val(tau) < 0), "dt requires tau > 0")
        self.mu, self.nu = mu, nu
        self.sigma = M.sqrt(1.0 / tau)
        _check(_val(nu) > 0 and _val(self.sigma) > 0, "dt requires nu > 0 and sigma > 0")

    def logpdf(self, M, x):
        # StatsFuns.tdistlogpdf: loggamma((nu+1)/2) - log(nu pi)/2 - loggamma(nu/2) - (nu+1)/2 log1p(z^2/nu)
        z = (x - self.mu) / self.sigma
        nu = self.nu
        return (M.lgamma((nu + 1.0) / 2.0) - M.log(nu * math.pi) / 2.0 - M.lgamma(nu / 2.0)
                - (nu + 1.0) / 2.0 * M.log1p(z * z / nu) - M.log(self.sigma))


class Flat(Dist):
    def logpdf(self, M, x):
        return 0.0 * x


class Censored(Dist):
    """`censored(d, lower, upper)` (Distributions.Censored; the BUGS suffix C(lower, upper), bugs_parser.jl:470-500):
    the density of `d` strictly inside (lower, upper), the tail mass of `d` on a bound, -Inf outside.  The tail masses
    need the cdf of `d`; only the interior case occurs in the reference's examples once the missing (censored) cells
    have dropped out of the target as generated quantities, so a value ON a bound raises instead of guessing."""

    def __init__(self, M, inner, lower, upper):
        self.inner, self.lower, self.upper = inner, lower, upper
        self.discrete = inner.discrete

    def support(self):
        lb, ub = self.inner.support()
        return (max(_val(lb), _val(self.lower)), min(_val(ub), _val(self.upper)))

    def logpdf(self, M, x):
        xv = _val(x)
        if xv < _val(self.lower) or xv > _val(self.upper):
            return -M.inf
        if xv == _val(self.lower) or xv == _val(self.upper):
            raise NotImplementedError("an observation on its censoring bound needs the cdf of the inner distribution")
        return self.inner.logpdf(M, x)


def _scipy_frozen(inner):
    """scipy.stats twin of an oracle distribution with constant arguments (for the cdf of a truncated prior)"""
    from scipy import stats as S
    const = lambda *vs: all(isinstance(v, (int, float)) for v in vs)   # noqa: E731
    if isinstance(inner, Normal) and const(inner.mu, inner.sigma):
        return S.norm(inner.mu, inner.sigma)
    if isinstance(inner, LogNormal) and const(inner.mu, inner.sigma):
        return S.lognorm(s=inner.sigma, scale=math.exp(inner.mu))
    if isinstance(inner, Logistic) and const(inner.mu, inner.s):
        return S.logistic(inner.mu, inner.s)
    if isinstance(inner, Laplace) and const(inner.mu, inner.b):
        return S.laplace(inner.mu, inner.b)
    if isinstance(inner, Exponential) and const(inner.theta):
        return S.expon(scale=inner.theta)
    if isinstance(inner, Gamma) and const(inner.k, inner.theta):
        return S.gamma(inner.k, scale=inner.theta)
    if isinstance(inner, Weibull) and const(inner.a, inner.theta):
        return S.weibull_min(inner.a, scale=inner.theta)
    if isinstance(inner, Chisq) and const(inner.k):
        return S.chi2(inner.k)
    if isinstance(inner, Uniform) and const(inner.a, inner.b):
        return S.uniform(inner.a, inner.b - inner.a)
    if isinstance(inner, Beta) and const(inner.a, inner.b):
        return S.beta(inner.a, inner.b)
    if isinstance(inner, TDistShiftedScaled) and const(inner.mu, inner.sigma, inner.nu):
        return S.t(inner.nu, loc=inner.mu, scale=inner.sigma)
    raise NotImplementedError("T(,) is restated for dnorm, dlnorm, dlogis, ddexp, dexp, dgamma, dweib, dchisqr, dunif, dbeta "
                              "and dt with constant arguments")


class TruncatedNormal(Dist):
    """`truncated(dist(args), lower, upper)` (Distributions.truncated; the BUGS suffix T(lower, upper),
    bugs_parser.jl:470-500): logpdf(x) = logpdf_inner(x) - log(cdf(upper) - cdf(lower)) inside the bounds, -Inf outside;
    the support -- hence the bijector -- is the intersection of [lower, upper] with the inner support.  The normaliser
    is written for constant arguments and bounds (half-normal / truncated priors, e.g. Magnesium's
    `tau.sqrd[6] ~ dnorm(0, p0) T(0,)`); it comes from scipy's cdf / sf on the side where the difference does not
    cancel -- an implementation independent of the product's include/jbugs_trunc.h.  (The class keeps the name it had
    when only dnorm was restated.)"""

    def __init__(self, M, inner, lower, upper):
        for v in (lower, upper):
            if not isinstance(v, (int, float)):
                raise NotImplementedError("T(,) needs constant bounds")
        fz = _scipy_frozen(inner)
        lo, hi = inner.support()
        self.inner, self.lower, self.upper = inner, max(float(lower), float(lo)), min(float(upper), float(hi))
        _check(self.lower < self.upper, "truncated requires lower < upper")
        mass = (fz.sf(self.lower) - fz.sf(self.upper)) if fz.cdf(self.lower) > 0.5 else (fz.cdf(self.upper) - fz.cdf(self.lower))
        self.logz = math.log(mass)

    def support(self):
        return (self.lower, self.upper)

    def logpdf(self, M, x):
        xv = _val(x)
        if xv < self.lower or xv > self.upper:
            return -M.inf
        return self.inner.logpdf(M, x) - self.logz


def make_dist(M, name, args):
    """BUGS distribution call -> distribution object (BUGSPrimitives/distributions.jl)."""
    if name.startswith("truncated:"):
        return TruncatedNormal(M, make_dist(M, name.split(":", 1)[1], args[:-2]), args[-2], args[-1])
    if name.startswith("censored:"):
        return Censored(M, make_dist(M, name.split(":", 1)[1], args[:-2]), args[-2], args[-1])
    if name == "dnorm":
        return Normal(M, *args)
    if name == "dgamma":
        return Gamma(M, *args)
    if name == "dexp":
        return Exponential(M, *args)
    if name == "dunif":
        return Uniform(M, *args)
    if name == "dbeta":
        return Beta(M, *args)
    if name == "dlnorm":
        return LogNormal(M, *args)
    if name == "dbern":
        return Bernoulli(M, *args)
    if name == "dbin":
        return Binomial(M, *args)
    if name == "dpois":
        return Poisson(M, *args)
    if name == "dcat":
        return Categorical(M, *args)
    if name == "dlogis":
        return Logistic(M, *args)
    if name == "ddexp":
        return Laplace(M, *args)
    if name == "dweib":
        return Weibull(M, *args)
    if name == "dnegbin":
        return NegativeBinomial(M, *args)
    if name == "dbetabin":
        return BetaBinomial(M, *args)
    if name == "dgeom":
        return Geometric(M, *args)
    if name == "dchisqr":
        return Chisq(M, *args)
    if name == "dt":
        return TDistShiftedScaled(M, *args)
    if name == "dflat":
        return Flat()
    raise NotImplementedError(f"distribution {name} is outside the restated subset")


# --------------------------------------------------------------------------- bijectors


def link(M, dist, x):
    """constrained value -> unconstrained slot (Bijectors.bijector(dist)(x)); used by getparams,
    `JuliaBUGS/src/model/bugsmodel.jl:619-665`."""
    if dist.discrete:
        return x
    lb, ub = dist.support()
    lo, hi = math.isfinite(_val(lb)), math.isfinite(_val(ub))
    if lo and hi:
        u = (x - lb) / (ub - lb)
        return M.log(u) - M.log(1.0 - u)
    if lo:
        return M.log(x - lb)
    if hi:
        return M.log(ub - x)
    return x


def invlink_with_logjac(M, dist, y):
    """unconstrained slot -> (value, logabsdetjac) -- `with_logabsdet_jacobian(inverse(bijector(d)), y)`
    at `evaluation.jl:244-251`."""
    if dist.discrete:
        return y, 0.0
    lb, ub = dist.support()
    lo, hi = math.isfinite(_val(lb)), math.isfinite(_val(ub))
    if lo and hi:
        x = lb + (ub - lb) * M.logistic(y)
        return x, M.log((x - lb) * (ub - x) / (ub - lb))
    if lo:
        return M.exp(y) + lb, y
    if hi:
        return ub - M.exp(y), y
    return y, 0.0
