"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Number systems for the oracle: plain float64, forward-mode dual numbers
(the restatement of the reference's `AutoForwardDiff` backend,
`JuliaBUGS/src/model/logdensityproblems.jl:193-232` differentiating
`evaluate_with_values!!` end to end), and mpmath for high-precision cross
checks.  The elementary functions below are written once against the small
`Math` namespace so the distribution code in `dists.py` runs unchanged in all
three systems.
"""
from __future__ import annotations

import math

import numpy as np

try:  # mpmath is optional at run time (only the high-precision checks need it)
    import mpmath as _mp
except Exception:  # pragma: no cover
    _mp = None

from scipy import special as _sp


class DomainError(ArithmeticError):
    """Mirror of Julia's DomainError (mapped to -Inf / NaN-gradient at the boundary,
    `logdensityproblems.jl:22-26,226-229`)."""


class Dual:
    """value + gradient vector (forward mode, all partials at once)."""

    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v = float(v)
        self.d = d

    @staticmethod
    def seed(x, i, n):
        d = np.zeros(n)
        d[i] = 1.0
        return Dual(x, d)

    def _lift(self, o):
        return o if isinstance(o, Dual) else Dual(o, 0.0)

    def __add__(self, o):
        o = self._lift(o)
        return Dual(self.v + o.v, self.d + o.d)

    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o)
        return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        return Dual(self.v * o.v, self.d * o.v + o.d * self.v)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o)
        q = self.v / o.v
        return Dual(q, (self.d - q * o.d) / o.v)

    def __rtruediv__(self, o):
        return self._lift(o) / self

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __pow__(self, o):
        if isinstance(o, Dual):
            return FLOAT_OR_DUAL.exp(o * FLOAT_OR_DUAL.log(self))
        if o == 2:
            return self * self
        return Dual(self.v**o, o * self.v ** (o - 1) * self.d)

    def __rpow__(self, o):
        return FLOAT_OR_DUAL.exp(self * math.log(o))

    # comparisons act on the value (as ForwardDiff does)
    def __lt__(self, o):
        return self.v < _val(o)

    def __le__(self, o):
        return self.v <= _val(o)

    def __gt__(self, o):
        return self.v > _val(o)

    def __ge__(self, o):
        return self.v >= _val(o)

    def __float__(self):
        return self.v

    def __repr__(self):
        return f"Dual({self.v!r})"


def _val(x):
    return x.v if isinstance(x, Dual) else x


def _phi_pdf(x):
    return math.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


class _FloatDualMath:
    """Elementary functions on float | Dual."""

    name = "float64"
    log2pi = math.log(2.0 * math.pi)
    inf = math.inf

    def _un(self, x, f, df):
        if isinstance(x, Dual):
            return Dual(f(x.v), df(x.v) * x.d)
        return f(float(x))

    def const(self, x):
        return float(x)

    def value(self, x):
        return float(_val(x))

    def exp(self, x):
        return self._un(x, math.exp, math.exp)

    def log(self, x):
        def f(v):
            if v < 0:
                raise DomainError(f"log({v})")
            return math.log(v) if v > 0 else -math.inf

        return self._un(x, f, lambda v: 1.0 / v if v != 0 else math.inf)

    def log1p(self, x):
        def f(v):
            if v < -1:
                raise DomainError(f"log1p({v})")
            return math.log1p(v) if v > -1 else -math.inf

        return self._un(x, f, lambda v: 1.0 / (1.0 + v) if v != -1 else math.inf)

    def sqrt(self, x):
        def f(v):
            if v < 0:
                raise DomainError(f"sqrt({v})")
            return math.sqrt(v)

        return self._un(x, f, lambda v: 0.5 / math.sqrt(v) if v > 0 else math.inf)

    def lgamma(self, x):
        return self._un(x, math.lgamma, lambda v: float(_sp.digamma(v)))

    def logistic(self, x):
        # LogExpFunctions.logistic: e = exp(x); e / (1 + e) with saturation in the tails
        def f(v):
            if v >= 0:
                return 1.0 / (1.0 + math.exp(-v))
            e = math.exp(v)
            return e / (1.0 + e)

        return self._un(x, f, lambda v: f(v) * (1.0 - f(v)))

    def cexpexp(self, x):
        # 1 - exp(-exp(x)) (functions.jl:15-17 -> LogExpFunctions.cexpexp = -expm1(-exp(x)))
        return self._un(x, lambda v: -math.expm1(-math.exp(v)), lambda v: math.exp(v - math.exp(v)))

    def phi(self, x):
        # standard normal cdf (functions.jl:87-89)
        return self._un(x, lambda v: 0.5 * math.erfc(-v / math.sqrt(2.0)), _phi_pdf)

    def abs(self, x):
        return self._un(x, abs, lambda v: 1.0 if v >= 0 else -1.0)

    def sin(self, x):
        return self._un(x, math.sin, math.cos)

    def cos(self, x):
        return self._un(x, math.cos, lambda v: -math.sin(v))

    def isint(self, x):
        return float(_val(x)).is_integer()


class _MpMath:
    """Elementary functions on mpmath.mpf (logp only, no derivatives)."""

    name = "mpmath"

    def __init__(self, dps=50):
        if _mp is None:  # pragma: no cover
            raise RuntimeError("mpmath is not available")
        _mp.mp.dps = dps
        self.log2pi = _mp.log(2 * _mp.pi)
        self.inf = _mp.inf

    def const(self, x):
        return _mp.mpf(x)

    def value(self, x):
        return float(x)

    def exp(self, x):
        return _mp.exp(x)

    def log(self, x):
        if x < 0:
            raise DomainError("log of negative")
        return _mp.log(x) if x > 0 else -_mp.inf

    def log1p(self, x):
        if x < -1:
            raise DomainError("log1p")
        return _mp.log1p(x) if x > -1 else -_mp.inf

    def sqrt(self, x):
        if x < 0:
            raise DomainError("sqrt of negative")
        return _mp.sqrt(x)

    def lgamma(self, x):
        return _mp.loggamma(x)

    def logistic(self, x):
        return 1 / (1 + _mp.exp(-x))

    def cexpexp(self, x):
        return -_mp.expm1(-_mp.exp(x))

    def phi(self, x):
        return _mp.ncdf(x)

    def abs(self, x):
        return abs(x)

    def sin(self, x):
        return _mp.sin(x)

    def cos(self, x):
        return _mp.cos(x)

    def isint(self, x):
        return x == int(x)


FLOAT_OR_DUAL = _FloatDualMath()


def mp_math(dps=50):
    return _MpMath(dps)
