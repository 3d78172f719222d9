"""Generated quantities for the draws of a batched sampler run.

The reference re-evaluates the whole model once per draw on the CPU to fill in the nodes that are not part of the
target: logical nodes (`sigma <- 1 / sqrt(tau)`, `alpha0 <- alpha.c - xbar * beta.c`) and unobserved stochastic
nodes without observed descendants (`delta.new ~ dnorm(d, tau)`), see `J/src/model/abstractmcmc.jl:92-110` and the
forward-sampling branch of `evaluate_with_rng!!` (`J/src/model/evaluation.jl:354-378`).  Here every statement is
evaluated ONCE for all draws: parameter arrays carry the draws on a trailing axis and the statement's loop domain is
a numpy index grid, so the cost is one vectorised expression per BUGS statement, independent of the number of draws.

Host-side post-processing only -- the hot path (log density + gradient) never goes through this module.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np

from .lowering import DataEval, Lowering, LoweringError, _grids
from .parser import Call, Colon, Index, Range, Var, free_vars


class _DrawEval(DataEval):
    """DataEval over an environment in which some variables carry a trailing draw axis of length S."""

    def __init__(self, data, draws: Dict[str, np.ndarray], S: int):
        super().__init__(data)
        self.draws = draws
        self.S = S

    def known(self, e, loop_vars=()) -> bool:
        return all((v in self.data) or (v in self.draws) or (v in loop_vars) for v in free_vars(e))

    def has_draws(self, e) -> bool:
        return any(v in self.draws for v in free_vars(e))

    def ev(self, e, lv):
        if isinstance(e, Var) and e.name in self.draws and e.name not in lv:
            return self.draws[e.name]
        if isinstance(e, Index) and e.name in self.draws:
            arr = self.draws[e.name]
            idx = []
            for ie in e.idx:
                if isinstance(ie, Colon):
                    idx.append(slice(None))
                elif isinstance(ie, Range):
                    lo, hi = self.ev(ie.lo, lv), self.ev(ie.hi, lv)
                    if np.ndim(lo) or np.ndim(hi):
                        raise LoweringError("loop-dependent index range over a parameter array")
                    idx.append(slice(int(lo) - 1, int(hi)))
                else:
                    if self.has_draws(ie):
                        raise LoweringError("array index depends on a parameter")
                    iv = np.asarray(self.ev(ie, lv))
                    if iv.ndim:
                        iv = iv[..., 0]  # grids carry a trailing unit axis for the draws
                    idx.append(iv.astype(np.int64) - 1)
            if any(isinstance(i, slice) for i in idx) and any(isinstance(i, np.ndarray) and i.ndim for i in idx):
                raise LoweringError("mixed slice / loop index over a parameter array")
            return arr[tuple(idx)]
        if isinstance(e, Call) and e.fn in ("mean", "sum", "sd", "inprod") and self.has_draws(e):
            args = [np.asarray(self.ev(a, lv), dtype=np.float64) for a in e.args]
            flat = [a.reshape(-1, self.S) if a.shape[-1:] == (self.S,) and self.has_draws(x) else a.reshape(-1, 1)
                    for a, x in zip(args, e.args)]
            if e.fn == "mean":
                return flat[0].mean(axis=0)
            if e.fn == "sum":
                return flat[0].sum(axis=0)
            if e.fn == "sd":
                return flat[0].std(axis=0, ddof=1)
            return (flat[0] * flat[1]).sum(axis=0)
        return super().ev(e, lv)


def _sample(rng, fn, args, shape):
    a = [np.broadcast_to(np.asarray(x, dtype=np.float64), shape) for x in args]
    if fn == "dnorm":
        return rng.normal(a[0], 1.0 / np.sqrt(a[1]))
    if fn == "dgamma":
        return rng.gamma(a[0], 1.0 / a[1])
    if fn == "dexp":
        return rng.exponential(1.0 / a[0])
    if fn == "dunif":
        return rng.uniform(a[0], a[1])
    if fn == "dbeta":
        return rng.beta(a[0], a[1])
    if fn == "dlnorm":
        return rng.lognormal(a[0], 1.0 / np.sqrt(a[1]))
    if fn == "dbern":
        return (rng.random(shape) < a[0]).astype(np.float64)
    if fn == "dbin":
        return rng.binomial(a[1].astype(np.int64), a[0]).astype(np.float64)
    if fn == "dpois":
        return rng.poisson(a[0]).astype(np.float64)
    if fn == "dlogis":
        return rng.logistic(a[0], 1.0 / np.sqrt(a[1]))
    if fn == "ddexp":
        return rng.laplace(a[0], 1.0 / np.sqrt(a[1]))
    if fn == "dweib":
        return rng.weibull(a[0]) / a[1]
    if fn == "dchisqr":
        return rng.chisquare(a[0])
    if fn == "dt":
        return a[0] + rng.standard_t(a[2]) / np.sqrt(a[1])
    if fn == "dnegbin":
        return rng.negative_binomial(a[1], a[0]).astype(np.float64)
    if fn == "dbetabin":
        return rng.binomial(a[2].astype(np.int64), rng.beta(a[0], a[1])).astype(np.float64)
    if fn == "dgeom":
        return (rng.geometric(a[0]) - 1).astype(np.float64)   # failures before the first success
    raise LoweringError(f"forward sampling of {fn} is not provided")


def generated_quantities(lowering: Lowering, draws: Dict[str, np.ndarray], seed: int = 0,
                         only_gq: bool = True) -> Dict[str, np.ndarray]:
    """`draws[label]` = array of S draws (any shape, flattened) for parameter labels such as 'tau' or 'beta[2]'
    (constrained space).  Returns name -> array of shape `dims + (S,)` for every logical node that can be computed
    from the supplied parameters and every unobserved stochastic node outside the target (forward-sampled).
    `only_gq=False` also returns the logical nodes that are part of the target (mu[i, j], p[i], ...)."""
    lw = lowering
    sizes = {np.asarray(v).size for v in draws.values()}
    if len(sizes) != 1:
        raise ValueError("all parameters must come with the same number of draws")
    S = sizes.pop()
    rng = np.random.default_rng(seed)
    env: Dict[str, np.ndarray] = {}
    # parameters: scalars as (S,), arrays as dims + (S,) when every element was supplied
    for s in lw.stmts:
        if s.kind != "param" or s.is_gq:
            continue
        labels = lw._labels_for(s, public=True)   # the names of the chain's columns
        if not all(l in draws for l in labels):
            continue
        if isinstance(s.node.lhs, Var):
            env[s.name] = np.asarray(draws[labels[0]], dtype=np.float64).reshape(S)
            continue
        lv = _grids(s.extents, s.loop_vars)
        idx = [np.broadcast_to(np.asarray(lw.de.ev(ie, lv)), s.shape if s.shape else ()).astype(np.int64).ravel()
               for ie in s.node.lhs.idx]
        top = tuple(int(i.max()) for i in idx)
        cur = env.get(s.name)
        if cur is None:
            cur = np.full(top + (S,), np.nan)
        elif any(c < t for c, t in zip(cur.shape[:-1], top)):
            new = np.full(tuple(max(c, t) for c, t in zip(cur.shape[:-1], top)) + (S,), np.nan)
            new[tuple(slice(0, c) for c in cur.shape[:-1])] = cur
            cur = new
        for k, lab in enumerate(labels):
            cur[tuple(int(i[k]) - 1 for i in idx)] = np.asarray(draws[lab], dtype=np.float64).reshape(S)
        env[s.name] = cur
    de = _DrawEval(lw.data, env, S)
    out: Dict[str, np.ndarray] = {}
    pending = [s for s in lw.stmts if s.kind == "logical" or (s.kind == "param" and s.is_gq)]
    progress = True
    while pending and progress:
        progress = False
        for s in list(pending):
            exprs = [s.node.rhs] if s.kind == "logical" else list(s.node.dist.args)
            if not all(de.known(e, s.loop_vars) for e in exprs):
                continue
            # a statement that reads an array some other pending statement still has to fill must wait
            needs = {v for e in exprs for v in free_vars(e)}
            if any(t is not s and t.name in needs for t in pending):
                continue
            pending.remove(s)
            progress = True
            lv = {k: v[..., None] for k, v in _grids(s.extents, s.loop_vars).items()}
            shape = tuple(s.shape) + (S,)
            try:
                if s.kind == "logical":
                    val = np.broadcast_to(np.asarray(de.ev(s.node.rhs, lv), dtype=np.float64), shape)
                else:
                    val = _sample(rng, s.node.dist.fn, [de.ev(a, lv) for a in s.node.dist.args], shape)
            except LoweringError:
                continue
            lhs = s.node.lhs
            if isinstance(lhs, Var):
                env[lhs.name] = np.asarray(val).reshape(S)
            else:
                glv = _grids(s.extents, s.loop_vars)
                idx = [np.broadcast_to(np.asarray(lw.de.ev(ie, glv)), s.shape if s.shape else ()).astype(np.int64)
                       for ie in lhs.idx]
                top = tuple(int(i.max()) for i in idx)
                cur = env.get(lhs.name)
                if cur is None:
                    cur = np.full(top + (S,), np.nan)
                elif any(c < t for c, t in zip(cur.shape[:-1], top)):
                    new = np.full(tuple(max(c, t) for c, t in zip(cur.shape[:-1], top)) + (S,), np.nan)
                    new[tuple(slice(0, c) for c in cur.shape[:-1])] = cur
                    cur = new
                cur[tuple(i - 1 for i in idx)] = val
                env[lhs.name] = cur
            if s.is_gq or not only_gq:
                out[s.name] = env[s.name]
    return out
