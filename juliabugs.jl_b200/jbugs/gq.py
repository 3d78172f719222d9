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
                    iv = np.asarray(self.ev(ie, lv))
                    if self.has_draws(ie):
                        # an index that varies from draw to draw (mu[z[i]] with z a recovered discrete latent)
                        idx.append(np.broadcast_to(iv, iv.shape[:-1] + (self.S,)).astype(np.int64) - 1)
                        continue
                    if iv.ndim:
                        iv = iv[..., 0]  # grids carry a trailing unit axis for the draws
                    idx.append(iv.astype(np.int64) - 1)
            if any(isinstance(i, slice) for i in idx) and any(isinstance(i, np.ndarray) and i.ndim for i in idx):
                raise LoweringError("mixed slice / loop index over a parameter array")
            if any(isinstance(i, np.ndarray) and i.ndim and i.shape[-1:] == (self.S,) and self.has_draws(ie)
                   for i, ie in zip(idx, e.idx)):
                full = [np.asarray(i)[..., None] if not (isinstance(i, np.ndarray) and i.ndim and i.shape[-1:] == (self.S,)
                                                           and self.has_draws(ie)) else i for i, ie in zip(idx, e.idx)]
                return arr[tuple(full) + (np.arange(self.S),)]
            return arr[tuple(idx)]
        if isinstance(e, Index) and e.name not in self.draws and any(
                not isinstance(ie, (Colon, Range)) and self.has_draws(ie) for ie in e.idx):
            arr = np.asarray(self.data[e.name])
            idx = []
            for ie in e.idx:
                if isinstance(ie, (Colon, Range)):
                    raise LoweringError("slice of a data array next to a draw-dependent index")
                iv = np.asarray(self.ev(ie, lv))
                idx.append(np.broadcast_to(iv, np.broadcast_shapes(iv.shape, (self.S,))).astype(np.int64) - 1)
            return arr[tuple(np.broadcast_arrays(*idx))]
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


def _logpdf(fn, y, a):
    """log density of the observation families a summed-out latent's children may have (Appendix A parameterisations)"""
    from scipy import special as sp
    if fn == "dnorm":
        return 0.5 * np.log(a[1]) - 0.5 * a[1] * (y - a[0]) ** 2 - 0.5 * np.log(2.0 * np.pi)
    if fn == "dpois":
        return y * np.log(a[0]) - a[0] - sp.gammaln(y + 1.0)
    if fn == "dbern":
        return np.where(y != 0, np.log(a[0]), np.log1p(-a[0]))
    if fn == "dbin":
        return (sp.gammaln(a[1] + 1.0) - sp.gammaln(y + 1.0) - sp.gammaln(a[1] - y + 1.0)
                + sp.xlogy(y, a[0]) + sp.xlog1py(a[1] - y, -a[0]))
    if fn == "dgamma":
        return a[0] * np.log(a[1]) - sp.gammaln(a[0]) + (a[0] - 1.0) * np.log(y) - a[1] * y
    if fn == "dexp":
        return np.log(a[0]) - a[0] * y
    if fn == "dlnorm":
        return 0.5 * np.log(a[1]) - 0.5 * a[1] * (np.log(y) - a[0]) ** 2 - 0.5 * np.log(2.0 * np.pi) - np.log(y)
    raise LoweringError(f"recovering a summed-out latent below a {fn} observation is not provided")


def _recover_latents(lw: Lowering, de: "_DrawEval", env, S, rng, out):
    """`forward_sample_generated_quantities!!` for a marginalised model (J/src/model/evaluation.jl:354-378 with the
    latent-recovery step the reference's test pins, J/test/model/auto_marginalization.jl:975-1041): every summed-out
    latent is drawn from p(z | theta, y) -- per cell proportional to p(z = k) * prod_children p(y_child | z = k), the
    weights the plate kernel sums -- so that generated quantities depending on it can be forward-sampled."""
    for z in lw.stmts:
        if z.kind != "latent" or z.is_gq:
            continue
        fn = z.node.dist.fn
        lv = {k: v[..., None] for k, v in _grids(z.extents, z.loop_vars).items()}
        shape = tuple(z.shape) + (S,)
        if fn == "dbern":
            states = [0.0, 1.0]
        else:
            wa = z.node.dist.args[0]
            states = [float(k) for k in range(1, int(lw.de.ev(wa.idx[-1].hi, {})) + 1)]
        children = [t for t in lw.stmts if t.kind == "obs" and not t.is_gq and z.sid in lw._latents_of(t)]
        if any(t.loops != z.loops for t in children):
            raise LoweringError(f"'{z.name}': its observations must share its loop nest")
        logw = []
        for val in states:
            env[z.name] = np.full((tuple(hi for _, hi in z.extents) or ()) + (S,), val) if z.loops else np.full(S, val)
            if fn == "dbern":
                p = np.broadcast_to(np.asarray(de.ev(lw._inline(z.node.dist.args[0]), lv), dtype=np.float64), shape)
                lp = np.log(p) if val == 1.0 else np.log1p(-p)
            else:
                wa = z.node.dist.args[0]
                zref = z.node.lhs if isinstance(z.node.lhs, Index) else Var(z.name)
                lp = np.log(np.broadcast_to(np.asarray(de.ev(lw._inline(Index(wa.name, wa.idx[:-1] + (zref,))), lv),
                                                       dtype=np.float64), shape))
            for t in children:
                a = [np.broadcast_to(np.asarray(de.ev(lw._inline(x), lv), dtype=np.float64), shape) for x in t.node.dist.args]
                yl = t.node.lhs if isinstance(t.node.lhs, Index) else Var(t.name)
                y = np.broadcast_to(np.asarray(lw.de.ev(yl, _grids(t.extents, t.loop_vars)), dtype=np.float64)[..., None], shape)
                lp = lp + _logpdf(t.node.dist.fn, y, a)
            logw.append(lp)
        logw = np.stack(logw)                                   # (K,) + dims + (S,)
        logw -= logw.max(axis=0, keepdims=True)
        cdf = np.cumsum(np.exp(logw), axis=0)
        u = rng.random(shape) * cdf[-1]
        pick = (u[None] > cdf).sum(axis=0).clip(0, len(states) - 1)
        zval = np.asarray(states)[pick]
        if z.zobs is not None:                                  # observed cells keep their value (state number + 1 recorded)
            obs = np.asarray(z.zobs)[..., None]
            zval = np.where(obs > 0, np.asarray(states)[np.clip(obs.astype(np.int64) - 1, 0, len(states) - 1)], zval)
        if z.loops:
            cur = np.full(tuple(hi for _, hi in z.extents) + (S,), np.nan)
            cur[tuple(slice(lo - 1, hi) for lo, hi in z.extents)] = zval
            env[z.name] = cur
        else:
            env[z.name] = zval.reshape(S)
        out[z.name] = env[z.name]


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
    if getattr(lw, "marginalize", False):
        _recover_latents(lw, de, env, S, rng, out)
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
    return _restore_split_arrays(lw, out)


def _restore_split_arrays(lw: Lowering, out):
    """arrays the lowering split when it wrote a short outer loop out (`odds.ratio<0:3>`, Lowering._unroll_outer_loops) come
    back under their own name with the dropped axis in place"""
    renamed = getattr(lw, "_renamed", None)
    if not renamed or not any(k in renamed for k in out):
        return out
    groups: Dict[str, list] = {}
    for k in list(out):
        if k in renamed:
            orig, q, v = renamed[k]
            groups.setdefault(orig, []).append((q, v, out.pop(k)))
    for orig, parts in groups.items():
        q = parts[0][0]
        top = max(v for _, v, _ in parts)
        shp = np.broadcast_shapes(*[np.shape(a) for _, _, a in parts])
        full = np.full(shp[:q] + (top,) + shp[q:], np.nan)
        for _, v, a in parts:
            full[(slice(None),) * q + (v - 1,)] = np.broadcast_to(a, shp)
        out[orig] = full
    return out
