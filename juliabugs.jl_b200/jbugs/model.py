"""Host-side mirror of the reference's model API for the hot path.

Same names and argument meaning as the reference (paths relative to /root/reference/JuliaBUGS):

    compile(model_def, data, inits)          src/JuliaBUGS.jl:328-371
    BUGSModel                                src/model/bugsmodel.jl:281-315
    getparams(model)                         src/model/bugsmodel.jl:619-665
    settrans(model, flag)                    src/model/bugsmodel.jl:697-707
    set_evaluation_mode(model, mode)         src/model/bugsmodel.jl:736-801
    condition / decondition                  src/model/abstractppl.jl:700-799 (new plan, as the reference drops its
                                             generated function); set_observed_values = data re-upload (:872-888)
    UseGraph / UseGeneratedLogDensityFunction / UseAutoMarginalization / UseCUDABatch
                                             (bugsmodel.jl:248-252 + the new mode)
    LogDensityProblems.logdensity / logdensity_and_gradient / dimension / capabilities
                                             src/model/logdensityproblems.jl:19-55,168-174,219-232

Julia is not installed in this image, so the host side above the C ABI is written in Python
(the Julia glue that a maintainer would add is shipped, unexecuted, in `julia/`).  Only the new
`UseCUDABatch` mode is implemented here: the CPU graph walk and the generated Julia function are
what this path replaces, and there is deliberately no CPU fallback.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import plan as P
from .cuda import CudaPlan, JBugsCudaError
from .lowering import Lowering, LoweringError


AUTO_SPECIALIZE_CELLS = 200_000


class EvaluationMode:
    pass


class UseGraph(EvaluationMode):
    """The reference's default node-by-node CPU walk -- not provided by this package."""


class UseGeneratedLogDensityFunction(EvaluationMode):
    """The reference's generated Julia function -- not provided by this package."""


@dataclass
class UseCUDABatch(EvaluationMode):
    """The new evaluation mode: plate plan + libjbugs_cuda on one B200.  `specialize`: compile a kernel for the exact
    shape of every plate that has no in-tree specialisation, expression tapes included (NVRTC in-process, cubins cached
    per user; the reference generates and compiles Julia code per model at the same point, bugsmodel.jl:736-801).
    True: when the device plan is built, blocking.  "auto" (default): for plates of at least `AUTO_SPECIALIZE_CELLS`
    observations, on a worker thread -- the interpreter kernels answer until the compiled ones are ready, and a box
    without libnvrtc simply keeps the interpreter.  False: never."""
    device: int = 0
    specialize: object = "auto"


@dataclass
class UseAutoMarginalization(UseCUDABatch):
    """`UseAutoMarginalization()` (bugsmodel.jl:803-876, evaluation.jl:739-961) on the CUDA path: unobserved finite
    discrete nodes (`z[i] ~ dcat(w[1:K])` selecting the parameters of `y[i]`) are summed out inside the plate kernel,
    theta holds the continuous parameters only.  The plate class covers one latent per observation cell -- `dcat` or
    `dbern`, possibly partially observed, with any number of observed children (mixtures, Cervix); chained latents
    (HMMs) raise `LoweringError`."""


class LogDensityOrder:
    def __init__(self, k):
        self.k = k

    def __eq__(self, o):
        return isinstance(o, LogDensityOrder) and o.k == self.k

    def __repr__(self):
        return f"LogDensityOrder{{{self.k}}}()"


class BUGSModel:
    def __init__(self, model_def, data, inits=None, transformed=True, theta_order=None, evaluation_mode=None):
        self.model_def = model_def
        self.data = dict(data)
        self.inits = dict(inits or {})
        self.transformed = transformed
        self.theta_order = theta_order
        self.evaluation_mode: EvaluationMode = evaluation_mode or UseCUDABatch()
        self.marginalize = isinstance(self.evaluation_mode, UseAutoMarginalization)
        self._lower()
        self._cuda: Optional[CudaPlan] = None

    def _lower(self):
        self.lowering = Lowering(self.model_def, self.data, transformed=self.transformed, theta_order=self.theta_order,
                                 marginalize=self.marginalize)
        self.plan: P.Plan = self.lowering.plan

    # -- device state ---------------------------------------------------------------------------
    @property
    def cuda(self) -> CudaPlan:
        if not isinstance(self.evaluation_mode, UseCUDABatch):
            raise JBugsCudaError(f"{type(self.evaluation_mode).__name__} is not implemented by this package; "
                                 "use set_evaluation_mode(model, UseCUDABatch())")
        if self._cuda is None:
            self._cuda = CudaPlan(self.plan, device=self.evaluation_mode.device)
            want = self.evaluation_mode.specialize
            if want is True:
                self._cuda.specialize()
            elif want == "auto" and self._interpreted_cells() >= AUTO_SPECIALIZE_CELLS:
                try:
                    self._cuda.specialize_async()
                except JBugsCudaError:
                    pass   # no libnvrtc on this box: the interpreter instantiations keep running
        return self._cuda

    def _interpreted_cells(self) -> int:
        """observation cells of the plates that run an interpreter instantiation (no in-tree specialised kernel)"""
        cells = 0
        for k, line in enumerate(self._cuda.describe().splitlines()[1:]):
            if k < len(self.plan.plates) and ("kernel=dynamic " in line or "kernel=expression_tape " in line):
                cells += self.plan.plates[k].N * self.plan.plates[k].T
        return cells

    def invalidate(self):
        """drop the device plan (what condition/fix/decondition do to the generated function,
        src/model/abstractppl.jl:796-799)."""
        if self._cuda is not None:
            self._cuda.close()
            self._cuda = None

    # -- observed data update without re-planning (abstractppl.jl:872-888) ----------------------------
    def set_observed_values(self, **values):
        self.data.update(values)
        old = self.plan
        self._lower()
        if [s.name for s in old.data] != [s.name for s in self.plan.data] or old.to_bytes() != self.plan.to_bytes():
            self.invalidate()  # structure changed: needs a new plan
        elif self._cuda is not None:
            for k, slot in enumerate(self.plan.data):
                self._cuda.upload(k, slot.values)
            self._cuda.plan = self.plan
        return self


def compile(model_def, data, inits=None, theta_order=None, evaluation_mode=None) -> BUGSModel:  # noqa: A001 - reference name
    """`evaluation_mode`: the reference compiles first and switches modes afterwards (`set_evaluation_mode`); a model
    whose discrete latents select parameters (`mu[z[i]]`) has no plan outside `UseAutoMarginalization()` -- a discrete
    theta entry cannot steer a gather on the device -- so that mode can be named at compile time."""
    return BUGSModel(model_def, data, inits, transformed=True, theta_order=theta_order, evaluation_mode=evaluation_mode)


def condition(model: BUGSModel, values=None, **kw) -> BUGSModel:
    """`AbstractPPL.condition(model, (; var = value))` (src/model/abstractppl.jl:700-799): the named variables become
    observations at the given values.  A new model with a new plan -- the reference likewise drops its generated
    function (`log_density_computation_function => nothing`, :796-797) and rebuilds it on demand."""
    vals = dict(values or {}, **kw)
    data = dict(model.data)
    for k, v in vals.items():
        if k in data and isinstance(data[k], np.ndarray) and np.ndim(v) == 0:
            raise ValueError(f"'{k}' is an array: pass the whole array (NaN = still unobserved)")
        data[k] = v
    m = BUGSModel(model.model_def, data, {k: v for k, v in model.inits.items() if k not in vals},
                  transformed=model.transformed, theta_order=None)
    m.evaluation_mode = model.evaluation_mode
    if model.marginalize:
        m.marginalize = True
        m._lower()
    m._conditioned = dict(getattr(model, "_conditioned", {}), **{k: model.data.get(k) for k in vals})
    return m


def decondition(model: BUGSModel, *names) -> BUGSModel:
    """undo `condition` for the named variables (all of them when none is named): they are parameters again."""
    was = dict(getattr(model, "_conditioned", {}))
    names = list(names) or list(was)
    data = dict(model.data)
    for k in names:
        if k not in was:
            raise KeyError(f"'{k}' was not conditioned on")
        prev = was.pop(k)
        if prev is None:
            data.pop(k, None)
        else:
            data[k] = prev
    m = BUGSModel(model.model_def, data, model.inits, transformed=model.transformed, theta_order=None)
    m.evaluation_mode = model.evaluation_mode
    if model.marginalize:
        m.marginalize = True
        m._lower()
    m._conditioned = was
    return m


def settrans(model: BUGSModel, flag: bool = True) -> BUGSModel:
    if flag == model.transformed:
        return model
    m = copy.copy(model)
    m.transformed = flag
    m._cuda = None
    m._lower()
    return m


def set_evaluation_mode(model: BUGSModel, mode: EvaluationMode) -> BUGSModel:
    if not isinstance(mode, UseCUDABatch):
        raise NotImplementedError(
            f"{type(mode).__name__}: only UseCUDABatch is provided (the CPU modes are the reference's own)")
    m = copy.copy(model)
    m.evaluation_mode = mode
    m._cuda = None
    if isinstance(mode, UseAutoMarginalization) != model.marginalize:
        m.marginalize = isinstance(mode, UseAutoMarginalization)
        m._lower()
    return m


def _link(tr, lb, ub, x):
    x = np.asarray(x, dtype=np.float64)
    if tr == P.TR_LOG:
        return np.log(x - lb)
    if tr == P.TR_UPPER:
        return np.log(ub - x)
    if tr == P.TR_LOGIT:
        u = (x - lb) / (ub - lb)
        return np.log(u) - np.log1p(-u)
    return x


def getparams(model: BUGSModel) -> np.ndarray:
    """Initial theta in the model's parameter order: inits pushed through the bijectors
    (bugsmodel.jl:619-665).  Parameters without an init get theta = 0 (the reference draws them
    from the prior with its own RNG, bugsmodel.jl:441-447 -- not reproducible here)."""
    lw = model.lowering
    theta = np.zeros(lw.D)
    fams = {}
    for g in model.plan.globals:
        fams[g.name] = (g.transform, g.lb, g.ub)
    for pl in model.plan.plates:
        for l in pl.locals:
            fams[l.name] = (l.transform, l.lb, l.ub)
    for s in lw.stmts:
        if s.sid not in lw.theta_map and lw.theta_order is None:
            continue
        if s.kind != "param" or s.is_gq or s.name not in fams:
            continue
        idx = lw.theta_index(s)
        if s.name not in model.inits:
            continue
        tr, lb, ub = fams[s.name]
        val = np.asarray(model.inits[s.name], dtype=np.float64)
        if idx.ndim == 0:
            theta[int(idx)] = float(_link(tr, lb, ub, val))
        else:
            lv = {nm: np.arange(lo, hi + 1, dtype=np.int64).reshape([-1 if d == k else 1 for d in range(len(s.shape))])
                  for k, (nm, (lo, hi)) in enumerate(zip(s.loop_vars, s.extents))}
            sel = tuple(np.broadcast_to(np.asarray(lw.de.ev(ie, lv)), s.shape).astype(np.int64) - 1
                        for ie in s.node.lhs.idx)
            vals = val[sel]
            ok = ~np.isnan(vals)
            theta[idx[ok]] = _link(tr, lb, ub, vals[ok])
    return theta


def evaluate_with_values(model: BUGSModel, theta, temperature: float = 1.0):
    """`evaluate_with_values!!(model, theta; temperature)` (src/model/evaluation.jl:217-275) for one theta or a (D, C)
    batch: dict(logprior, loglikelihood, tempered_logjoint) -- floats for one theta, arrays over chains for a batch.
    (The evaluation environment the reference also returns is host-side state: `Chains.generated_quantities`.)"""
    th = np.asarray(theta, dtype=np.float64)
    single = th.ndim == 1
    lpr, ll, lp, _, _ = model.cuda.eval_tempered(th.reshape(-1, 1) if single else th, temperature, want_grad=False)
    pick = (lambda a: float(a[0])) if single else (lambda a: a)
    return {"logprior": pick(lpr), "loglikelihood": pick(ll), "tempered_logjoint": pick(lp)}


class LogDensityProblems:
    """Namespace mirroring LogDensityProblems.jl's generic functions for `BUGSModel`."""

    @staticmethod
    def dimension(model: BUGSModel) -> int:
        return model.plan.D

    @staticmethod
    def capabilities(model) -> LogDensityOrder:
        # the CUDA mode carries its own reverse pass: gradient without an AD backend
        return LogDensityOrder(1)

    @staticmethod
    def logdensity(model: BUGSModel, x) -> float:
        x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
        logp, _, _ = model.cuda.eval_host(x, layout="param_fast", want_grad=False)
        return float(logp[0])

    @staticmethod
    def logdensity_and_gradient(model: BUGSModel, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
        logp, grad, _ = model.cuda.eval_host(x, layout="param_fast", want_grad=True)
        return float(logp[0]), grad[:, 0].copy()

    # ---- the new batched entry points (the reference has none: one theta per call) -------------------
    @staticmethod
    def logdensity_batch(model: BUGSModel, theta, layout="param_fast"):
        """theta: (D, C); `param_fast` == Julia Matrix{Float64}(D, C) (one column per chain)."""
        logp, _, status = model.cuda.eval_host(theta, layout=layout, want_grad=False)
        return logp

    @staticmethod
    def logdensity_and_gradient_batch(model: BUGSModel, theta, layout="param_fast"):
        logp, grad, status = model.cuda.eval_host(theta, layout=layout, want_grad=True)
        return logp, grad
