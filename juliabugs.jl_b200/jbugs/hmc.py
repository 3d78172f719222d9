"""On-device batched HMC around the CUDA evaluator (the caller of the hot path).

Reference counterpart: `AbstractMCMC.sample(model, NUTS(0.8), n; n_adapts, ...)` through
`JuliaBUGS/ext/JuliaBUGSAdvancedHMCExt.jl:40-66`, where AdvancedHMC calls `logdensity_and_gradient` once per
leapfrog step for one chain.  Here every chain of the batch advances in lock step on the GPU; only the monitored
parameters of each draw come back to the host.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import plan as P
from .cuda import _check, load
from .model import BUGSModel, getparams


@dataclass
class HMC:
    """static-trajectory HMC: `n_leapfrog` steps per transition, step size and diagonal metric adapted in warm-up."""
    n_leapfrog: int = 16
    step_size: float = 0.01
    jitter: float = 0.1


@dataclass
class NUTS:
    """No-U-Turn sampler, the reference's default (`NUTS(0.8)`): dynamic trajectory length by tree doubling up to
    `max_depth`, multinomial sampling, generalised U-turn criterion; step size adapted to the target acceptance
    statistic 0.8 and a diagonal metric, as for `HMC`."""
    max_depth: int = 10
    step_size: float = 0.01
    jitter: float = 0.1


@dataclass
class Chains:
    names: List[str]
    samples: np.ndarray          # [n_samples, n_monitored, n_chains], constrained space
    accept_rate: float
    step_size: float
    model: Optional[BUGSModel] = None
    n_divergent: int = 0
    tree_depth: float = 0.0     # NUTS: mean tree depth of the last transition
    _gq: Optional[Dict[str, np.ndarray]] = None
    raw: Optional[np.ndarray] = None      # the same draws in unconstrained space (what the device program reads)
    rows: Optional[np.ndarray] = None     # theta index of every monitored row

    def __getitem__(self, name: str) -> np.ndarray:
        """draws of a monitored parameter ('tau', 'beta[2]') or of a generated quantity ('sigma', 'b0', 'outlier[21]'),
        shape [n_samples, n_chains]."""
        if name in self.names:
            return self.samples[:, self.names.index(name), :]
        gq = self.generated_quantities()
        base, _, rest = name.partition("[")
        if base in gq:
            arr = gq[base]
            if rest:
                arr = arr[tuple(int(t) - 1 for t in rest.rstrip("]").split(","))]
            return arr.reshape(self.samples.shape[0], self.samples.shape[2])
        raise KeyError(name)

    def summary(self, names: Optional[Sequence[str]] = None) -> Dict[str, Dict[str, float]]:
        """mean, sd, split-R-hat and effective sample size per quantity (what `summarize(chains)` of MCMCChains shows
        for the reference's runs; `J/test/ext/JuliaBUGSAdvancedHMCExt.jl:56-62` reads mean and std from it)."""
        return {n: dict(mean=float(self[n].mean()), std=float(self[n].std(ddof=1)), rhat=split_rhat(self[n]),
                        ess=effective_sample_size(self[n])) for n in (names or self.names)}

    def generated_quantities(self, seed: int = 0, device: bool = False) -> Dict[str, np.ndarray]:
        """logical nodes outside the target and forward draws of unobserved stochastic leaves, computed from the
        monitored parameters for every draw at once (reference: J/src/model/abstractmcmc.jl:92-110).
        `device=True`: the lowered program on the GPU (jbugs/gq_device.py, one thread per draw) instead of numpy on
        the host; the monitored rows must cover the parameters the quantities read."""
        if self._gq is None and device and any(s.kind == "latent" for s in self.model.lowering.stmts):
            raise ValueError("a marginalised model recovers its summed-out latents from p(z | theta, y) on the host: "
                             "call generated_quantities(device=False)")
        if self._gq is None and device:
            from .gq_device import DeviceGQ
            dq = DeviceGQ(self.model.lowering, device=self.model.evaluation_mode.device, only_gq=True)
            try:
                n, k, c = self.raw.shape
                th = np.ascontiguousarray(self.raw.transpose(1, 0, 2).reshape(k, n * c))
                self._gq = dq.run(th, rows=self.rows, seed=seed)
            finally:
                dq.close()
        if self._gq is None:
            from .gq import generated_quantities
            draws = {n: self.samples[:, k, :].ravel() for k, n in enumerate(self.names)}
            self._gq = generated_quantities(self.model.lowering, draws, seed=seed)
        return self._gq

    def mean(self, name):
        return float(self[name].mean())

    def std(self, name):
        return float(self[name].std())


def split_rhat(x: np.ndarray) -> float:
    """split-R-hat (Gelman et al., BDA3) of draws x[n_samples, n_chains]: every chain is cut in two halves."""
    n = x.shape[0] // 2
    if n < 2:
        return float("nan")
    h = np.concatenate([x[:n], x[n:2 * n]], axis=1)           # [n, 2 * chains]
    w = h.var(axis=0, ddof=1).mean()
    b = n * h.mean(axis=0).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * w + b / n) / w)) if w > 0 else float("nan")


def effective_sample_size(x: np.ndarray) -> float:
    """ESS of draws x[n_samples, n_chains] from the chain-averaged autocorrelation with Geyer's initial positive
    sequence truncation (the estimator Stan and MCMCChains use, without rank normalisation)."""
    n, m = x.shape
    if n < 4:
        return float("nan")
    xc = x - x.mean(axis=0, keepdims=True)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, nfft, axis=0)
    acov = np.fft.irfft(f * np.conj(f), nfft, axis=0)[:n] / n   # biased autocovariance per chain
    w = acov[0].mean() * n / (n - 1)
    var_plus = w * (n - 1) / n + (x.mean(axis=0).var(ddof=1) if m > 1 else 0.0)
    if not var_plus > 0:
        return float("nan")
    rho = 1.0 - (w - acov.mean(axis=1)) / var_plus
    tau, t = -1.0, 0
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        tau += 2.0 * pair
        t += 2
    return float(n * m / max(tau, 1.0 / np.log10(max(n * m, 10))))


def _bind(L):
    c = ctypes
    L.jbugs_hmc_create.argtypes = [c.c_void_p, c.c_int64, c.c_uint64, c.POINTER(c.c_void_p)]
    L.jbugs_hmc_destroy.argtypes = [c.c_void_p]
    L.jbugs_hmc_init.argtypes = [c.c_void_p, c.c_void_p, c.c_double, c.c_double]
    L.jbugs_hmc_run.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int, c.c_void_p, c.POINTER(c.c_double)]
    L.jbugs_hmc_get_state.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.POINTER(c.c_double), c.c_void_p]
    L.jbugs_hmc_run_nuts.argtypes = L.jbugs_hmc_run.argtypes
    L.jbugs_hmc_nuts_stats.argtypes = [c.c_void_p, c.POINTER(c.c_int64), c.POINTER(c.c_double)]
    L.jbugs_hmc_draw_momentum.argtypes = [c.c_void_p, c.c_uint32, c.c_void_p]


def _constrain(tr, lb, ub, y):
    if tr == P.TR_LOG:
        return lb + np.exp(y)
    if tr == P.TR_UPPER:
        return ub - np.exp(y)
    if tr == P.TR_LOGIT:
        return lb + (ub - lb) / (1.0 + np.exp(-y))
    return y


def monitor_table(model: BUGSModel, monitor: Optional[Sequence[str]] = None):
    """(names, name -> (theta slot, bijector code, lb, ub)) of the parameters a run records: every scalar parameter by
    default, else the labels asked for ('tau', 'beta[2]', 'alpha[17]').  Names are the PUBLIC labels: a model whose short
    outer loop was written out by the lowering (Magnesium) keeps `mu[3]`, `theta[3,5]` although its plan calls them
    `mu<0:3>`, `theta<0:3>[5]`."""
    lw = model.lowering
    public = {}
    for st in lw.stmts:
        if st.kind == "param" and not st.is_gq and (st.sid in lw.theta_map or lw.theta_order is not None) and st.size <= 100_000:
            public.update(zip(lw._labels_for(st), lw._labels_for(st, public=True)))
    by_name = {public.get(g.name, g.name): (g.theta_idx, g.transform, g.lb, g.ub) for g in model.plan.globals}
    names = list(by_name) if monitor is None else list(monitor)
    if any(n not in by_name for n in names):
        # elements of plate-level parameters: slot from the theta map, bijector from the plate's local record
        slot = {lab: k for k, lab in enumerate(lw.param_names())}
        tr = {l.name: (l.transform, l.lb, l.ub) for pl in model.plan.plates for l in pl.locals}
        internal = {b: a for a, b in public.items()}
        for n in names:
            if n not in by_name:
                base = internal.get(n, n).partition("[")[0]
                if n not in slot or base not in tr:
                    raise KeyError(f"'{n}' is not a parameter of the model")
                by_name[n] = (slot[n],) + tr[base]
    return names, by_name


def sample(model: BUGSModel, sampler, n_samples: int, n_adapts: int = 500, n_chains: int = 256, seed: int = 1234,
           monitor: Optional[Sequence[str]] = None, init: Optional[np.ndarray] = None, specialize: bool = False) -> Chains:
    """`AbstractMCMC.sample(model, sampler, n_samples; n_adapts, ...)` for `n_chains` lock-step chains.
    `monitor`: parameter labels to record ('tau', 'beta[2]', 'alpha[17]'); default: every scalar parameter."""
    L = load()
    _bind(L)
    cp = model.cuda
    if specialize:
        cp.specialize()   # a kernel for this model's plate shapes (nvcc at run time, cached): worth it for long runs
    names, by_name = monitor_table(model, monitor)
    rows = np.array([by_name[n][0] for n in names], dtype=np.int64)
    th0 = np.ascontiguousarray(getparams(model) if init is None else init, dtype=np.float64)
    th0 = np.where(np.isfinite(th0), th0, 0.0)
    h = ctypes.c_void_p()
    _check(L.jbugs_hmc_create(cp.h, n_chains, seed, ctypes.byref(h)))
    try:
        _check(L.jbugs_hmc_init(h, th0.ctypes.data, sampler.jitter, sampler.step_size))
        acc = ctypes.c_double()
        nuts = isinstance(sampler, NUTS)
        run = L.jbugs_hmc_run_nuts if nuts else L.jbugs_hmc_run
        length = sampler.max_depth if nuts else sampler.n_leapfrog
        if n_adapts > 0:
            _check(run(h, n_adapts, length, 1, None, 0, None, ctypes.byref(acc)))
        out = np.empty((n_samples, len(names), n_chains))
        _check(run(h, n_samples, length, 0, rows.ctypes.data, len(names), out.ctypes.data, ctypes.byref(acc)))
        eps = ctypes.c_double()
        _check(L.jbugs_hmc_get_state(h, None, None, ctypes.byref(eps), None))
        n_div, depth = ctypes.c_int64(), ctypes.c_double()
        _check(L.jbugs_hmc_nuts_stats(h, ctypes.byref(n_div), ctypes.byref(depth)))
    finally:
        L.jbugs_hmc_destroy(h)
    raw = out.copy()
    for k, n in enumerate(names):
        _, tr_k, lb, ub = by_name[n]
        out[:, k, :] = _constrain(tr_k, lb, ub, out[:, k, :])
    return Chains(names, out, float(acc.value), float(eps.value), model, n_divergent=int(n_div.value),
                  tree_depth=float(depth.value), raw=raw, rows=rows)
