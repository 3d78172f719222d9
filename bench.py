#!/usr/bin/env python
"""bench.py -- logdensity+gradient obs-evals/s of the batched B200 evaluator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rats|seeds|dense|pumps|epil] [--no-secondary] ...

A "step" is one batched `logdensity_and_gradient` (jbugs_eval_batch) over the whole synthetic data
plate for every chain of the batch.  Default workload = BASELINE.json configs[1]: Rats scaled to
10^6 rats x 5 timepoints, 4096 chains, fp64, one B200.  With N > 1 (torchrun) chains shard across
GPUs with no data-path collective (weak scaling: 4096 chains per GPU).

  value      obs-evals/s with theta/grad resident in HBM (CUDA events around exactly K steps)
  e2e        same metric through the C ABI with pinned HOST buffers (H2D of theta, D2H of grad/logp
             inside the timed region), every chain of the batch, in slabs; both host layouts
  roofline   algorithmic bytes (2*8*C*D) / MEAN plate-kernel time over the K timed steps (the
             library's own CUDA events, one set per step) vs MEASURED_PEAKS.json hbm_gbs
  sustained  the same step repeated for >= 3 s right after the timed region (per-step p50 / p95)
  secondary  the other BASELINE configurations (3: Seeds 10^7 obs, observation-sharded + NCCL
             all-reduce when N > 1; 4: dense logistic 2^22 x 512 x 1024 on the FP64 tensor cores,
             rows sharded when N > 1; 5: Pumps / Epil ensembles, 8192 chains per GPU), each with its
             own roofline, measured in the same process after the primary workload
  cpu_baseline  the C restatement of the reference evaluator (oracle/) on the host cores, bounded sample

`--impl reference` times that CPU restatement alone (Julia is not installed here or on the GPU box,
so the reference itself cannot run; kind = "port").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "juliabugs.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "logdensity+gradient obs-evals/s"
UNIT = "obs-evals/s"
DEFAULT_CHAINS = {"rats": 4096, "seeds": 256, "pumps": 8192, "epil": 8192, "dense": 1024, "rats_variant": 1024, "mixture": 1024}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------- workloads
def make_workload(name, groups, dev=None):
    """returns (model_text, data, theta0 (D,), description, device_slots).  `device_slots` maps a data-slot name
    prefix to a device tensor that is uploaded straight from HBM (the 17 GB design matrix of the dense workload is
    generated on the GPU; its host entry is an untouched placeholder of the right shape)."""
    from jbugs import examples as ex
    from helpers import synth_rats, synth_seeds
    if name == "rats":
        N = groups or 1_000_000
        data, alpha, beta = synth_rats(N)
        D = 2 * N + 5
        th0 = np.empty(D)
        th0[:5] = [np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)]  # beta.tau, beta.c, alpha.tau, alpha.c, tau.c
        th0[5::2] = beta
        th0[6::2] = alpha
        return ex.RATS, data, th0, f"rats N={N} T=5", {}
    if name == "seeds":
        N = groups or 10_000_000
        data, b = synth_seeds(N)
        th0 = np.empty(N + 5)
        th0[:5] = [np.log(1 / 0.09), -0.84, 1.36, 0.09, -0.55]  # tau, alpha12, alpha2, alpha1, alpha0
        th0[5:] = b
        return ex.SEEDS, data, th0, f"seeds N={N}", {}
    if name == "dense":
        # BASELINE.json config 4: dense hierarchical logistic regression, X (N x P) fp64 column-major, beta over chains
        import torch
        N, Pn = groups or (1 << 22), 512
        rng = np.random.default_rng(1234)
        bt = rng.normal(0.0, 0.1, Pn)
        th0 = np.concatenate([[np.log(100.0), 0.0], bt])  # tau.b, mu.b, beta[1..P]
        desc = f"dense logistic N={N} P={Pn}"
        if dev is None:  # CPU arm: host arrays
            gen = torch.Generator().manual_seed(1234)
            X = torch.randn(Pn * N, dtype=torch.float64, generator=gen).numpy().reshape((Pn, N)).T  # F-contiguous N x P
            y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ bt)))).astype(np.float64)
            return ex.DENSE_LOGISTIC, {"N": N, "P": Pn, "X": X, "y": y}, th0, desc, {}
        gen = torch.Generator(device=dev).manual_seed(1234)
        Xd = torch.empty((Pn, N), dtype=torch.float64, device=dev)  # [P][N] row-major == N x P column-major
        for j0 in range(0, Pn, 64):
            Xd[j0:j0 + 64] = torch.randn((min(64, Pn - j0), N), dtype=torch.float64, device=dev, generator=gen)
        eta = torch.from_numpy(bt).to(dev) @ Xd
        yd = (torch.rand(N, dtype=torch.float64, device=dev, generator=gen) < torch.sigmoid(eta)).to(torch.float64)
        y = yd.cpu().numpy()
        X = np.empty((Pn, N), dtype=np.float64).T  # placeholder (never touched): shape and layout only
        return ex.DENSE_LOGISTIC, {"N": N, "P": Pn, "X": X, "y": y}, th0, desc, {"dense|X": Xd}
    if name == "rats_variant":
        # a model WITHOUT an in-tree kernel (the Rats plate with one more regression term on a group-level covariate):
        # it runs the interpreter until run-time specialisation (NVRTC, tiered) has compiled its kernel
        N = groups or 1_000_000
        data, alpha, beta = synth_rats(N)
        data["z"] = np.random.default_rng(7).standard_normal(N)
        text = ex.RATS.replace("mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar)", "mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar) + gam * z[i]")
        text = text.replace("alpha.c ~ dnorm(0.0, 1.0E-6)", "alpha.c ~ dnorm(0.0, 1.0E-6)\n    gam ~ dnorm(0.0, 0.01)")
        import jbugs
        names = jbugs.lower(text, {k: (v[:4] if isinstance(v, np.ndarray) and v.shape[:1] == (N,) else v) for k, v in dict(data, N=4).items()})[1].param_names()
        ng = sum(1 for n in names if "[" not in n)
        th0 = np.empty(2 * N + ng)
        start = {"beta.tau": np.log(4.0), "beta.c": 6.2, "alpha.tau": np.log(1 / 196.0), "alpha.c": 242.0, "tau.c": np.log(1 / 36.0), "gam": 0.0}
        th0[:ng] = [start[n] for n in names[:ng]]
        first = names[ng].split("[")[0]
        th0[ng::2] = beta if first == "beta" else alpha
        th0[ng + 1::2] = alpha if first == "beta" else beta
        return text, data, th0, f"rats + one regression term (no in-tree kernel; specialised at run time) N={N} T=5", {}
    if name == "mixture":
        # two-component normal mixture, the discrete latents summed out inside an expression-tape kernel (compiled at run time)
        N = groups or 2_000_000
        rng = np.random.default_rng(2)
        z = rng.random(N) < 0.3
        y = np.where(z, rng.normal(-2, 1.0, N), rng.normal(2, 0.7, N))
        text = """model {
  w[1] <- 0.3
  w[2] <- 0.7
  mu[1] ~ dnorm(-2, 0.04)
  mu[2] ~ dnorm(2, 0.04)
  sigma[1] ~ dexp(1)
  sigma[2] ~ dexp(1)
  for (k in 1:2) { tau[k] <- 1 / (sigma[k] * sigma[k]) }
  for (i in 1:N) {
    z[i] ~ dcat(w[1:2])
    y[i] ~ dnorm(mu[z[i]], tau[z[i]])
  }
}"""
        import jbugs
        names = jbugs.lower(text, {"N": 4, "y": y[:4]}, marginalize=True)[1].param_names()
        start = {"sigma[1]": 0.0, "sigma[2]": np.log(0.7), "mu[2]": 2.0, "mu[1]": -2.0}
        return text, {"N": N, "y": y}, np.array([start[n] for n in names]), f"normal mixture, latents summed out (tape kernel compiled at run time) N={N}", {}
    if name in ("pumps", "epil"):
        # BASELINE.json config 5: the shipped fixtures, a large ensemble of independent chains (chain-sharded)
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "examples_vol1.json")))[name]
        data = {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in fx["data"].items()}
        inits = {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in fx["inits"].items()}
        import jbugs
        m = jbugs.compile(ex.PUMPS if name == "pumps" else ex.EPIL, data, inits)
        th0 = jbugs.getparams(m)
        if name == "pumps":
            th0[2:] = np.log(np.maximum(data["x"], 0.5) / data["t"])  # theta_i near x_i / t_i
        return (ex.PUMPS if name == "pumps" else ex.EPIL), data, th0, f"{name} (shipped fixture)", {}
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """clocks / throttle reasons sampled between the wall-clock instants t0 and t1"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        sm, mx, reasons = [], [], set()
        for t, line in list(self.samples):
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(np.min(sm)), "sm_max_mhz": float(np.max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ----------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(plan, th0, n_obs, budget_s=15.0, seed=7, cores=None):
    """the C restatement of the reference evaluator on the host cores, one chain per thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from c_oracle import COracle, native_lib
    co = COracle(plan, native=True)   # compiled here with -O3 -march=native: the strongest build of the port
    # every host core this process may run on; not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    if cores is None:
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or co.max_threads()
    rng = np.random.default_rng(seed)

    def run(C):
        theta = np.ascontiguousarray(th0[:, None] + 0.05 * rng.standard_normal((th0.size, C)))
        t = time.perf_counter()
        co.eval_batch(theta, layout="param_fast", want_grad=True, nthreads=cores)
        return time.perf_counter() - t

    t1 = run(cores)  # one chain per thread
    reps = int(max(1, min(64, budget_s / max(t1, 1e-3))))
    C = cores * reps
    if reps > 1:
        t1 = run(C)
    else:
        C = cores
    return {"value": n_obs * C / t1, "unit": UNIT, "cores": cores, "kind": "port", "seconds": t1,
            "sample": f"{C} chains x {n_obs} obs, one chain per thread, {t1:.2f} s",
            "note": "C restatement of the reference's generated evaluator (oracle/jbugs_oracle.c, " + native_lib()[1] +
                    ", analytic gradient) -- NOT JuliaBUGS + Mooncake itself (Julia is not installed)"}


# ----------------------------------------------------------------------------------- one workload on the GPU(s)
class Ctx:
    pass


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pinned(nbytes, write_combined=False):
    """(torch tensor view, raw address) of library-allocated pinned host memory (cudaHostAlloc on the NUMA node this
    process was bound to)."""
    import torch
    from jbugs import cuda
    ptr = cuda.host_alloc(nbytes, write_combined=write_combined)
    arr = np.ctypeslib.as_array((ctypes.c_double * (nbytes // 8)).from_address(ptr))
    return torch.from_numpy(arr), ptr


def measure(name, ctx, args, K, W, C, primary):
    """build the workload, time exactly K steps (CUDA events on the launch stream, max over ranks), return the
    fields of a bench line for it"""
    import torch
    import jbugs
    from jbugs import cuda as jcuda
    dist = ctx.dist
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    shard = (args.shard if primary and args.shard else {"seeds": "obs", "dense": "obs"}.get(name, "chains"))
    model_text, data, th0, desc, dev_slots = make_workload(name, args.groups if primary else 0, dev)
    t0 = time.time()
    model = jbugs.compile(model_text, data, evaluation_mode=(jbugs.UseAutoMarginalization(device=ctx.local_rank, specialize=False)
                                                             if name == "mixture" else jbugs.UseCUDABatch(device=ctx.local_rank, specialize=False)))
    plan = model.plan
    n_obs = plan.n_obs()
    D = plan.D
    log(f"[bench] lowered {desc}: D={D} obs={n_obs} in {time.time() - t0:.1f}s")
    obs_sharded = shard == "obs" and world > 1
    cp = jcuda.CudaPlan(plan, device=ctx.local_rank, upload=False)
    model._cuda = cp
    for k, slot in enumerate(plan.data):
        src = next((t for pre, t in dev_slots.items() if slot.name.startswith(pre)), None)
        if src is None:
            cp.upload(k, slot.values)
        else:  # device-resident source: device-to-device copy inside the library
            jcuda._check(cp.L.jbugs_plan_upload_data(cp.h, k, src.data_ptr(), src.numel() * 8, 0))
    dev_slots.clear()
    del data
    if primary:
        if args.dynamic:
            cp.set_option("dynamic", 1)
        if args.tile:
            cp.set_option("tile", args.tile)
        if args.tma >= 0:
            cp.set_option("tma", args.tma)
        for kv in filter(None, args.opt.split(",")):
            k, v = kv.split("=")
            cp.set_option(k, int(v))
    jit = None
    if name in ("rats_variant", "mixture"):
        # run-time specialisation, tiered: the call returns with the quick variant, the best one follows from a worker thread
        t_j = time.time()
        cp.specialize("none")
        t_quick = time.time() - t_j
        jit = {"quick_variant_s": t_quick, "nvrtc_s": cp.last_compile_seconds()}
    local_groups = [pl.N for pl in plan.plates]
    dense_rows = plan.dense[0].N if plan.dense else 0
    if obs_sharded:
        from jbugs.dist import shard_bounds, shard_observations
        shard_observations(model, rank, world)
        local_groups = [(b - a) if pl.has_obs else pl.N
                        for (a, b), pl in ((shard_bounds(pl.N, world, rank), pl) for pl in plan.plates)]
        if plan.dense:
            a, b = shard_bounds(plan.dense[0].N, world, rank)
            dense_rows = b - a
    par = ((f"data plate sharded by observation x{world} + one ncclAllReduce of C x (2+G"
            f"{'+P' if plan.dense else ''}) doubles per evaluation") if shard == "obs"
           else f"chains sharded x{world}, no collective")
    # theta rows this rank really reads / writes: the globals + the locals of its groups
    D_local = len(plan.globals) + sum(
        ng * sum(pl.T if l.level == 1 else 1 for l in pl.locals) for ng, pl in zip(local_groups, plan.plates))
    ld = (C + 15) // 16 * 16
    theta = torch.empty((D, ld), dtype=torch.float64, device=dev)
    grad = torch.empty((D, ld), dtype=torch.float64, device=dev)
    logp = torch.empty(ld, dtype=torch.float64, device=dev)
    status = torch.empty(ld, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + (0 if obs_sharded else rank))  # observation sharding: every rank holds the SAME chains
    th0_d = torch.from_numpy(th0).to(dev)
    rows = max(1, (1 << 28) // (8 * ld))
    for r0 in range(0, D, rows):
        r1 = min(D, r0 + rows)
        theta[r0:r1] = th0_d[r0:r1, None] + 0.05 * torch.randn((r1 - r0, ld), dtype=torch.float64, device=dev, generator=g)
    grad.zero_()
    torch.cuda.synchronize()
    stream = ctx.stream.cuda_stream
    want_grad = not (primary and args.logdensity_only)
    work_gb = 2 * 8 * C * D_local / 1e9
    config = {"workload": f"{desc} x {C} chains{'' if shard == 'obs' else '/GPU'}, fp64, theta+grad resident in HBM (chain-fast layout)",
              "chains_per_gpu": C, "D": D, "n_obs": n_obs, "parallelism": par,
              "l2": ("inputs (X = %.1f GB, R = %.1f GB per GPU) exceed the 126 MB L2; no flush needed"
                     % (8.0 * n_obs * (D - 2) / 1e9 / (world if shard == "obs" else 1), 8.0 * n_obs * C / 1e9 / (world if shard == "obs" else 1)))
              if name == "dense" else "inputs (theta+grad = %.1f GB/GPU) exceed the 126 MB L2; no flush needed" % work_gb}
    if not want_grad:
        config["evaluation"] = "logdensity only (no gradient): theta is read, nothing but logp is written"

    def step(flags=1):
        cp.eval_raw(theta.data_ptr(), ld, C, 0, logp.data_ptr(), grad.data_ptr(), status.data_ptr(), want_grad, flags, stream)

    step(0)  # first call also builds the one-off data constants
    if jit is not None:
        # wait for the best variant before anything is timed (a synchronous evaluation is the adoption point)
        t_j = time.time()
        while "quick_variant" in cp.describe() and time.time() - t_j < 90:
            time.sleep(0.25)
            step(0)
        jit["best_variant_after_s"] = time.time() - t_j
        jit["kernel"] = cp.describe().splitlines()[1].split("kernel=")[1].split()[0]
        config["run_time_specialisation"] = jit
    log("[bench] " + cp.describe().replace("\n", " | "))
    for _ in range(W):
        step(0)
    launches0 = cp.launch_count()
    step(0)
    launches_per_step = cp.launch_count() - launches0
    # working set could sit in the 126 MB L2: flush between timed steps (not the dense workload: its X and R matrices are GBs)
    small = 2 * 8 * C * D_local < (1 << 28) and name != "dense"
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tw0 = time.time()
    Ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    if small:
        # event pair per step, a 256 MB write (> L2) between steps and outside the pair.  The library's own phase events
        # are NOT recorded here: four event records are a visible share of a 30 us step (timed separately below)
        flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
        evs = [(Ev(), Ev()) for _ in range(K)]
        for a, b in evs:
            flush.zero_()
            a.record()
            step(1)
            b.record()
        torch.cuda.synchronize()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        ms_total = float(sum(step_ms))
        del flush
        config["l2"] = "working set fits in L2: a 256 MB buffer is rewritten before every timed step (outside the event pair)"
        if launches_per_step == 1:
            # one-launch evaluation (fused_small_kernel): the step IS the kernel, the event pair above brackets exactly it
            plate_ms, total_ms = np.array(step_ms), np.array(step_ms)
        else:
            cp.set_option("timing", 2)
            for _ in range(min(K, 20)):
                step(0)
            plate_ms, total_ms = cp.timing_history()
            cp.set_option("timing", 0)
    else:
        # one boundary event per step (K + 1 events) -> per-step times; the library records its own four events per step
        # into a ring ("timing" = 2), so the plate-kernel time of EVERY timed step is known (mean and max reported)
        cp.set_option("timing", 2)
        evs = [Ev() for _ in range(K + 1)]
        evs[0].record()
        for k in range(K):
            step(1)
            evs[k + 1].record()
        torch.cuda.synchronize()
        ms_total = evs[0].elapsed_time(evs[K])
        step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(K)]
        plate_ms, total_ms = cp.timing_history()
        plate_ms, total_ms = plate_ms[-K:], total_ms[-K:]
        cp.set_option("timing", 0)
    if world > 1:
        dist.barrier()
    tw1 = time.time()
    clocks = ctx.sampler.window(tw0, tw1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = n_obs * C * (1 if shard == "obs" else world) / (ms_per_step * 1e-3)
    lp_host = logp[:C].cpu().numpy()
    st_host = status[:C].cpu().numpy()
    if not (np.all(np.isfinite(lp_host)) and np.all(st_host == 0)):
        raise SystemExit(f"bench ({name}) produced non-finite log densities: invalid run")

    # ---- sustained region (primary workload): the same step for >= 3 s, per-step boundary events
    sustained = None
    if primary and args.sustain_s > 0 and not small:
        n_s = int(min(1000, max(K, np.ceil(args.sustain_s * 1e3 / ms_per_step))))
        cp.set_option("timing", 2)
        evs = [Ev() for _ in range(n_s + 1)]
        ts0 = time.time()
        evs[0].record()
        for k in range(n_s):
            step(1)
            evs[k + 1].record()
        torch.cuda.synchronize()
        ts1 = time.time()
        sm = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(n_s)])
        pk, _ = cp.timing_history()
        pk = pk[-n_s:]
        cp.set_option("timing", 0)
        sustained = {"steps": n_s, "seconds": float(sm.sum() * 1e-3), "ms_per_step_mean": float(sm.mean()),
                     "ms_per_step_p50": float(np.percentile(sm, 50)), "ms_per_step_p95": float(np.percentile(sm, 95)),
                     "kernel_ms_mean": float(pk.mean()), "kernel_ms_max": float(pk.max()),
                     "value": n_obs * C * (1 if shard == "obs" else world) / (float(sm.mean()) * 1e-3),
                     "clocks": ctx.sampler.window(ts0, ts1)}

    # ---- roofline of the dominant kernel: mean over the K timed steps of the library's per-step events
    peak, which = hbm_peak()
    alg_bytes = (2.0 if want_grad else 1.0) * 8.0 * C * D_local
    pk_ms = float(np.mean(plate_ms))
    achieved = alg_bytes / (pk_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not (primary and args.groups):
        try:
            t = json.load(open(tpath)).get(name)
            # the ncu --set full capture replays the kernel ~40x and must save/restore what it writes, so it was
            # taken at fewer chains than fit beside a 131 GB working set; DRAM bytes are linear in the chain count
            traffic = t["dram_bytes_per_launch"] * C / t["chains"] * (D_local / D)
            traffic_note = f"dram__bytes_read+write from {t['source']} (captured at {t['chains']} chains, scaled to {C})"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "fused_small_kernel (prologue + plate tile + finalize, one launch)" if (small and launches_per_step == 1) else "plate_kernel",
                "kernel_ms": pk_ms, "kernel_ms_max": float(np.max(plate_ms)), "kernel_ms_samples": int(len(plate_ms)),
                "eval_ms": float(np.mean(total_ms)), "algorithmic_bytes": alg_bytes, "peak_source": which}
    if name == "mixture":
        # D = 4 parameters per chain live in registers and the observations are shared by all chains: the HBM figure above
        # is kept for the contract's shape but bounds nothing here
        roofline["note"] = ("not HBM-bound: the kernel's time is FP64 / special-function arithmetic (two log-density branches, a "
                            "log-sum-exp and their reverse sweeps per observation per chain); no instruction count of the "
                            "run-time-compiled kernel was captured, so no FP64-pipe fraction is claimed -- read ms_per_step / value")
    elif small:
        roofline["note"] = ("latency-bound ensemble of a small model (working set in L2, one launch per evaluation): the HBM "
                            "fraction is reported for the contract's shape; the figure of merit is ms_per_step (DESIGN section 3)")
    fpath = os.path.join(ROOT, "profiles", "fp64_counts.json")
    if name in ("seeds",) and os.path.exists(fpath):
        # the logit plate kernel is FP64-pipe bound, not HBM bound: report both.  FP64 instructions per
        # observation-evaluation are counted in the SASS of the timed kernel (scripts/count_fp64_sass.py)
        try:
            fc = json.load(open(fpath))[name]
            fpeak = jcuda.fp64_peak(ctx.local_rank)  # DFMA / s, measured now
            n_local_obs = sum(ng * pl.T for ng, pl in zip(local_groups, plan.plates) if pl.has_obs)
            instr = fc["fp64_per_obs"] * n_local_obs * C
            ach = instr / (pk_ms * 1e-3)
            roofline = {"bound": "fp64", "achieved": ach / 1e12, "peak": fpeak / 1e12, "unit": "T FP64-instr/s",
                        "frac": ach / fpeak, "traffic": traffic, "traffic_note": traffic_note, "kernel": "plate_kernel",
                        "kernel_ms": pk_ms, "kernel_ms_max": float(np.max(plate_ms)), "kernel_ms_samples": int(len(plate_ms)),
                        "eval_ms": float(np.mean(total_ms)),
                        "fp64_instr_per_obs_eval": fc["fp64_per_obs"], "fp64_count_source": fc.get("source"),
                        "peak_source": "independent DFMA chains measured in this run (jbugs_fp64_peak): one FP64 "
                                       "instruction per lane-slot, so a DADD/DMUL costs the same slot as a DFMA",
                        "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "algorithmic_bytes": alg_bytes, "peak_source": which}}
        except Exception as e:  # noqa: BLE001
            roofline["fp64_note"] = f"fp64 bound not reported: {e}"
    if plan.dense:
        # dense linear predictor: the two X products dominate -> FP64 tensor pipe.  Denominator: cuBLAS DGEMM measured
        # here, in the same process (MEASURED_PEAKS.json has no FP64 entry).  Rows of X local to this rank.
        dn = plan.dense[0]
        flops = 4.0 * dense_rows * dn.P * C
        n = 8192
        a = torch.randn((n, n), dtype=torch.float64, device=dev)
        b = torch.randn((n, n), dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s0, s1 = Ev(), Ev()
            s0.record()
            torch.matmul(a, b)
            s1.record()
            torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        dgemm_tf = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        del a, b
        ach = flops / (pk_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach / dgemm_tf,
                    "traffic": None, "kernel": "dense_fwd_kernel + dense_bwd_kernel (DMMA m8n8k4 f64)", "kernel_ms": pk_ms,
                    "kernel_ms_max": float(np.max(plate_ms)), "kernel_ms_samples": int(len(plate_ms)),
                    "eval_ms": float(np.mean(total_ms)), "algorithmic_flops": flops, "local_rows": dense_rows,
                    "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run, best of 3"}

    out = {"metric": METRIC if want_grad else "logdensity obs-evals/s", "value": value, "unit": UNIT, "n_gpus": ctx.n_gpus,
           "steps": K, "warmup": W, "ms_per_step": ms_per_step,
           "ms_per_step_p50": float(np.percentile(step_ms, 50)), "ms_per_step_p95": float(np.percentile(step_ms, 95)),
           "higher_is_better": True, "scaling": "strong" if shard == "obs" else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "roofline": roofline,
           "gpu_launches": launches_per_step * K}
    if sustained:
        out["sustained"] = sustained

    # ---- e2e: host pinned theta -> C ABI -> host grad; every chain of the batch, slab by slab, both host layouts
    if primary and not args.no_e2e and not obs_sharded:
        out["e2e"] = e2e(ctx, args, cp, theta, grad, lp_host, n_obs, D, C, K)
    res = (out, plan, th0, n_obs)
    del theta, grad
    cp.close()
    model._cuda = None
    torch.cuda.empty_cache()
    return res


def pcie_ceiling(dev):
    """what this box's host link gives plain cudaMemcpyAsync on pinned 1 GiB buffers: each direction alone and both at
    once (two streams).  The e2e step moves 16 bytes per parameter per chain, so this is its bound."""
    import torch
    n = (1 << 30) // 8
    h_in, p1 = pinned(8 * n)
    h_out, p2 = pinned(8 * n)
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.empty(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h, reps=3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return 8 * n / ((time.perf_counter() - t) / reps) / 1e9

    run(True, True, 1)
    out = {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "both_each_way_gbs": run(True, True),
           "how": "cudaMemcpyAsync of pinned 1 GiB buffers on two streams, measured in this run"}
    del h_in, h_out, d_in, d_out
    from jbugs import cuda as jcuda
    jcuda.host_free(p1)
    jcuda.host_free(p2)
    return out


def e2e(ctx, args, cp, theta, grad, lp_host, n_obs, D, C, K):
    """the metric through jbugs_eval_batch with HOST buffers: one step = every chain of the batch, in slabs of
    `Ce` chains that are staged through ONE pair of pinned host buffers (pinning 2 x 65.5 GB for all 4096 chains at
    once is not reasonable on a shared host; each slab's bytes cross PCIe in both directions inside the timed region)."""
    import torch
    from jbugs import cuda as jcuda
    dist, world, dev = ctx.dist, ctx.world, ctx.dev
    stream = ctx.stream.cuda_stream
    # slab size: 2 x 16 GB of pinned host memory per rank at one GPU; less per rank when several ranks share the host
    Ce = min(args.e2e_chains or (1024 if world == 1 else (512 if world == 2 else 256)), C)
    while Ce > 64 and 2 * 8 * D * Ce > (40 << 30):
        Ce //= 2
    n_slabs = (C + Ce - 1) // Ce
    res = {}
    th_h, th_p = pinned(8 * D * Ce, write_combined=True)  # written by the host, read by the copy engine only
    g_h, g_p = pinned(8 * D * Ce)
    lp_h, lp_p = pinned(8 * Ce)
    st_p = jcuda.host_alloc(4 * Ce)
    th_h = th_h.view(D, Ce)
    th_h.copy_(theta[:, :Ce])  # slab 0's values (the other slabs re-send the same bytes: the copies are what is timed)
    torch.cuda.synchronize()
    # free the resident batch: the staging slots of the host path get the device memory (chunks of 256 chains)
    theta.untyped_storage().resize_(0)
    grad.untyped_storage().resize_(0)
    torch.cuda.empty_cache()
    res["pcie_ceiling"] = pcie_ceiling(dev)
    for lay_name, lay in (("chain_fast", 0), ("param_fast", 1)):
        if lay == 1:
            # a Julia Matrix{Float64}(D, Ce): one contiguous column per chain -- what eval_batch! of the Julia glue sends
            tmp = th_h.clone()
            th_h.view(Ce, D).copy_(tmp.t())
            del tmp
        ld = Ce if lay == 0 else D

        def e2e_step():
            # asynchronous host calls: the library's copy-in / kernels / copy-out pipeline runs on from one slab into the
            # next (no fill / drain bubble per call); one synchronisation when the whole batch is back in host memory
            for _ in range(n_slabs):
                cp.eval_raw(th_p, ld, Ce, lay, lp_p, g_p, st_p, True, 1, stream)
            cp.synchronize(stream)

        e2e_step()
        reps = max(2, min(K, 3))
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / reps
        dt_local = dt
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        ok = bool(np.allclose(lp_h.numpy()[:Ce], lp_host[:Ce], rtol=1e-12, atol=0))
        h2d = 8 * D * Ce * n_slabs
        d2h = (8 * D * Ce + 12 * Ce) * n_slabs
        res[lay_name] = {"value": n_obs * Ce * n_slabs * world / dt, "ms_per_step": dt * 1e3, "matches_resident": ok,
                         "h2d_gbs_this_rank": h2d / dt_local / 1e9, "d2h_gbs_this_rank": d2h / dt_local / 1e9}
    ceiling = res.pop("pcie_ceiling")
    best = max(res, key=lambda k: res[k]["value"])
    out = {"value": res["chain_fast"]["value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "chains": Ce * n_slabs, "slab_chains": Ce, "slabs_per_step": n_slabs,
           "ms_per_step": res["chain_fast"]["ms_per_step"], "layout": "chain_fast",
           "matches_resident": res["chain_fast"]["matches_resident"] and res["param_fast"]["matches_resident"],
           "h2d_gbs_per_rank": res["chain_fast"]["h2d_gbs_this_rank"], "d2h_gbs_per_rank": res["chain_fast"]["d2h_gbs_this_rank"],
           "param_fast": res["param_fast"], "best_layout": best,
           "numa": ctx.numa, "pcie_ceiling": ceiling,
           "sample": f"all {Ce * n_slabs} chains per GPU in {n_slabs} slabs of {Ce} through jbugs_eval_batch with pinned "
                     "host buffers (theta write-combined; JBUGS_EVAL_ASYNC calls, one jbugs_plan_synchronize per step), H2D and D2H inside the timed region; `value` is the "
                     "chain_fast host layout, `param_fast` = a Julia Matrix(D, C) incl. the two on-device transposes"}
    for p in (th_p, g_p, lp_p, st_p):
        jcuda.host_free(p)
    return out


# ----------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: enough for a >= 3 s timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rats")
    ap.add_argument("--groups", type=int, default=0, help="plate size override (default: the BASELINE config)")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default 4096 for rats, 256 for seeds)")
    ap.add_argument("--e2e-chains", type=int, default=0, help="chains per host slab of the e2e measurement")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the sustained region after the K timed steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--secondary", default="seeds,seeds_c64,dense,pumps,epil,rats_variant,mixture")
    ap.add_argument("--dynamic", type=int, default=0, help="1 = force the interpreter kernel")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--tma", type=int, default=-1, help="0 = stage plate data with cp.async only (A/B against the TMA path)")
    ap.add_argument("--logdensity-only", action="store_true", help="time logdensity without the gradient (the reference benchmark's first column)")
    ap.add_argument("--opt", default="", help="comma-separated key=value options passed to jbugs_plan_set_option (A/B runs)")
    ap.add_argument("--shard", default="", choices=["", "chains", "obs"],
                    help="N>1: 'chains' = every GPU its own chains (weak, no collective; default for rats); "
                         "'obs' = the data plate is split by observation + NCCL all-reduce (strong; default for seeds)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    W = max(args.warmup, 3)
    C = args.chains or DEFAULT_CHAINS.get(args.workload, 1024)
    # default K: a timed region of >= 3 s at the step times measured on B200 (rats 21 ms, seeds 15 ms, dense 300 ms)
    K = args.steps if args.steps > 0 else {"rats": 150, "seeds": 220, "dense": 12, "pumps": 200, "epil": 200}.get(args.workload, 100)

    # ---------------- reference arm: CPU restatement only (rank 0)
    if args.impl == "reference":
        if rank != 0:
            return
        import jbugs
        model_text, data, th0, desc, _ = make_workload(args.workload, args.groups)
        plan = jbugs.compile(model_text, data).plan
        n_obs, D = plan.n_obs(), plan.D
        shard = args.shard or ("obs" if args.workload in ("seeds", "dense") else "chains")
        config = {"workload": f"{desc} x {C} chains{'' if shard == 'obs' else '/GPU'}, fp64, theta+grad resident in HBM (chain-fast layout)",
                  "chains_per_gpu": C, "D": D, "n_obs": n_obs,
                  "parallelism": "host cores, OpenMP, one chain per thread"}
        vals = []
        # bounded: the whole run (W warm-up + K timed samples) stays near two minutes whatever K is -- a step is at least
        # one chain per host thread (~0.5 s for the Rats plate), more chains per thread when K is small
        for _ in range(min(W, 3)):
            cpu_baseline(plan, th0, n_obs, budget_s=0.5)
        for _ in range(K):
            vals.append(cpu_baseline(plan, th0, n_obs, budget_s=max(0.25, 100.0 / K)))
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": n_gpus,
                          "steps": K, "warmup": W, "ms_per_step": float(np.mean([x["seconds"] for x in vals])) * 1e3,
                          "higher_is_better": True, "scaling": "strong" if shard == "obs" else "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "Julia is not installed: the reference arm is the C restatement of the reference's "
                                  "generated evaluator (oracle/jbugs_oracle.c), OpenMP one chain per thread"}))
        return

    # ---------------- our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    from jbugs import cuda as jcuda
    try:
        all_cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        all_cores = list(range(os.cpu_count() or 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local_rank, ctx.dev, ctx.n_gpus = rank, world, local_rank, dev, n_gpus
    ctx.dist = None
    # pinned staging of the host path must sit on the GPU's own NUMA node (8 ranks on one node share its memory system)
    node, ncpu = jcuda.bind_host_to_device(local_rank)
    ctx.numa = {"node": node, "cpus_bound": ncpu, "cpus_before": len(all_cores),
                "how": "jbugs_bind_host_to_device: sched_setaffinity to the node's CPUs + set_mempolicy(MPOL_PREFERRED) "
                       "before any pinned allocation" if node >= 0 else "platform reports no NUMA node for the GPU: not bound"}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        ctx.dist = dist
    ctx.sampler = ClockSampler(local_rank)
    ctx.sampler.start()
    # a real (non-default) stream: the kernels are launched on it and the events of the timed region recorded on it
    ctx.stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ctx.stream)
    time.sleep(0.3)

    out, plan, th0, n_obs = measure(args.workload, ctx, args, K, W, C, primary=True)

    if not args.no_secondary and not args.groups and not args.chains:
        sec = []
        for nm in [s for s in args.secondary.split(",") if s and s != args.workload]:
            Ks = {"seeds": 20, "seeds_c64": 40, "dense": 3, "pumps": 50, "epil": 50, "rats_variant": 40, "mixture": 10}.get(nm, 10)
            try:
                if nm == "seeds_c64":   # BASELINE config 3 leaves the chain count open (64 or 256): both are reported
                    o, _, _, _ = measure("seeds", ctx, args, Ks, 3, 64, primary=False)
                else:
                    o, _, _, _ = measure(nm, ctx, args, Ks, 3, DEFAULT_CHAINS[nm], primary=False)
                sec.append({k: o[k] for k in ("value", "unit", "ms_per_step", "ms_per_step_p50", "ms_per_step_p95", "steps",
                                              "warmup", "scaling", "config", "clocks", "roofline", "gpu_launches")})
                sec[-1]["name"] = nm
            except BaseException as e:  # noqa: BLE001 - a secondary workload must never void the primary line
                sec.append({"name": nm, "failed": f"{type(e).__name__}: {e}"})
                if world > 1:
                    break  # the ranks may be out of step after a failure: stop here
        out["secondary"] = sec
    ctx.sampler.stop()
    if rank == 0 and not args.no_cpu and world == 1:
        try:
            os.sched_setaffinity(0, all_cores)  # the CPU baseline gets every core, not just the GPU's node
        except Exception:
            pass
        try:
            out["cpu_baseline"] = cpu_baseline(plan, th0, n_obs, cores=len(all_cores))
        except Exception as e:  # the oracle is a checker; never let it break the bench line
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
