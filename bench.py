#!/usr/bin/env python
"""bench.py -- logdensity+gradient obs-evals/s of the batched B200 evaluator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload rats|seeds|epil|pumps]

A "step" is one batched `logdensity_and_gradient` (jbugs_eval_batch) over the whole synthetic data
plate for every chain of the batch.  Default workload = BASELINE.json configs[1]: Rats scaled to
10^6 rats x 5 timepoints, 4096 chains, fp64, one B200.  With N > 1 (torchrun) chains shard across
GPUs with no data-path collective (weak scaling: 4096 chains per GPU).

  value     obs-evals/s with theta/grad resident in HBM (CUDA events around exactly K steps)
  e2e       same metric through the C ABI with pinned HOST buffers (H2D of theta, D2H of grad/logp
            inside the timed region) on a bounded chain sample
  roofline  algorithmic bytes (2*8*C*D) / measured plate-kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the C restatement of the reference evaluator (oracle/) on the host cores, bounded sample

`--impl reference` times that CPU restatement alone (Julia is not installed here or on the GPU box,
so the reference itself cannot run; kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "juliabugs.jl_b200"))

METRIC = "logdensity+gradient obs-evals/s"
UNIT = "obs-evals/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------- workloads
def make_workload(name, groups):
    """returns (model_text, data, theta0 (D,), description)"""
    from jbugs import examples as ex
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import synth_rats, synth_seeds
    if name == "rats":
        N = groups or 1_000_000
        data, alpha, beta = synth_rats(N)
        D = 2 * N + 5
        th0 = np.empty(D)
        th0[:5] = [np.log(4.0), 6.2, np.log(1 / 196.0), 242.0, np.log(1 / 36.0)]  # beta.tau, beta.c, alpha.tau, alpha.c, tau.c
        th0[5::2] = beta
        th0[6::2] = alpha
        return ex.RATS, data, th0, f"rats N={N} T=5"
    if name == "seeds":
        N = groups or 10_000_000
        data, b = synth_seeds(N)
        th0 = np.empty(N + 5)
        th0[:5] = [np.log(1 / 0.09), -0.84, 1.36, 0.09, -0.55]  # tau, alpha12, alpha2, alpha1, alpha0
        th0[5:] = b
        return ex.SEEDS, data, th0, f"seeds N={N}"
    if name == "dense":
        # BASELINE.json config 4: dense hierarchical logistic regression, X (N x P) fp64, beta over chains
        import torch
        N, Pn = groups or (1 << 22), 512
        gen = torch.Generator().manual_seed(1234)
        X = torch.randn(Pn * N, dtype=torch.float64, generator=gen).numpy().reshape((Pn, N)).T  # F-contiguous N x P
        rng = np.random.default_rng(1234)
        bt = rng.normal(0.0, 0.1, Pn)
        y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ bt)))).astype(np.float64)
        th0 = np.concatenate([[np.log(100.0), 0.0], bt])  # tau.b, mu.b, beta[1..P]
        return ex.DENSE_LOGISTIC, {"N": N, "P": Pn, "X": X, "y": y}, th0, f"dense logistic N={N} P={Pn}"
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
        return (ex.PUMPS if name == "pumps" else ex.EPIL), data, th0, f"{name} (shipped fixture)"
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

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.samples:
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
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(plan, th0, n_obs, budget_s=15.0, seed=7):
    """the C restatement of the reference evaluator on the host cores, one chain per thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from c_oracle import COracle
    co = COracle(plan)
    # every host core this process may run on; not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
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
            "sample": f"{C} chains x {n_obs} obs, one chain per thread, {t1:.2f} s"}


# ----------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rats")
    ap.add_argument("--groups", type=int, default=0, help="plate size override (default: the BASELINE config)")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default 4096 for rats, 256 for seeds)")
    ap.add_argument("--e2e-chains", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dynamic", type=int, default=0, help="1 = force the interpreter kernel")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--tma", type=int, default=-1, help="0 = stage plate data with cp.async only (A/B against the TMA path)")
    ap.add_argument("--logdensity-only", action="store_true", help="time logdensity without the gradient (the reference benchmark's first column)")
    ap.add_argument("--opt", default="", help="comma-separated key=value options passed to jbugs_plan_set_option (A/B runs)")
    ap.add_argument("--shard", default="", choices=["", "chains", "obs"],
                    help="N>1: 'chains' = every GPU its own chains (weak, no collective; default for rats); "
                         "'obs' = the data plate is split by observation + NCCL all-reduce (strong; default for seeds)")
    args = ap.parse_args()
    shard = args.shard or ("obs" if args.workload == "seeds" else "chains")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    C = args.chains or {"rats": 4096, "seeds": 256, "pumps": 8192, "epil": 8192, "dense": 1024}.get(args.workload, 1024)

    import jbugs
    model_text, data, th0, desc = make_workload(args.workload, args.groups)
    t0 = time.time()
    model = jbugs.compile(model_text, data)
    plan = model.plan
    n_obs = plan.n_obs()
    D = plan.D
    log(f"[bench] lowered {desc}: D={D} obs={n_obs} in {time.time() - t0:.1f}s")
    obs_sharded = shard == "obs" and world > 1
    par = (f"data plate sharded by observation x{n_gpus} + ncclAllReduce of C x (1+G) doubles" if shard == "obs"
           else f"chains sharded x{n_gpus}, no collective")
    config = {"workload": f"{desc} x {C} chains{'' if shard == 'obs' else '/GPU'}, fp64, theta+grad resident in HBM (chain-fast layout)",
              "chains_per_gpu": C, "D": D, "n_obs": n_obs, "parallelism": par,
              "l2": "inputs (theta+grad = %.1f GB/GPU) exceed the 126 MB L2; no flush needed"
                    % (2 * 8 * C * D / 1e9 / (world if obs_sharded else 1))}
    scaling = "strong" if shard == "obs" else "weak"

    # ---------------- reference arm: CPU restatement only (rank 0)
    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for _ in range(W):
            cpu_baseline(plan, th0, n_obs, budget_s=2.0)
        for _ in range(K):
            vals.append(cpu_baseline(plan, th0, n_obs, budget_s=max(5.0, 60.0 / K)))
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": n_gpus,
                          "steps": K, "warmup": W, "ms_per_step": float(np.mean([x["seconds"] for x in vals])) * 1e3,
                          "higher_is_better": True, "scaling": scaling,
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
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    model = jbugs.set_evaluation_mode(model, jbugs.UseCUDABatch(device=local_rank))
    cp = model.cuda
    if args.dynamic:
        cp.set_option("dynamic", 1)
    if args.tile:
        cp.set_option("tile", args.tile)
    if args.tma >= 0:
        cp.set_option("tma", args.tma)
    for kv in filter(None, args.opt.split(",")):
        k, v = kv.split("=")
        cp.set_option(k, int(v))
    local_groups = [pl.N for pl in plan.plates]
    if obs_sharded:
        from jbugs.dist import shard_bounds, shard_observations
        shard_observations(model, rank, world)
        local_groups = [b - a for a, b in (shard_bounds(pl.N, world, rank) for pl in plan.plates)]
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
    # a real (non-default) stream: the kernels are launched on it and the events of the timed region recorded on it
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    want_grad = not args.logdensity_only
    if not want_grad:
        config["evaluation"] = "logdensity only (no gradient): theta is read, nothing but logp is written"

    def step(flags=1):
        cp.eval_raw(theta.data_ptr(), ld, C, 0, logp.data_ptr(), grad.data_ptr(), status.data_ptr(), want_grad, flags, stream)

    step(0)  # first call also builds the one-off data constants
    log("[bench] " + cp.describe().replace("\n", " | "))
    for _ in range(W):
        step(0)
    launches0 = cp.launch_count()
    step(0)
    launches_per_step = cp.launch_count() - launches0
    plate_ms = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    small = 2 * 8 * C * D_local < (1 << 28)  # working set could sit in the 126 MB L2: flush between timed steps
    tw0 = time.time()
    if small:
        flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for a, b in evs:
            flush.zero_()  # 256 MB write > L2, outside the timed event pair
            a.record()
            step(1)
            b.record()
        torch.cuda.synchronize()
        ms_total = float(sum(a.elapsed_time(b) for a, b in evs))
        config["l2"] = "working set fits in L2: a 256 MB buffer is rewritten before every timed step (outside the event pair)"
    else:
        # the library's own events around the plate kernel are recorded inside the timed region (four event records
        # per call: nothing against a 20 ms step); the last step's are read after the region
        cp.set_option("timing", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step(1)
        e1.record()
        torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        plate_ms.append(cp.last_timing()[0])
        cp.set_option("timing", 0)
    if world > 1:
        dist.barrier()
    tw1 = time.time()
    # per-kernel time of the dominant (plate) kernel: the library's own CUDA events on the launch stream
    if not plate_ms:  # small workloads: timed separately, the event records would be a visible share of a 30 us step
        cp.set_option("timing", 1)
        for _ in range(min(K, 5)):
            step(0)
            plate_ms.append(cp.last_timing()[0])
        cp.set_option("timing", 0)
    clocks = sampler.stop(tw0, tw1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = n_obs * C * (1 if obs_sharded else world) / (ms_per_step * 1e-3)
    lp_host = logp[:C].cpu().numpy()
    st_host = status[:C].cpu().numpy()
    if not (np.all(np.isfinite(lp_host)) and np.all(st_host == 0)):
        raise SystemExit("bench produced non-finite log densities: invalid run")

    # roofline of the plate kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, which = 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_bytes = (2.0 if want_grad else 1.0) * 8.0 * C * D_local
    pk_ms = float(np.mean(plate_ms))
    achieved = alg_bytes / (pk_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not args.groups:
        try:
            t = json.load(open(tpath)).get(args.workload)
            # the ncu --set full capture replays the kernel ~40x and must save/restore what it writes, so it was
            # taken at fewer chains than fit beside a 131 GB working set; DRAM bytes are linear in the chain count
            traffic = t["dram_bytes_per_launch"] * C / t["chains"]
            traffic_note = f"dram__bytes_read+write from {t['source']} (captured at {t['chains']} chains, scaled to {C})"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note, "kernel": "plate_kernel", "kernel_ms": pk_ms,
                "algorithmic_bytes": alg_bytes, "peak_source": which}

    if plan.dense:
        # dense linear predictor: the two X products dominate -> FP64 tensor pipe.  Denominator: cuBLAS DGEMM measured
        # here, in the same process (MEASURED_PEAKS.json has no FP64 entry).
        dn = plan.dense[0]
        flops = 4.0 * dn.N * dn.P * C
        n = 8192
        a = torch.randn((n, n), dtype=torch.float64, device=dev)
        b = torch.randn((n, n), dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
                    "algorithmic_flops": flops, "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run, best of 3"}

    out = {"metric": METRIC if want_grad else "logdensity obs-evals/s", "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "roofline": roofline,
           "gpu_launches": launches_per_step * K}

    # e2e: host pinned theta -> C ABI -> host grad, bounded chain sample
    if rank == 0 or world > 1:
        if not args.no_e2e and not obs_sharded:
            Ce = min(args.e2e_chains, C)
            th_h = torch.empty((D, Ce), dtype=torch.float64).pin_memory()
            g_h = torch.empty((D, Ce), dtype=torch.float64).pin_memory()
            lp_h = torch.empty(Ce, dtype=torch.float64).pin_memory()
            st_h = torch.empty(Ce, dtype=torch.int32).pin_memory()
            th_h.copy_(theta[:, :Ce])
            torch.cuda.synchronize()

            def e2e_step():
                cp.eval_raw(th_h.data_ptr(), Ce, Ce, 0, lp_h.data_ptr(), g_h.data_ptr(), st_h.data_ptr(), True, 0, stream)

            for _ in range(2):
                e2e_step()
            reps = max(2, min(K, 5))
            if world > 1:
                dist.barrier()
            t = time.perf_counter()
            for _ in range(reps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t) / reps
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            ok = bool(np.allclose(lp_h.numpy(), lp_host[:Ce], rtol=1e-12, atol=0))
            out["e2e"] = {"value": n_obs * Ce * world / dt, "unit": UNIT, "h2d_bytes_per_step": 8 * D * Ce,
                          "d2h_bytes_per_step": 8 * D * Ce + 12 * Ce, "chains": Ce, "ms_per_step": dt * 1e3,
                          "matches_resident": ok,
                          "sample": f"{Ce} of {C} chains per GPU through jbugs_eval_batch with pinned host buffers"}
            del th_h, g_h
    if rank == 0 and not args.no_cpu and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline(plan, th0, n_obs)
        except Exception as e:  # the oracle is a checker; never let it break the bench line
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
