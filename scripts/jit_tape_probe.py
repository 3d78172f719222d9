"""interpreter vs run-time compiled expression tape (and marginalised mixture) on large plates; CUDA events around the
evaluations, device-resident buffers.  Writes one line per case (profiles/r2_jit_tape_probe.txt)."""
import sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch, jbugs
from test_marginalization import GMM2

def bench(m, th0, C, n_obs, bytes_per_chain, tag, reps=5):
    cp = m.cuda
    D = m.plan.D
    theta = torch.from_numpy(th0)[:, None].cuda() + 0.02 * torch.randn((D, C), dtype=torch.float64, device='cuda')
    grad = torch.empty_like(theta); lp = torch.empty(C, dtype=torch.float64, device='cuda'); st = torch.empty(C, dtype=torch.int32, device='cuda')
    def run(what):
        for _ in range(2): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 0, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        kern = cp.describe().splitlines()[1].split("kernel=")[1].split()[0]
        print(f"{tag:28s} {what:12s} {kern:42s} {ms:9.2f} ms  {n_obs * C / ms * 1e3:.3e} obs-evals/s  hbm frac {bytes_per_chain * C / ms / 1e6 / 6457.4:.3f}", flush=True)
        return lp.clone(), grad.clone()
    a, ga = run("interpreter")
    t = time.time(); cp.specialize("none"); dt = time.time() - t
    b, gb = run("compiled")
    print(f"{tag:28s} specialise {dt:.1f} s; max rel diff logp {float(((a - b).abs() / a.abs()).max()):.2e} grad {float(((ga - gb).abs() / (ga.abs() + 1e-3 * ga.abs().max())).max()):.2e}", flush=True)

rng = np.random.default_rng(1)
# (1) two-component normal mixture, latents summed out: globals only (theta is 4 x C), data-bound streams
N = 2_000_000
z = rng.random(N) < 0.3
y = np.where(z, rng.normal(-2, 1.0, N), rng.normal(2, 0.7, N))
m = jbugs.compile(GMM2 % (0.3, 0.7), {"N": N, "y": y}, evaluation_mode=jbugs.UseAutoMarginalization(specialize=False))
names = m.lowering.param_names()
vals = {"sigma[1]": 0.0, "sigma[2]": np.log(0.7), "mu[2]": 2.0, "mu[1]": -2.0}
bench(m, np.array([vals[n] for n in names]), 1024, N, 0.0, "gmm2 2e6 x 1024")
# (2) Dugongs-style growth curve with a random effect per group: y[i,j] ~ dnorm(a[i] - b * pow(g, x[j]), tau)
Ng, T = 200_000, 8
text = """model {
  for (i in 1:N) {
    a[i] ~ dnorm(mu.a, tau.a)
    for (j in 1:T) {
      y[i, j] ~ dnorm(m[i, j], tau)
      m[i, j] <- a[i] - beta * pow(gamma, x[j])
    }
  }
  mu.a ~ dnorm(0, 1.0E-4)
  tau.a ~ dgamma(0.01, 0.01)
  beta ~ dunif(0.1, 5)
  gamma ~ dunif(0.5, 1.0)
  tau ~ dgamma(0.001, 0.001)
}"""
x = np.linspace(1.0, 30.0, T)
a = rng.normal(2.6, 0.2, Ng)
Y = a[:, None] - 1.0 * 0.87 ** x[None, :] + rng.normal(0, 0.1, (Ng, T))
m2 = jbugs.compile(text, {"N": Ng, "T": T, "x": x, "y": Y}, evaluation_mode=jbugs.UseCUDABatch(specialize=False))
names = m2.lowering.param_names()
th = np.zeros(m2.plan.D)
for k, n in enumerate(names):
    if n.startswith("a["): th[k] = a[int(n[2:-1]) - 1]
    elif n == "mu.a": th[k] = 2.6
    elif n == "tau.a": th[k] = np.log(25.0)
    elif n == "tau": th[k] = np.log(100.0)
    elif n == "beta": th[k] = np.log((1.0 - 0.1) / (5 - 1.0))
    elif n == "gamma": th[k] = np.log((0.87 - 0.5) / (1.0 - 0.87))
bench(m2, th, 1024, Ng * T, 16.0 * Ng, "growth curve 2e5x8 x 1024")
