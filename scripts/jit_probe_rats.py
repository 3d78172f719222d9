"""a Rats variant that no in-tree kernel matches (one more global regression term): interpreter vs run-time specialised
kernel at 10^6 groups x 5 x 1024 chains, against the HBM roofline."""
import sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests')
import numpy as np, torch, jbugs
from helpers import synth_rats
N, C = 1_000_000, 1024
data, alpha, beta = synth_rats(N)
data["z"] = np.random.default_rng(3).standard_normal(N)
text = jbugs.examples.RATS.replace("mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar)", "mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar) + gam * z[i]") \
                          .replace("alpha0 <- alpha.c - xbar * beta.c", "alpha0 <- alpha.c - xbar * beta.c\n    gam ~ dnorm(0.0, 1.0E-2)")
m = jbugs.compile(text, data, evaluation_mode=jbugs.UseCUDABatch(specialize=False))
cp = m.cuda
D = m.plan.D
names = m.lowering.param_names(limit=10)[:8] if False else None
th0 = np.zeros(D)
g = {gl.name: gl.theta_idx for gl in m.plan.globals}
th0[g["tau.c"]] = np.log(1 / 36.0); th0[g["alpha.c"]] = 242.0; th0[g["alpha.tau"]] = np.log(1 / 196.0)
th0[g["beta.c"]] = 6.2; th0[g["beta.tau"]] = np.log(4.0)
loc = {l.name: l for l in m.plan.plates[0].locals}
th0[loc["alpha"].theta_base + loc["alpha"].stride_i * np.arange(N)] = alpha
th0[loc["beta"].theta_base + loc["beta"].stride_i * np.arange(N)] = beta
ld = C
theta = torch.from_numpy(th0)[:, None].cuda() + 0.02 * torch.randn((D, C), dtype=torch.float64, device='cuda')
grad = torch.empty_like(theta); lp = torch.empty(C, dtype=torch.float64, device='cuda'); st = torch.empty(C, dtype=torch.int32, device='cuda')
def run(tag, reps):
    for _ in range(2): cp.eval_raw(theta.data_ptr(), ld, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 0, None)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): cp.eval_raw(theta.data_ptr(), ld, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    print(tag, cp.describe().splitlines()[1].split("kernel=")[1].split()[0], "%.2f ms" % (dt * 1e3), "%.3e obs-evals/s" % (5 * N * C / dt), "hbm frac %.3f" % (16.0 * C * D / dt / 6457.4e9))
    return lp.clone()
a = run("interpreter ", 2)
t = time.time(); cp.specialize("none"); print("specialise: %.1f s" % (time.time() - t))
b2 = run("specialised ", 10)
print("max rel diff of logp", float(((a - b2).abs() / a.abs()).max()), "all finite", bool(torch.isfinite(b2).all()))
# the best variant arrives from the worker thread: eval_raw with flags = 0 is the polling entry point
t = time.time()
while "quick_variant" in cp.describe() and time.time() - t < 120:
    time.sleep(0.5)
    cp.eval_raw(theta.data_ptr(), ld, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 0, None)
print("best variant adopted after %.1f s" % (time.time() - t))
b3 = run("upgraded    ", 10)
print("max rel diff of logp", float(((a - b3).abs() / a.abs()).max()))
