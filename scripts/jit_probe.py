"""interpreter vs run-time specialised kernel on a plate shape that has no in-tree specialisation (a probit variant of the
Seeds model, 2e6 observations x 256 chains)."""
import sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests')
import numpy as np, torch, jbugs
from helpers import synth_seeds
N, C = 2_000_000, 256
data, b = synth_seeds(N, seed=1)
text = jbugs.examples.SEEDS.replace("logit(p[i])", "probit(p[i])")
m = jbugs.compile(text, data)
cp = m.cuda
D = m.plan.D
th0 = np.concatenate([[np.log(1 / 0.09), -0.4, 0.7, 0.05, -0.3], 0.5 * b])
theta = torch.from_numpy(th0)[:, None].cuda() + 0.02 * torch.randn((D, C), dtype=torch.float64, device='cuda')
grad = torch.empty_like(theta); lp = torch.empty(C, dtype=torch.float64, device='cuda'); st = torch.empty(C, dtype=torch.int32, device='cuda')
def run(tag):
    for _ in range(3): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 0, None)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(tag, cp.describe().splitlines()[1].split("kernel=")[1].split()[0], "%.2f ms" % (dt * 1e3), "%.3e obs-evals/s" % (N * C / dt), "hbm frac %.3f" % (16.0 * C * N / dt / 6457.4e9))
    return lp.clone()
a = run("interpreter ")
t = time.time(); cp.specialize(); print("specialise: %.1f s" % (time.time() - t))
b2 = run("specialised ")
print("max rel diff of logp", float(((a - b2).abs() / a.abs()).max()))
