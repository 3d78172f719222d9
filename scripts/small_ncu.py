import sys, json
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch, jbugs
fx = json.load(open('tests/golden/examples_vol1.json'))
def arr(d): return {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in d.items()}
C = 8192
for name, text in (("pumps", jbugs.examples.PUMPS), ("epil", jbugs.examples.EPIL)):
    m = jbugs.compile(text, arr(fx[name]["data"]), arr(fx[name]["inits"]))
    cp = m.cuda
    D = m.plan.D
    th0 = jbugs.getparams(m); th0 = np.where(np.isfinite(th0), th0, 0.0)
    theta = torch.from_numpy(th0)[:, None].cuda() + 0.02 * torch.randn((D, C), dtype=torch.float64, device='cuda')
    grad = torch.empty_like(theta); lp = torch.empty(C, dtype=torch.float64, device='cuda'); st = torch.empty(C, dtype=torch.int32, device='cuda')
    for one in (0, 1):
        cp.set_option("one_launch", one)
        for _ in range(6): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
        torch.cuda.synchronize()
