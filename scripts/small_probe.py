"""Pumps / Epil at 8192 chains: microseconds per evaluation as a function of the group tile (option "tile")."""
import sys, json
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch, jbugs
fx = json.load(open('tests/golden/examples_vol1.json'))
def arr(d): return {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in d.items()}
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for name, text in (("pumps", jbugs.examples.PUMPS), ("epil", jbugs.examples.EPIL)):
    m = jbugs.compile(text, arr(fx[name]["data"]), arr(fx[name]["inits"]))
    cp = m.cuda
    if __import__("os").environ.get("JB_DYNAMIC"): cp.set_option("dynamic", 1)
    D = m.plan.D
    th0 = jbugs.getparams(m); th0 = np.where(np.isfinite(th0), th0, 0.0)
    theta = torch.from_numpy(th0)[:, None].cuda() + 0.02 * torch.randn((D, C), dtype=torch.float64, device='cuda')
    grad = torch.empty_like(theta); lp = torch.empty(C, dtype=torch.float64, device='cuda'); st = torch.empty(C, dtype=torch.int32, device='cuda')
    for one, tile in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (0, 2), (0, 8)):
        cp.set_option("one_launch", one)
        cp.set_option("tile", tile)
        for _ in range(20): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 200
        for _ in range(n): cp.eval_raw(theta.data_ptr(), C, C, 0, lp.data_ptr(), grad.data_ptr(), st.data_ptr(), True, 1, None)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} C={C} one_launch={one} tile={tile}: {e0.elapsed_time(e1) / n * 1e3:.2f} us  | {cp.describe().splitlines()[1][:110]}", flush=True)
