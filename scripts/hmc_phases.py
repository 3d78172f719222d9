"""wall time of the phases of one on-device sampler run (create / init / warm-up / sampling) -- a profiling aid."""
import ctypes, json, sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests')
import numpy as np, jbugs
from jbugs import hmc
from jbugs.cuda import load, _check
fx = json.load(open('tests/golden/examples_vol1.json'))['rats']
data = {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in fx['data'].items()}
m = jbugs.compile(jbugs.examples.RATS, data, fx['inits'])
L = load(); hmc._bind(L)
cp = m.cuda
th0 = np.ascontiguousarray(jbugs.getparams(m))
for rep in range(2):
    h = ctypes.c_void_p(); t0 = time.time()
    _check(L.jbugs_hmc_create(cp.h, 256, 1234, ctypes.byref(h))); t1 = time.time()
    _check(L.jbugs_hmc_init(h, th0.ctypes.data, 0.05, 0.005)); t2 = time.time()
    acc = ctypes.c_double()
    _check(L.jbugs_hmc_run(h, 1500, 32, 1, None, 0, None, ctypes.byref(acc))); t3 = time.time()
    rows = np.arange(5, dtype=np.int64); out = np.empty((1000, 5, 256))
    _check(L.jbugs_hmc_run(h, 1000, 32, 0, rows.ctypes.data, 5, out.ctypes.data, ctypes.byref(acc))); t4 = time.time()
    L.jbugs_hmc_destroy(h)
    print("create %.3f init %.3f warmup(1500) %.3f sample(1000) %.3f  -> %.1f us per leapfrog step in sampling" %
          (t1 - t0, t2 - t1, t3 - t2, t4 - t3, (t4 - t3) / (1000 * 32) * 1e6))
