"""the on-device sampler on a model without an in-tree kernel (Stacks), interpreter vs run-time specialised kernel."""
import json, sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests')
import numpy as np, jbugs
from jbugs import hmc
fx = json.load(open('tests/golden/examples_vol1.json'))['stacks']
data = {k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in fx['data'].items()}
for spec in (False, True):
    m = jbugs.compile(jbugs.examples.STACKS, data, fx['inits'])
    t = time.time()
    ch = hmc.sample(m, hmc.HMC(n_leapfrog=16, step_size=0.05, jitter=0.05), n_samples=1000, n_adapts=1000, n_chains=1024, seed=11, specialize=spec)
    dt = time.time() - t
    print("specialize", spec, m.cuda.describe().splitlines()[1].split("kernel=")[1].split()[0], "wall %.2f s" % dt,
          "b0 %.2f +- %.2f" % (ch["b0"].mean(), ch["b0"].std()), "acc %.2f" % ch.accept_rate)
