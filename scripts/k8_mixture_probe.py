import sys; sys.path.insert(0,'tests')
import conftest, numpy as np, jbugs
from c_oracle import COracle
K=8
text="model {\n"+"".join(f"  w[{k}] <- {1.0/K}\n" for k in range(1,K+1))+f"""  for (k in 1:{K}) {{ mu[k] ~ dnorm(0, 0.04)
    sigma[k] ~ dexp(1)
    tau[k] <- 1 / (sigma[k] * sigma[k]) }}
  for (i in 1:N) {{ z[i] ~ dcat(w[1:{K}])
    y[i] ~ dnorm(mu[z[i]], tau[z[i]]) }}
}}"""
rng=np.random.default_rng(0); N=200
y=rng.normal(rng.integers(0,K,N)*2.0-7, 0.5)
m=jbugs.compile(text,{"N":N,"y":y},evaluation_mode=jbugs.UseAutoMarginalization())
print("D",m.plan.D, "xops", [len(p.xops) for p in m.plan.plates])
th=0.3*rng.standard_normal((m.plan.D,70))
lp,g,st=m.cuda.eval_host(th,layout="param_fast")
lp0,g0,st0=COracle(m.plan).eval_batch(th)
print(np.max(np.abs(lp-lp0)/np.abs(lp0)), np.max(np.abs(g-g0))/np.max(np.abs(g0)))
m.cuda.specialize()
lp,g,st=m.cuda.eval_host(th,layout="param_fast")
print("jit", np.max(np.abs(lp-lp0)/np.abs(lp0)), np.max(np.abs(g-g0))/np.max(np.abs(g0)), m.cuda.describe()[:200])
