import sys, json, time
sys.path.insert(0,'juliabugs.jl_b200'); sys.path.insert(0,'tests')
import numpy as np, jbugs
from jbugs import hmc
fx=json.load(open('tests/golden/examples_vol1.json'))['rats']
data={k:(np.asarray(v,dtype=float) if isinstance(v,list) else v) for k,v in fx['data'].items()}
inits={k:(np.asarray(v,dtype=float) if isinstance(v,list) else v) for k,v in fx['inits'].items()}
m=jbugs.compile(jbugs.examples.RATS,data,inits)
for (L,na,ns,C) in [(24,700,500,256),(32,1500,1000,256),(48,2000,1000,512),(16,3000,1000,256)]:
    t=time.time()
    ch=hmc.sample(m,hmc.HMC(n_leapfrog=L,step_size=0.005,jitter=0.05),n_samples=ns,n_adapts=na,n_chains=C,seed=1234)
    a0=ch['alpha.c']-22.0*ch['beta.c']; sg=1/np.sqrt(ch['tau.c'])
    # per-chain means spread = convergence indicator
    print(L,na,ns,C,'time %.1fs'%(time.time()-t),'acc %.2f eps %.4f'%(ch.accept_rate,ch.step_size),
          'alpha0 %.2f±%.2f (chain-mean sd %.2f)'%(a0.mean(),a0.std(),a0.mean(axis=0).std()),
          'beta.c %.3f±%.3f'%(ch['beta.c'].mean(),ch['beta.c'].std()),'sigma %.3f±%.3f'%(sg.mean(),sg.std()))
fx=json.load(open('tests/golden/examples_vol1.json'))['seeds']
data={k:(np.asarray(v,dtype=float) if isinstance(v,list) else v) for k,v in fx['data'].items()}
inits={k:(np.asarray(v,dtype=float) if isinstance(v,list) else v) for k,v in fx['inits'].items()}
m=jbugs.compile(jbugs.examples.SEEDS,data,inits)
for (L,na,ns,C) in [(32,1500,1000,256),(64,2000,1000,256)]:
    t=time.time()
    ch=hmc.sample(m,hmc.HMC(n_leapfrog=L,step_size=0.01,jitter=0.05),n_samples=ns,n_adapts=na,n_chains=C,seed=7)
    sg=1/np.sqrt(ch['tau'])
    print('seeds',L,na,ns,C,'time %.1fs'%(time.time()-t),'acc %.2f eps %.4f'%(ch.accept_rate,ch.step_size),
          ' '.join('%s %.3f±%.3f'%(k,ch[k].mean(),ch[k].std()) for k in ['alpha0','alpha1','alpha12','alpha2']),'sigma %.3f±%.3f'%(sg.mean(),sg.std()))
