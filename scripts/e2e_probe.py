import sys, time, os
sys.path.insert(0,'juliabugs.jl_b200'); sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch, jbugs
from bench import make_workload
txt,data,th0,desc=make_workload('rats',0)
m=jbugs.compile(txt,data); cp=m.cuda; D=m.plan.D
Ce=256
th=torch.empty((D,Ce),dtype=torch.float64).pin_memory(); g=torch.empty((D,Ce),dtype=torch.float64).pin_memory()
lp=torch.empty(Ce,dtype=torch.float64).pin_memory(); st=torch.empty(Ce,dtype=torch.int32).pin_memory()
th.copy_(torch.from_numpy(th0)[:,None]+0.05*torch.randn((D,Ce),dtype=torch.float64))
# raw PCIe reference: contiguous copies
d=torch.empty((D,Ce),dtype=torch.float64,device='cuda')
for _ in range(2): d.copy_(th,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); d.copy_(th,non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()-t
t=time.perf_counter(); g.copy_(d,non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()-t
print('contiguous H2D GB/s',D*Ce*8/t1/1e9,'D2H',D*Ce*8/t2/1e9)
del d
for hc in (16,32,64,128,256):
    cp.set_option('host_chunk',hc)
    for _ in range(2): cp.eval_raw(th.data_ptr(),Ce,Ce,0,lp.data_ptr(),g.data_ptr(),st.data_ptr(),True,0,None)
    t=time.perf_counter()
    for _ in range(3): cp.eval_raw(th.data_ptr(),Ce,Ce,0,lp.data_ptr(),g.data_ptr(),st.data_ptr(),True,0,None)
    dt=(time.perf_counter()-t)/3
    print('host_chunk',hc,'ms',dt*1e3,'GB/s each way',D*Ce*8/dt/1e9)
# param-fast host layout (Julia D x C): contiguous chain columns
thp=torch.empty((Ce,D),dtype=torch.float64).pin_memory(); gp=torch.empty((Ce,D),dtype=torch.float64).pin_memory()
thp.copy_(th.t())
for hc in (8,16,32,64):
    cp.set_option('host_chunk',hc)
    for _ in range(2): cp.eval_raw(thp.data_ptr(),D,Ce,1,lp.data_ptr(),gp.data_ptr(),st.data_ptr(),True,0,None)
    t=time.perf_counter()
    for _ in range(3): cp.eval_raw(thp.data_ptr(),D,Ce,1,lp.data_ptr(),gp.data_ptr(),st.data_ptr(),True,0,None)
    dt=(time.perf_counter()-t)/3
    print('param_fast host_chunk',hc,'ms',dt*1e3,'GB/s each way',D*Ce*8/dt/1e9)
