import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from ilqr_iterative_tasks_b200 import _abi, engine
params=_abi.default_params(); N=6; B=1<<20
d,obs=bench.make_workload(B,N,0)
dev=torch.device('cuda',0)
g=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x0,xt,lamb=g(d['x0']),g(d['x_term']),g(d['lamb_in']); gobs=torch.from_numpy(obs.view(np.uint8).reshape(B,-1).copy()).to(dev)
u=torch.empty((B,2,N),dtype=torch.float64,device=dev); x=torch.empty((B,4,N+1),dtype=torch.float64,device=dev); lo=torch.empty(B,dtype=torch.float64,device=dev); it=torch.empty(B,dtype=torch.int32,device=dev)
e=engine.Engine(params,device=0)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
n=348*1024*1024//8
hbuf=torch.empty(n,dtype=torch.float64).pin_memory(); gbuf=torch.empty(n,dtype=torch.float64,device=dev)
hin=torch.empty(n//3,dtype=torch.float64).pin_memory(); gin=torch.empty(n//3,dtype=torch.float64,device=dev)
def solve(st): e.solve_batch_device(B,x0.data_ptr(),xt.data_ptr(),lamb.data_ptr(),gobs.data_ptr(),B,u.data_ptr(),x.data_ptr(),lo.data_ptr(),it.data_ptr(),stream=st.cuda_stream)
ev=[torch.cuda.Event(enable_timing=True) for _ in range(6)]
for rep in range(3):
    solve(s2); torch.cuda.synchronize()
    with torch.cuda.stream(s2):
        ev[0].record(); solve(s2); ev[1].record()
    torch.cuda.synchronize(); alone=ev[0].elapsed_time(ev[1])
    with torch.cuda.stream(s1):
        ev[2].record(); hbuf.copy_(gbuf,non_blocking=True); ev[3].record()
    with torch.cuda.stream(s2):
        ev[0].record(); gin.copy_(hin,non_blocking=True); ev[4].record(); solve(s2); ev[1].record()
    torch.cuda.synchronize()
    print('solve alone %.2f ms | with concurrent D2H: h2d %.2f solve %.2f ms, D2H %.2f ms'%(alone, ev[0].elapsed_time(ev[4]), ev[4].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])))
