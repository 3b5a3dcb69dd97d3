import sys, time, threading, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from ilqr_iterative_tasks_b200 import _abi, engine
params=_abi.default_params(); N=6; B=1<<20
d,obs=bench.make_workload(B,N,0)
pin=lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
x0,xt,lamb=pin(d['x0']),pin(d['x_term']),pin(d['lamb_in'])
obs_p=np.frombuffer(pin(obs.view(np.uint8)),dtype=_abi.OBSTACLE_DTYPE)
def mkout(): return {"uvar":pin(np.empty((B,2,N))),"xvar":pin(np.empty((B,4,N+1))),"lamb_out":pin(np.empty(B)),"iters":pin(np.empty(B,dtype=np.int32))}
for nthr in (1,2,3):
    engs=[engine.Engine(params,device=0) for _ in range(nthr)]; outs=[mkout() for _ in range(nthr)]
    def work(i,n):
        for _ in range(n): engs[i].solve_batch(x0,xt,lamb,obs_p,out=outs[i])
    for i in range(nthr): work(i,1)
    torch.cuda.synchronize()
    steps=48
    th=[threading.Thread(target=work,args=(i,steps//nthr)) for i in range(nthr)]
    t0=time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('threads %d: %.2f ms/step  %.1f M solves/s; last call pipeline_ms %.2f'%(nthr, dt/steps*1e3, B*steps/dt/1e6, engs[0].last_stats()['kernel_ms']))
    for e in engs: e.close()
