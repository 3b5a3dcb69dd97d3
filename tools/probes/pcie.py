import torch, time
dev=torch.device('cuda',0)
n=348*1024*1024//8
h=torch.empty(n,dtype=torch.float64).pin_memory(); g=torch.empty(n,dtype=torch.float64,device=dev)
h2=torch.empty(n//3,dtype=torch.float64).pin_memory(); g2=torch.empty(n//3,dtype=torch.float64,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
d2h=t(lambda: h.copy_(g,non_blocking=True)); print('D2H %.1f GB/s'%(n*8/d2h/1e9))
h2d=t(lambda: g.copy_(h,non_blocking=True)); print('H2D %.1f GB/s'%(n*8/h2d/1e9))
def both():
    with torch.cuda.stream(s1): h.copy_(g,non_blocking=True)
    with torch.cuda.stream(s2): g2.copy_(h2,non_blocking=True)
b=t(both); print('duplex: D2H %.1f GB/s + H2D %.1f GB/s in %.2f ms'%(n*8/b/1e9, n//3*8/b/1e9, b*1e3))
# chunked D2H in 8 pieces on 3 streams
ss=[torch.cuda.Stream() for _ in range(3)]
def chunked():
    c=n//8
    for i in range(8):
        with torch.cuda.stream(ss[i%3]): h[i*c:(i+1)*c].copy_(g[i*c:(i+1)*c],non_blocking=True)
c=t(chunked); print('chunked D2H %.1f GB/s'%(n*8/c/1e9))
