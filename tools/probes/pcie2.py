import torch, time
dev=torch.device('cuda',0)
n=348*1024*1024//8
h=torch.empty(n,dtype=torch.float64).pin_memory(); g=torch.empty(n,dtype=torch.float64,device=dev)
h2=torch.empty(n,dtype=torch.float64).pin_memory(); g2=torch.empty(n,dtype=torch.float64,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
e=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        e[0].record(); h.copy_(g,non_blocking=True); e[1].record()
    with torch.cuda.stream(s2):
        e[2].record(); g2.copy_(h2,non_blocking=True); e[3].record()
    torch.cuda.synchronize()
    print('concurrent equal sizes: D2H %.1f GB/s (%.2f ms)  H2D %.1f GB/s (%.2f ms)'%(n*8/e[0].elapsed_time(e[1])/1e6, e[0].elapsed_time(e[1]), n*8/e[2].elapsed_time(e[3])/1e6, e[2].elapsed_time(e[3])))
