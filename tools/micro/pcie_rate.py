import torch, time
x=torch.empty(655360000//4,dtype=torch.float32).pin_memory(); y=torch.empty(327024640//4,dtype=torch.float32).pin_memory()
dx=torch.empty_like(x,device='cuda'); dy=torch.empty_like(y,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for mode in ('h2d','d2h','both'):
    for _ in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter()
        if mode in ('h2d','both'):
            with torch.cuda.stream(s1): dx.copy_(x,non_blocking=True)
        if mode in ('d2h','both'):
            with torch.cuda.stream(s2): y.copy_(dy,non_blocking=True)
        torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(mode, 'ms', dt*1e3, 'GB/s h2d', 0.65536/dt if mode!='d2h' else None, 'd2h', 0.327/dt if mode!='h2d' else None)
