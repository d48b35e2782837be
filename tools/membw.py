import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
for mb in (105, 512, 2048):
    n=mb*1024*1024//4
    a=torch.empty(n,device='cuda'); b=torch.empty(n,device='cuda')
    s=t(lambda: a.zero_()); print(f"{mb} MB fill : {n*4/s/1e9:8.1f} GB/s")
    s=t(lambda: b.copy_(a)); print(f"{mb} MB copy : {2*n*4/s/1e9:8.1f} GB/s (r+w)")
    s=t(lambda: a.sum()); print(f"{mb} MB read : {n*4/s/1e9:8.1f} GB/s")
