import torch
x=torch.empty(1<<30,dtype=torch.uint8,device='cuda'); y=torch.empty_like(x)
def t(f,n=10):
    for _ in range(3): f()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
ms=t(lambda: x.fill_(1)); print('fill 1GiB GB/s', 1.0737/ms*1e3)
ms=t(lambda: y.copy_(x)); print('copy 1GiB r+w GB/s', 2*1.0737/ms*1e3)
ms=t(lambda: x.view(torch.float32).sum()); print('read 1GiB GB/s', 1.0737/ms*1e3)
