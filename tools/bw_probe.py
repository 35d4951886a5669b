import torch, time
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return best
N=1<<31
a=torch.empty(N,dtype=torch.float64,device='cuda'); b=torch.empty(N,dtype=torch.float64,device='cuda')
ms=t(lambda: a.fill_(1.0)); print('fill  write-only GB/s', N*8/ms/1e6)
ms=t(lambda: b.copy_(a)); print('copy  r+w GB/s', 2*N*8/ms/1e6)
ms=t(lambda: a.sum()); print('sum   read-only GB/s', N*8/ms/1e6)
# 1 read : 9 write mix like K1: out planes = 53, in planes = 6 -> emulate with a broadcasted expand write
c=torch.empty((9,N//8),dtype=torch.float64,device='cuda'); d=torch.empty(N//8,dtype=torch.float64,device='cuda')
ms=t(lambda: torch.add(d.unsqueeze(0).expand(9,-1), 1.0, out=c)); print('1r:9w GB/s', 10*(N//8)*8/ms/1e6)
