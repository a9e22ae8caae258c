import torch
x=torch.empty(50176*2048,device='cuda'); y=torch.empty_like(x)
def t(f,n=10):
    for _ in range(3): f()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True); torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
b=x.numel()*4
for name,f,by in [("fill",lambda:x.fill_(1.0),b),("zero_",lambda:x.zero_(),b),("copy",lambda:y.copy_(x),2*b)]:
    us=t(f); print(f"{name:8s} {us:7.1f} us {by/us/1e6:6.0f} GB/s")
