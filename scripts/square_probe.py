import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pplp_b200 import engine
n=8192
ctx = engine.Context(n, t=1<<20, device=0)
k=ctx.k
nq=512
a = ctx.empty(nq,2,k,n)
for j in range(k): a[:,:,j].random_(0, ctx.q[j])
for _ in range(2): out=ctx.square(a)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out=ctx.square(a)
e1.record(); torch.cuda.synchronize()
print('squares/s', nq*5/(e0.elapsed_time(e1)*1e-3))
