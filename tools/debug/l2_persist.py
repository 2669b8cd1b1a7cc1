import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
def gpu_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return float(np.median(ts))
rng=np.random.default_rng(0)
n,m=1<<27,1<<20
t=DeviceSumTree(torch.from_numpy(rng.random(n,dtype=np.float32)).cuda()); t.reserve(m)
idx=[torch.from_numpy(rng.integers(0,n,m)).cuda() for _ in range(4)]; val=torch.from_numpy(rng.random(m,dtype=np.float32)).cuda(); u=[torch.from_numpy(rng.random(m)).cuda() for _ in range(4)]
out=torch.empty(m,dtype=torch.int64,device='cuda'); k=[0]
def upd(): k[0]+=1; t.update_(idx[k[0]%4],val)
def smp(): k[0]+=1; t.sample_uniform(u[k[0]%4],out=out)
def both(): upd(); smp()
print('baseline   update', gpu_time(upd), 'sample', gpu_time(smp), 'step', gpu_time(both))
for mb in (16, 32, 64, 96):
    try:
        g = t.handle.set_l2_persistence(mb << 20, torch.cuda.current_stream().cuda_stream)
    except Exception as e:
        print('persist', mb, 'failed', e); continue
    print(f'persist {mb} MiB (granted {g>>20}) update', gpu_time(upd), 'sample', gpu_time(smp), 'step', gpu_time(both))
t.handle.set_l2_persistence(0, torch.cuda.current_stream().cuda_stream)
print('removed    update', gpu_time(upd), 'sample', gpu_time(smp))
