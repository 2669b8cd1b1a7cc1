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
for dt,tdt in ((np.int32,torch.int32),(np.int64,torch.int64),(np.float64,torch.float64)):
    leaves = rng.integers(0,1000,n).astype(dt) if dt!=np.float64 else rng.random(n)
    t=DeviceSumTree(torch.from_numpy(leaves).cuda()); t.reserve(m)
    idx=[torch.from_numpy(rng.integers(0,n,m)).cuda() for _ in range(4)]
    val=torch.from_numpy(rng.integers(0,1000,m).astype(dt) if dt!=np.float64 else rng.random(m)).cuda()
    u=torch.from_numpy(rng.random(m)).cuda(); out=torch.empty(m,dtype=torch.int64,device='cuda'); k=[0]
    def upd(): k[0]+=1; t.update_(idx[k[0]%4],val)
    print(np.dtype(dt).name, 'update us', gpu_time(upd), 'sample us', gpu_time(lambda: t.sample_uniform(u,out=out)), 'build us', gpu_time(lambda: t.build_(), reps=3))
    del t
