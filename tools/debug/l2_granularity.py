import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
import glob; rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so.12'))[0])
def gpu_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return float(np.median(ts))
rng=np.random.default_rng(0)
n,m=1<<27,1<<20
torch.zeros(1).cuda()
for gran in (None, 32, 128, 64):
    if gran is not None:
        rc = rt.cudaDeviceSetLimit(ctypes.c_int(0x05), ctypes.c_size_t(gran))  # cudaLimitMaxL2FetchGranularity = 5
        v = ctypes.c_size_t(); rt.cudaDeviceGetLimit(ctypes.byref(v), ctypes.c_int(0x05))
        print('set', gran, 'rc', rc, 'now', v.value)
    t=DeviceSumTree(torch.from_numpy(rng.random(n,dtype=np.float32)).cuda()); t.reserve(m)
    idx=[torch.from_numpy(rng.integers(0,n,m)).cuda() for _ in range(4)]; val=torch.from_numpy(rng.random(m,dtype=np.float32)).cuda(); u=[torch.from_numpy(rng.random(m)).cuda() for _ in range(4)]
    out=torch.empty(m,dtype=torch.int64,device='cuda'); k=[0]
    def upd(): k[0]+=1; t.update_(idx[k[0]%4],val)
    def smp(): k[0]+=1; t.sample_uniform(u[k[0]%4],out=out)
    print(gran, 'update us', gpu_time(upd), 'sample us', gpu_time(smp), 'build us', gpu_time(lambda: t.build_(), reps=3))
    del t
