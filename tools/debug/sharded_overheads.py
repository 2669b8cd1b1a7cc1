"""1-GPU microbenchmark of the O(M_global) replicated work a rank does per sharded step at G ranks."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
def gpu_time(fn, reps=8, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return float(np.median(ts))
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = G << 20
rng = np.random.default_rng(0)
dev = 'cuda'
n_local = 1 << 27
shift = 27
rank = 3 % G
idx = torch.from_numpy(rng.integers(0, n_local * G, M)).to(dev)
vals = torch.from_numpy(rng.random(M, dtype=np.float32)).to(dev)
u = torch.from_numpy(rng.random(M)).to(dev)
top = DeviceSumTree(torch.from_numpy((rng.random(G) * 6.7e7).astype(np.float32)).to(dev))
d_all = torch.from_numpy((rng.random(M, dtype=np.float32) - 0.5)).to(dev)
shard = idx >> shift
print('G', G, 'M', M)
print('shard = idx >> shift        ', gpu_time(lambda: idx >> shift))
print('mask + nonzero              ', gpu_time(lambda: torch.nonzero(shard == rank).reshape(-1)))
mine = torch.nonzero(shard == rank).reshape(-1)
print('idx[mine] - off, vals[mine] ', gpu_time(lambda: (idx[mine] - 5, vals[mine])))
print('zeros + scatter d_all[mine] ', gpu_time(lambda: torch.zeros(M, dtype=torch.float32, device=dev).index_put_((mine,), vals[mine])))
print('top.apply_deltas_ (masked)  ', gpu_time(lambda: top.apply_deltas_(shard, d_all)))
print('top.route_uniform           ', gpu_time(lambda: top.route_uniform(u)))
out = torch.empty(M, dtype=torch.int64, device=dev)
loc = torch.zeros(mine.numel(), dtype=torch.int64, device=dev)
def comb():
    out.zero_(); out[mine] = loc + 7
print('out.zero_ + out[mine] = ... ', gpu_time(comb))
