"""Run a few PER-sized steps (config #2) -- used under ncu to profile the small-batch kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starr_b200 import DeviceSumTree
rng = np.random.default_rng(0)
n, m = 1 << 20, 4096
t = DeviceSumTree(torch.from_numpy(rng.random(n, dtype=np.float32)).cuda())
t.reserve(m)
out = torch.empty(m, dtype=torch.int64, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    idx = torch.from_numpy(rng.integers(0, n, m)).cuda()
    val = torch.from_numpy(rng.random(m, dtype=np.float32)).cuda()
    u = torch.from_numpy(rng.random(m)).cuda()
    t.update_(idx, val)
    t.sample_uniform(u, out=out)
torch.cuda.synchronize()
t.poll()
print("ok", int(out[0]))
