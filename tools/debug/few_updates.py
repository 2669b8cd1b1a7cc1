"""A handful of fresh 2^20-update batches on a 2^27-leaf float32 tree (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from starr_b200 import DeviceSumTree
n, m = 1 << 27, 1 << 20
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
t = DeviceSumTree(torch.rand(n, device=dev)); t.reserve(m)
bat = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev), torch.rand(m, device=dev, dtype=torch.float64)) for _ in range(reps)]
torch.cuda.synchronize()
for i, v, u in bat:
    t.update_(i, v)
    t.sample_uniform(u)
torch.cuda.synchronize()
t.poll()
print("ok")
