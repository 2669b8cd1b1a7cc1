"""1-GPU timing of the sharded plan / expand kernels at the 8-GPU batch size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = G << 20
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
shift = 27
idx = torch.from_numpy(rng.integers(0, G << shift, M)).to(dev)
vals = torch.from_numpy(rng.random(M, dtype=np.float32)).to(dev)
tree = DeviceSumTree(1 << 20, dtype=torch.float32, device=dev)
for rep in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    shard, slot, own_ids, own_vals, counts = tree.shard_plan(None, idx, shift, G, 3 % G, vals, True)
    b.record()
    torch.cuda.synchronize()
    print("plan: device %.1f us, wall %.1f us" % (a.elapsed_time(b) * 1e3, (time.perf_counter() - t0) * 1e6))
cap = max(counts)
slabs = torch.zeros((G, cap), dtype=torch.float32, device=dev)
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tree.shard_expand_values(shard, slot, slabs); b.record(); torch.cuda.synchronize()
    print("expand values: %.1f us" % (a.elapsed_time(b) * 1e3))
