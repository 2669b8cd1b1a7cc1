"""1-GPU: device time of update_with_deltas_ (2^20 updates, 2^27 leaves) when the host is ahead of the
GPU (back-to-back calls) versus right after a host sync (what a sharded step sees after its plan)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
n, m = 1 << 27, 1 << 20
tree = DeviceSumTree(torch.from_numpy(rng.random(n, dtype=np.float32)).to(dev))
tree.reserve(m)
batches = [(torch.from_numpy(rng.integers(0, n, m)).to(dev), torch.from_numpy(rng.random(m, dtype=np.float32)).to(dev)) for _ in range(4)]
out = torch.empty(m, dtype=torch.float32, device=dev)
for i, v in batches: tree.update_with_deltas_(i, v, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
ev[0].record()
for k in range(20):
    tree.update_with_deltas_(*batches[k % 4], out=out); ev[k + 1].record()
torch.cuda.synchronize()
print("back-to-back: %.1f us per call" % (ev[0].elapsed_time(ev[20]) * 1e3 / 20))
ts, hs = [], []
for k in range(20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); tree.update_with_deltas_(*batches[k % 4], out=out); b.record()
    hs.append((time.perf_counter() - t0) * 1e6)
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print("after a sync: %.1f us per call (median), host enqueue time %.1f us" % (np.median(ts), np.median(hs)))
