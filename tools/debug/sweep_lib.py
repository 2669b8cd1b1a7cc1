"""Tuning aid: update-phase timings on fresh batches for the library named by STARR_B200_LIB."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree
n, m = 1 << 27, 1 << 20
dev = torch.device("cuda")
t = DeviceSumTree(torch.rand(n, device=dev)); t.reserve(m)
bat = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev)) for _ in range(16)]
for i, v in bat[:4]: t.update_(i, v)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
ev[0].record()
for k, (i, v) in enumerate(bat[4:]):
    t.update_(i, v); ev[k + 1].record()
torch.cuda.synchronize()
ts = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(12)]
t.handle.set_profiling(True)
bat2 = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev)) for _ in range(8)]
for i, v in bat2: t.update_(i, v)
torch.cuda.synchronize()
bd = t.handle.update_breakdown()
print(os.environ.get("STARR_B200_LIB", "default"), "update us: median %.1f min %.1f max %.1f" % (np.median(ts), min(ts), max(ts)),
      {k: round(v * 1e3, 1) for k, v in bd.items() if k != "calls"})
