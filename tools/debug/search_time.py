"""Sample (prefix-search) timing at 2^27 float32 leaves after a few fresh update batches; ST_SEARCH_INDEX=1 enables the
sector-packed index.  Prints us per 2^20 samples and the update time next to it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree
n, m = 1 << 27, 1 << 20
dev = torch.device("cuda")
t = DeviceSumTree(torch.rand(n, device=dev)); t.reserve(m)
bat = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev), torch.rand(m, device=dev, dtype=torch.float64)) for _ in range(14)]
out = torch.empty(m, dtype=torch.int64, device=dev)
for i, v, u in bat[:3]:
    t.update_(i, v); t.sample_uniform(u, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * 11 + 1)]
ev[0].record()
for k, (i, v, u) in enumerate(bat[3:]):
    t.update_(i, v); ev[2 * k + 1].record()
    t.sample_uniform(u, out=out); ev[2 * k + 2].record()
torch.cuda.synchronize()
up = [ev[2 * k].elapsed_time(ev[2 * k + 1]) * 1e3 for k in range(11)]
sm = [ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) * 1e3 for k in range(11)]
t.poll()
print("ST_SEARCH_INDEX=%s update us median %.1f | sample us median %.1f min %.1f  (%.2fe9 samples/s)" % (
    os.environ.get("ST_SEARCH_INDEX", "0"), np.median(up), np.median(sm), min(sm), m / np.median(sm) / 1e3))
