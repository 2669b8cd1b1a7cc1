"""Update / sample time against the batch size on one 2^27-leaf float32 tree (fresh batches): looks for a batch size
between the small path (<= 4096) and 2^20 where the pipeline falls off a cliff."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree
lgn = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << lgn
dev = torch.device("cuda")
t = DeviceSumTree(torch.rand(n, device=dev)); t.reserve(1 << 20)
for lgm in range(12, 21):
    m = 1 << lgm
    bat = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev), torch.rand(m, device=dev, dtype=torch.float64)) for _ in range(12)]
    for i, v, u in bat[:3]:
        t.update_(i, v); t.sample_uniform(u)
    torch.cuda.synchronize()
    tu, ts = [], []
    for i, v, u in bat[3:]:
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); t.update_(i, v); b.record(); t.sample_uniform(u); c.record(); torch.cuda.synchronize()
        tu.append(a.elapsed_time(b) * 1e3); ts.append(b.elapsed_time(c) * 1e3)
    t.handle.set_profiling(True)
    for i, v, u in bat[3:7]:
        t.update_(i, v)
    torch.cuda.synchronize()
    bd = t.handle.update_breakdown()
    t.handle.set_profiling(False)
    print(f"m=2^{lgm}: update {np.median(tu):8.1f} us (max {max(tu):8.1f})  sample {np.median(ts):7.1f} us  ",
          {k: round(v * 1e3, 1) for k, v in bd.items() if k != "calls"}, flush=True)
t.poll()
