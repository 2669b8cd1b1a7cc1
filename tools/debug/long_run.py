"""Does the update slow down over a long run?  (uniform priorities on a power-of-two tree put every upper node next to
a binade boundary; nodes whose ulp is about |d| random-walk into it and then hop across it all the time)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
lgn = int(sys.argv[2]) if len(sys.argv) > 2 else 27
lgm = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n, m = 1 << lgn, 1 << lgm
dev = torch.device("cuda")
t = DeviceSumTree(torch.from_numpy(np.random.default_rng(1000).random(n, dtype=np.float32)).to(dev)); t.reserve(m)
gen = torch.Generator(device=dev); gen.manual_seed(20260)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
idx = torch.empty(m, dtype=torch.int64, device=dev); val = torch.empty(m, device=dev)
ts = []
s = torch.cuda.current_stream().cuda_stream
prev = t.handle.update_stats(s)
for k in range(steps):
    idx.random_(0, n, generator=gen); val.uniform_(0, 1, generator=gen)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); t.update_(idx, val); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
    if (k + 1) % 25 == 0:
        st = t.handle.update_stats(s)
        top = t.heap[1:16].cpu().numpy()
        near = [(h + 1, float(v - 2.0 ** round(np.log2(v)))) for h, v in enumerate(top) if abs(v - 2.0 ** round(np.log2(v))) < 300]
        print(f"steps {k-24:4d}..{k:4d}: update median {np.median(ts[-25:]):7.1f} max {max(ts[-25:]):7.1f} us; exact chunks in these steps "
              f"{st['chunks_folded_exactly'] - prev['chunks_folded_exactly']:6d}; nodes < 300 from a power of two: {near}", flush=True)
        prev = st
t.poll()
