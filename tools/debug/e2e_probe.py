"""Where does the end-to-end step lose ~0.1 ms against the device-resident step?  (a) the device step alone, (b) the same
with independent H2D + D2H traffic of the step's size on side streams, (c) the C-ABI host pipeline with 2/3/4 slots."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree, _lib
n, m = 1 << 27, 1 << 20
dev = torch.device("cuda")
t = DeviceSumTree(torch.rand(n, device=dev)); t.reserve(m)
K = 24
bat = [(torch.randint(0, n, (m,), device=dev), torch.rand(m, device=dev), torch.rand(m, device=dev, dtype=torch.float64)) for _ in range(K + 4)]
out = torch.empty(m, dtype=torch.int64, device=dev)
hin = torch.empty(20 << 20, dtype=torch.uint8).pin_memory(); din = torch.empty(20 << 20, dtype=torch.uint8, device=dev)
hout = torch.empty(8 << 20, dtype=torch.uint8).pin_memory(); dout = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def loop(traffic):
    for i, v, u in bat[:4]:
        t.update_(i, v); t.sample_uniform(u, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i, v, u in bat[4:]:
        if traffic:
            with torch.cuda.stream(s_in): din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s_out): hout.copy_(dout, non_blocking=True)
        t.update_(i, v); t.sample_uniform(u, out=out)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K
print("device step alone            %.3f ms" % loop(False))
print("device step + PCIe traffic   %.3f ms" % loop(True))
print("device step alone            %.3f ms" % loop(False))
h = t.handle
rng = np.random.default_rng(1)
pins = []
for _ in range(8):
    pb = _lib.PinnedBuffer(m * 20)
    i = pb.array(np.int64, m, 0); v = pb.array(np.float32, m, 8 * m); u = pb.array(np.float64, m, 12 * m)
    i[:] = rng.integers(0, n, m); v[:] = rng.random(m, dtype=np.float32); u[:] = rng.random(m)
    pins.append((pb, i, v, u))
for NS in (2, 3, 4):
    outs = []
    for _ in range(NS):
        pb = _lib.PinnedBuffer(m * 8); outs.append((pb, pb.array(np.int64, m, 0)))
    def run(first, nsteps):
        for k in range(nsteps):
            slot = k % NS
            if k >= NS: h.per_step_host_wait(slot)
            _, i, v, u = pins[(first + k) % len(pins)]
            v[:8] = rng.random(8, dtype=np.float32)
            h.per_step_host_submit(slot, i, v, u, outs[slot][1])
        for k in range(max(0, nsteps - NS), nsteps): h.per_step_host_wait(k % NS)
    run(0, NS); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(NS, 40); dt = (time.perf_counter() - t0) / 40 * 1e3
    print("host pipeline, %d slots       %.3f ms" % (NS, dt))
