"""Device time of the non-contiguous axis sum (BASELINE config #5: 4096 x 4096, sum(axis=0)) with an L2
flush between calls; prints microseconds and the HBM fraction.  ST_AXIS_BC=32|16|8|-1 overrides the strip
width (-1 = generic one-thread-per-column kernel)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starr_b200 import DeviceSumTree  # noqa: E402

PEAK = 6544.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(0)
    rows = []
    for name, tdt, shape in (("float32", torch.float32, (1, 4096, 4096)), ("int32", torch.int32, (1, 4096, 4096)),
                             ("float64", torch.float64, (1, 4096, 4096)), ("float32", torch.float32, (64, 1024, 256)),
                             ("float32", torch.float32, (1, 16384, 1024))):
        A, R, C = shape
        x = torch.from_numpy(rng.random(A * R * C).astype(np.float64)).to(dev)
        x = (x * 1000).to(tdt)
        trees = [DeviceSumTree(x) for _ in range(4)]          # 4 copies (> L2 together), cycled inside the timed region
        want = x.view(A, R, C).cpu().numpy().sum(axis=1)
        got = trees[0].axis_fold(A, R, C).cpu().numpy().reshape(want.shape)
        ok = got.tobytes() == want.tobytes()
        times = []
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()                         # 8 calls replayed as one graph: no host gaps inside the timing
        with torch.cuda.graph(graph):
            keep = [trees[q & 3].axis_fold(A, R, C) for q in range(8)]
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3 / 8)
        ok = ok and keep[5].cpu().numpy().reshape(want.shape).tobytes() == want.tobytes()
        times.sort()
        us = times[len(times) // 2]
        nbytes = x.numel() * x.element_size() + want.nbytes
        rows.append((name, shape, us, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK, ok))
        del trees, x
    print(f"ST_AXIS_BC={os.environ.get('ST_AXIS_BC', 'auto')}")
    for name, shape, us, gbs, frac, ok in rows:
        print(f"{name:8s} {str(shape):20s} {us:8.1f} us  {gbs:7.0f} GB/s  {frac:5.2f} of {PEAK:.0f}  bits_equal_numpy={ok}")


if __name__ == "__main__":
    main()
