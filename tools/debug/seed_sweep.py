"""1-GPU: update breakdown for the leaf seeds the sharded bench gives each rank.  With uniform [0,1)
priorities and a power-of-two leaf count EVERY node's expected value is a power of two, so some node
always hovers at a binade boundary; this shows how much the huge-segment stitch pays for it."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from starr_b200 import DeviceSumTree
dev = torch.device("cuda", 0)
n, m = 1 << 27, 1 << 20
for seed in [0] + list(range(1000, 1008)):
    leaves = np.random.default_rng(seed).random(n, dtype=np.float32)
    tree = DeviceSumTree(torch.from_numpy(leaves).to(dev)); tree.reserve(m)
    rng = np.random.default_rng(5)
    B = [(torch.from_numpy(rng.integers(0, n, m)).to(dev), torch.from_numpy(rng.random(m, dtype=np.float32)).to(dev)) for _ in range(6)]
    for k in range(60): tree.update_(*B[k % 6])          # also brings the clocks up
    tree.handle.set_profiling(True)
    for k in range(12): tree.update_(*B[k % 6])
    torch.cuda.synchronize()
    last = tree.handle.update_breakdown()
    st = tree.handle.update_stats()
    print(f"seed {seed}: huge_maps {last['huge_maps_ms']*1e3:.0f} us first stitch {last['huge_stitch_ms']*1e3:.0f} us second stitch + folds {last['fold_scan_ms']*1e3:.0f} us "
          f"(avg over {last['calls']} calls); 72 calls: {st['huge_segments_with_exact_chunks']} of "
          f"{st['huge_segments_with_exact_chunks'] + st['huge_segments_maps_only']} huge segments needed {st['chunks_folded_exactly']} exact chunk folds")
    del tree
