import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from starr_b200 import DeviceSumTree
dev = torch.device("cuda"); nl = 1 << 27
roots = []
for r in range(8):
    g = torch.Generator(device=dev); g.manual_seed(4242 + r)
    t = DeviceSumTree(torch.rand(nl, device=dev, generator=g)); roots.append(t.heap[1].item()); del t
top = DeviceSumTree(torch.tensor(roots, device=dev))
h = top.heap.cpu().numpy()
for i, v in enumerate(h): print(i, repr(float(v)), hex(int(v.view(np.uint32))), float(v) - 2.0 ** round(np.log2(max(v, 1))))
