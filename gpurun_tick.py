import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from starr_b200 import DeviceSumTree, _lib
rng = np.random.default_rng(0)
n, m = 1 << 20, 4096
t = DeviceSumTree(torch.from_numpy(rng.random(n, dtype=np.float32)).cuda())
t.reserve(m)
for it in range(3):
    idx = torch.from_numpy(rng.integers(0, n, m)).cuda(); val = torch.from_numpy(rng.random(m, dtype=np.float32)).cuda()
    t.update_(idx, val)
torch.cuda.synchronize()
# read status words 32..63 via update_stats trick: copy d_status through st_update_stats (words 4..7 only) -> use raw cudaMemcpy via torch? expose pointer
import ctypes
lib=_lib.lib()
class ST(ctypes.Structure): pass
# st_tree layout: int64 n; int dtype, device, esize, depth; void* heap; bool owns; uint* d_status ...
h = t.handle.ptr
raw = (ctypes.c_ubyte*64).from_address(h.value)
import struct
n_, dtype, device, esize, depth, heap = struct.unpack_from('qiiiiP', bytes(raw), 0)
owns, = struct.unpack_from('?', bytes(raw), 32)
d_status, = struct.unpack_from('P', bytes(raw), 40)
buf = torch.empty(64, dtype=torch.int32, device='cuda')
cudart = ctypes.CDLL('libcudart.so.12') if False else None
import torch.cuda
# use torch to copy from raw pointer via cudaMemcpy in libstarr? simpler: ctypes to libcudart from torch
rt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), 'lib', 'libcudart.so.12')) if os.path.exists(os.path.join(os.path.dirname(torch.__file__), 'lib', 'libcudart.so.12')) else ctypes.CDLL('libcudart.so')
host = (ctypes.c_uint*64)()
rt.cudaMemcpy(host, ctypes.c_void_p(d_status), 256, 2)
ticks = list(host)[32:60]
print('ticks', ticks)
names = ['valid','sort','leaf']+[f'L{l}' for l in range(21)]
prev=0
for nme, tk in zip(names, ticks):
    print(f'{nme:6s} {tk:8d} (+{tk-prev})'); prev=tk
