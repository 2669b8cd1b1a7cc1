import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from starr_b200 import SumTreeArray
def t(fn, reps=200):
    fn(); fn()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter()-t0)/reps*1e6
x = SumTreeArray(np.ones(1_000_000))
st = x._state; h = st.handle
idx = np.arange(999990, 1000000, dtype=np.intp); vals = np.full(10, 2.0)
u = np.random.rand(100); out = np.zeros(100, np.intp)
print('x[-10:] = 2                 ', t(lambda: x.__setitem__(slice(-10, None), 2)))
print('x.sample(100)               ', t(lambda: x.sample(100)))
print('x.sum()                     ', t(lambda: x.sum()))
print('raw update_gather_host(10)  ', t(lambda: h.update_gather_host(idx, vals)))
print('raw update_host(10)         ', t(lambda: h.update_host(idx, vals)))
print('raw sample_root_host(100)   ', t(lambda: h.sample_root_host(u, out)))
print('raw root()                  ', t(lambda: h.root()))
print('python: _indices[slice]     ', t(lambda: x._indices[slice(-10, None)]))
print('x.sum(axis=0) n/a 1d; get_prefix_sum_id(100)', t(lambda: x.get_prefix_sum_id(u * 1e6)))
