"""TEST HELPER ONLY: a CPU stand-in for DeviceSumTree built on the oracle, so that the host-side
logic of ShardedSumTree (ownership, value cycling, delta exchange, routing, result combination)
can be exercised under the gloo backend on a machine without GPUs.  Never used by the product."""
import numpy as np
import torch

import oracle


class CpuTree:
    def __init__(self, n, dtype=torch.float32, device=None):
        self.n, self.dtype = int(n), dtype
        self.heap = torch.zeros(2 * self.n, dtype=dtype)
        self._np = self.heap.numpy()          # shares memory

    @property
    def leaves(self):
        return self.heap[self.n:]

    @property
    def tree(self):
        return self.heap[:self.n]

    def reserve(self, m):
        pass

    def poll(self):
        pass

    def build_(self, leaves=None):
        if leaves is not None:
            self.heap[self.n:] = leaves.reshape(-1).to(self.dtype)
        oracle.c.build_sumtree_from_array(self._np[self.n:], self._np[:self.n])
        return self

    def update_with_deltas_(self, idx, vals):
        idx_np = idx.numpy().astype(np.intp)
        vals_np = vals.numpy()
        before = self._np[self.n:].copy()
        d = np.empty(idx_np.size, self._np.dtype)
        for j, l in enumerate(idx_np):            # the leaf chain of pyx:43-45
            d[j] = vals_np[j % vals_np.size] - before[l]
            before[l] = before[l] + d[j]
        oracle.c.update_sumtree(idx_np, np.ascontiguousarray(vals_np), self._np[self.n:], self._np[:self.n])
        assert before.tobytes() == self._np[self.n:].tobytes()
        return torch.from_numpy(d)

    def apply_deltas_(self, idx, deltas):
        tree = self._np[:self.n]
        d = deltas.numpy()
        for j, l in enumerate(idx.numpy()):
            h = (int(l) + self.n) // 2
            while h > 0:
                tree[h] = tree[h] + d[j]
                h //= 2
        return self

    def route_uniform(self, u):
        root = self._np[1]
        v = (root * u.numpy()).astype(self._np.dtype)
        h = np.ones(v.size, np.int64)
        for _ in range(64):
            act = h < self.n
            if not act.any():
                break
            c = np.where(act, 2 * h, 0)
            lv = self._np[c]
            right = act & (v >= lv)
            v = np.where(right, (v - lv).astype(v.dtype), v)
            h = np.where(act, c + right, h)
        return torch.from_numpy(h - self.n), torch.from_numpy(v)

    def search(self, vals):
        out = np.zeros(vals.numel(), np.intp)
        oracle.c.get_prefix_sum_idx(out, np.ascontiguousarray(vals.numpy()), self._np[self.n:], self._np[:self.n])
        return torch.from_numpy(out.astype(np.int64))
