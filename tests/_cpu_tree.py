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

    def shard_plan(self, ids, shift, shards, rank, payload, check_negative, want_ids=True):
        ids = ids.reshape(-1)
        payload = payload.reshape(-1)
        owner = ids >> shift
        if bool(((ids < 0) | (owner >= shards)).any()):
            raise IndexError("Out of bounds on buffer access (axis 0)")
        if check_negative and bool((payload < 0).any()):
            raise ValueError("vals must be non-negative")
        slot = torch.empty(ids.numel(), dtype=torch.int32)
        counts = []
        for s in range(shards):
            pos = torch.nonzero(owner == s).reshape(-1)
            slot[pos] = torch.arange(pos.numel(), dtype=torch.int32)
            counts.append(int(pos.numel()))
        mine = torch.nonzero(owner == rank).reshape(-1)
        own_ids = (ids[mine] - (rank << shift)) if want_ids else None
        return owner.to(torch.uint8), slot, own_ids, payload[mine % payload.numel()], counts

    def shard_expand_values(self, shard, slot, slabs):
        return slabs[shard.long(), slot.long()]

    def shard_expand_indices(self, shard, slot, slabs, leaves_per_shard, out=None):
        res = shard.long() * leaves_per_shard + slabs[shard.long(), slot.long()].long()
        if out is None:
            return res
        out.copy_(res)
        return out

    def update_with_deltas_(self, idx, vals, out=None):
        idx_np = idx.numpy().astype(np.intp)
        vals_np = vals.numpy()
        before = self._np[self.n:].copy()
        d = np.empty(idx_np.size, self._np.dtype)
        for j, l in enumerate(idx_np):            # the leaf chain of pyx:43-45
            d[j] = vals_np[j % vals_np.size] - before[l]
            before[l] = before[l] + d[j]
        oracle.c.update_sumtree(idx_np, np.ascontiguousarray(vals_np), self._np[self.n:], self._np[:self.n])
        assert before.tobytes() == self._np[self.n:].tobytes()
        if out is not None:
            out.copy_(torch.from_numpy(d))
            return out
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
