"""``ShardedSumTree``: one logical sum tree whose leaf range is sharded by subtree over the
GPUs of one node (one process per GPU, ``torch.distributed``; NCCL over NVLink on the B200 box).

The reference is single-process (no collective anywhere), so there is no reference
file to cite for the exchange itself; what IS pinned is the result: tree contents, leaves and
sampled indices are bit-identical to ONE reference tree over all ``n`` leaves
(``starr/src/_sumtree_array.pyx:22-122``).

Geometry: n = 2^k leaves, G = 2^g ranks.  Rank r owns leaves [r*n/G, (r+1)*n/G) = the
heap subtree rooted at global node G + r.  Global nodes 1 .. G-1 ("top tree") plus the G shard
roots are replicated on every rank as a tiny tree ``top`` with G leaves: top heap index h IS the
global heap index h.

Per batch (inputs are replicated on every rank, as a PER learner broadcasts them):
  build   local pairwise build -> all-gather of the G shard roots -> pairwise top build.
  update  every rank applies the updates it owns (batch order preserved) and gets their
          differences d_j; the d_j are exchanged (sum all-reduce of a zero-filled vector: one
          non-zero contributor per slot, hence exact); every rank folds all d_j into the top tree
          in batch order (the reference adds EVERY update's d to the top nodes in order, so
          roots alone are not enough for floats); shard roots are re-gathered.
  sample  every rank routes all queries through the top tree (same compares / subtractions as
          the first g levels of the reference descent), finishes the ones it owns locally with
          the left-over prefix, and the global indices are combined with one all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _hostlogic


def _default_factory(n, dtype, device):
    from .device_tree import DeviceSumTree
    return DeviceSumTree(n, dtype=dtype, device=device)


class ShardedSumTree:
    def __init__(self, n_global: int, dtype=torch.float32, local_leaves: torch.Tensor | None = None,
                 group=None, device=None, tree_factory=None):
        if not dist.is_initialized():
            raise RuntimeError("ShardedSumTree needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world < 2:
            raise ValueError("ShardedSumTree needs >= 2 ranks; use DeviceSumTree on one GPU")
        self.n = int(n_global)
        self.n_local, self.g = _hostlogic.shard_geometry(self.n, self.world)
        self.shift = self.n_local.bit_length() - 1
        self.dtype = dtype
        factory = tree_factory or _default_factory
        if device is None and tree_factory is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.local = factory(self.n_local, dtype, self.device)
        self.top = factory(self.world, dtype, self.device)
        self.offset = self.rank * self.n_local
        if local_leaves is not None:
            self.build_(local_leaves)

    # ---- helpers --------------------------------------------------------------------------------
    def reserve(self, max_global_batch: int) -> None:
        self.local.reserve(max_global_batch)
        self.top.reserve(max_global_batch)

    def poll(self) -> None:
        self.local.poll()
        self.top.poll()

    def _gather_roots(self) -> None:
        """All-gather of the G shard roots (G elements) into the top tree's leaves."""
        root = self.local.heap[1:2].contiguous()
        roots = torch.empty(self.world, dtype=self.dtype, device=root.device)
        dist.all_gather_into_tensor(roots, root, group=self.group)
        self.top.leaves.copy_(roots)

    def total(self) -> torch.Tensor:
        return self.top.heap[1]

    # ---- build ----------------------------------------------------------------------------------
    def build_(self, local_leaves: torch.Tensor | None = None) -> "ShardedSumTree":
        self.local.build_(local_leaves)
        self._gather_roots()
        self.top.build_()      # nodes G-1 .. 1 pairwise from the shard roots: same adds as one tree
        return self

    # ---- update ---------------------------------------------------------------------------------
    def update_(self, idx: torch.Tensor, vals: torch.Tensor, check: bool = False) -> "ShardedSumTree":
        """Replicated batch: ``idx`` are GLOBAL leaf ids, ``vals`` cycles over the global batch
        position (reference pyx:44)."""
        idx = idx.reshape(-1)
        vals = vals.reshape(-1)
        ni, nv = idx.numel(), vals.numel()
        if check:
            if nv and bool((vals < 0).any()):
                raise ValueError("vals must be non-negative")
            if ni and (int(idx.min()) < 0 or int(idx.max()) >= self.n):
                raise IndexError("Out of bounds on buffer access (axis 0)")
        if ni == 0:
            return self
        if nv == 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        shard = idx >> self.shift
        mine = torch.nonzero(shard == self.rank).reshape(-1)        # batch order preserved
        loc_idx = idx[mine] - self.offset
        loc_vals = vals[mine] if nv >= ni else vals[mine % nv]
        d_all = torch.zeros(ni, dtype=self.dtype, device=idx.device)
        if mine.numel():
            d_all[mine] = self.local.update_with_deltas_(loc_idx, loc_vals)
        dist.all_reduce(d_all, op=dist.ReduceOp.SUM, group=self.group)   # one contributor per slot
        self._fold_top(shard, d_all)                                    # ordered fold, all ranks alike
        self._gather_roots()
        return self

    def _fold_top(self, shard: torch.Tensor, d_all: torch.Tensor) -> None:
        """Ordered fold of ALL differences into the replicated top tree.  With the CUDA tree the
        chunk summaries (the parallel part) are split over the ranks and exchanged with one small
        all-gather; the stitch is replicated.  Results are identical on every rank."""
        top = self.top
        ni = shard.numel()
        if not hasattr(top, "top_fold_maps_") or ni <= 4096 or not self.dtype.is_floating_point:
            top.apply_deltas_(shard, d_all)
            return
        nodes = self.world - 1
        rb = top.top_fold_record_bytes()
        nchunks = (ni + 1023) // 1024
        per = (nchunks + self.world - 1) // self.world
        recs = torch.empty(per * self.world * nodes * rb, dtype=torch.uint8, device=shard.device)
        lo = self.rank * per
        shard = shard.contiguous()
        d_all = d_all.contiguous()
        top.top_fold_maps_(shard, d_all, lo, lo + per, recs)
        slab = recs[lo * nodes * rb:(lo + per) * nodes * rb].clone()
        dist.all_gather_into_tensor(recs, slab, group=self.group)
        top.top_fold_apply_(shard, d_all, recs)

    # ---- sample ---------------------------------------------------------------------------------
    def sample_uniform(self, u: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Global leaf indices for replicated uniforms ``u`` (float64), identical to the
        reference's ``get_prefix_sum_id((root * u).astype(dtype))`` on the whole tree."""
        u = u.reshape(-1)
        shard, resid = self.top.route_uniform(u)
        mine = torch.nonzero(shard == self.rank).reshape(-1)
        if out is None:
            out = torch.empty(u.numel(), dtype=torch.int64, device=u.device)
        out.zero_()
        if mine.numel():
            out[mine] = self.local.search(resid[mine]) + self.offset
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- inspection -------------------------------------------------------------------------------
    def gather_sumtree(self) -> torch.Tensor:
        """The whole tree in the reference's flat layout (``sumtree()``), assembled on every rank:
        level l of the global tree = concatenation of level l-g of the shards."""
        parts = [self.top.heap[: self.world]]
        level = 0
        while (1 << level) < self.n_local:
            mine = self.local.heap[(1 << level): (2 << level)].contiguous()
            allp = torch.empty(self.world * mine.numel(), dtype=self.dtype, device=mine.device)
            dist.all_gather_into_tensor(allp, mine, group=self.group)
            parts.append(allp)
            level += 1
        return torch.cat(parts)

    def gather_leaves(self) -> torch.Tensor:
        mine = self.local.leaves.contiguous()
        allp = torch.empty(self.n, dtype=self.dtype, device=mine.device)
        dist.all_gather_into_tensor(allp, mine, group=self.group)
        return allp
