"""``DeviceSumTree``: the batched device entry point for Prioritized-Experience-Replay loops.

Torch tensors in, torch tensors out, everything enqueued on the current CUDA stream, no host
synchronisation unless asked for.  torch is plumbing only (device memory + streams): the tensors'
raw pointers go straight into the C ABI of ``libstarr_b200.so``.

It replaces the same reference call sites as ``SumTreeArray`` (``starr/sumtree_array.py:185,
233, 354, 508`` and ``sample`` ``:417-421``) for callers whose indices / priorities / uniforms
already live on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_TORCH_TO_NP = {
    torch.int16: np.int16, torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8,
    torch.float32: np.float32, torch.float64: np.float64,
}
for _n, _np in (("uint16", np.uint16), ("uint32", np.uint32), ("uint64", np.uint64)):
    if hasattr(torch, _n):
        _TORCH_TO_NP[getattr(torch, _n)] = _np
_NP_TO_TORCH = {np.dtype(v): k for k, v in _TORCH_TO_NP.items()}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class DeviceSumTree:
    """A sum tree over ``n`` leaves resident on one GPU.

    ``heap`` is one tensor of ``2n`` elements: ``heap[:n]`` is the tree in the reference's layout
    (``[0, root, node 2, ...]``, what ``SumTreeArray.sumtree()`` returns) and ``heap[n:]`` are
    the leaves.  Both are exposed as zero-copy views (``tree`` / ``leaves``).
    """

    def __init__(self, n_or_leaves, dtype=None, device=None):
        if isinstance(n_or_leaves, torch.Tensor):
            init = n_or_leaves
            n = init.numel()
            dtype = dtype or init.dtype
            device = device if device is not None else (init.device if init.is_cuda else None)
        else:
            init, n = None, int(n_or_leaves)
            dtype = dtype or torch.float32
        if dtype not in _TORCH_TO_NP:
            raise TypeError("No matching signature found")
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceSumTree needs a CUDA device; starr_b200 has no CPU fallback")
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("DeviceSumTree needs a CUDA device; starr_b200 has no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = n
        self.dtype = dtype
        with torch.cuda.device(self.device):
            self.heap = torch.zeros(2 * n, dtype=dtype, device=self.device)
            torch.cuda.current_stream().synchronize()
            self._h = _lib.TreeHandle(n, _TORCH_TO_NP[dtype], self.device.index,
                                      external_heap=self.heap.data_ptr())
        if init is not None:
            self.build_(init)

    # ---- views ------------------------------------------------------------------------------
    @property
    def leaves(self) -> torch.Tensor:
        return self.heap[self.n:]

    @property
    def tree(self) -> torch.Tensor:
        return self.heap[:self.n]

    def total(self) -> torch.Tensor:
        """0-d view of the root (``sumtree[1]``); stays on the device."""
        return self.heap[1]

    @property
    def handle(self) -> _lib.TreeHandle:
        return self._h

    def reserve(self, max_batch: int) -> None:
        """Pre-allocate the update workspace (call before CUDA-graph capture)."""
        self._h.reserve(max_batch)

    def _chk(self, t: torch.Tensor, dtype, name: str) -> torch.Tensor:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device):
            raise ValueError(f"{name} must be a CUDA tensor on {self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
        return t.contiguous()

    # ---- build ------------------------------------------------------------------------------
    def build_(self, leaves: torch.Tensor | None = None) -> "DeviceSumTree":
        """Validate (no negative leaf) and rebuild every node pairwise.  Synchronises once to
        report the validation result (``ValueError`` like the reference)."""
        with torch.cuda.device(self.device):
            if leaves is not None:
                if leaves.numel() != self.n:
                    raise TypeError("array and sumtree must have the same size")
                src = leaves.to(device=self.device, dtype=self.dtype).reshape(-1).contiguous()
                self._h.build_from_device(src.data_ptr(), _stream())
            else:
                self._h.build(_stream())
        return self

    # ---- update -----------------------------------------------------------------------------
    def update_(self, idx: torch.Tensor, vals: torch.Tensor, check: bool = False,
                assign_leaf: bool = False) -> "DeviceSumTree":
        """``leaves[idx[j]] = vals[j % len(vals)]`` for j in batch order, exact reference
        semantics (duplicates apply sequentially; float ancestors are ordered folds).

        Asynchronous by default: a batch that fails validation (negative value, index out of
        range) is skipped as a whole on the device and reported by the next ``poll()``.
        ``check=True`` waits and raises immediately.  A failed batch does not block the batches
        enqueued behind it.  ``assign_leaf=True`` stores ``vals[j]`` itself in the leaf (the
        reference's experimental C++ twin, ``starr/experimental/src/sumtree_array.cpp:19-23``)
        instead of ``fl(leaf + fl(vals[j] - leaf))``."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        with torch.cuda.device(self.device):
            self._h.update_device(idx.data_ptr(), idx.numel(), vals.data_ptr(), vals.numel(),
                                  _stream(), sync_check=check, assign_leaf=assign_leaf)
        return self

    def append_(self, vals: torch.Tensor) -> "DeviceSumTree":
        """Ring-buffer append (reference ``starr/experimental/slow.py:81-88`` ``CircularSumTree``):
        ``vals[k]`` goes to leaf ``(cursor + k) % n`` in order, the cursor advances by ``len(vals)``."""
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        k = vals.numel()
        if k == 0:
            return self
        cursor = getattr(self, "_cursor", 0)
        idx = (torch.arange(k, dtype=torch.int64, device=self.device) + cursor) % self.n
        self.update_(idx, vals)
        self._cursor = (cursor + k) % self.n
        return self

    # ---- in-place leaf arithmetic (reference sumtree_array.py:150-173, :202-205) --------
    def inplace_(self, op: str, operand, rebuild: bool = True) -> "DeviceSumTree":
        """``leaves = leaves (op) operand`` on the device, ``op`` in add / subtract / multiply /
        true_divide / fill, ``operand`` a Python or NumPy scalar or a tensor of ``n`` elements of the
        tree dtype; then (``rebuild``) validate and rebuild the tree like the reference's in-place
        ufuncs do (``ValueError`` if a leaf went negative: the leaves stay changed, the tree does not)."""
        with torch.cuda.device(self.device):
            if isinstance(operand, torch.Tensor):
                y = self._chk(operand.reshape(-1), self.dtype, "operand")
                if y.numel() != self.n:
                    raise TypeError("operand must have the size of the array")
                self._h.inplace_device(op, None, y.data_ptr(), y.numel(), _stream())
            else:
                sc = np.empty(1, dtype=_TORCH_TO_NP[self.dtype])
                sc.fill(operand)
                self._h.inplace_device(op, sc, 0, 0, _stream())
            if rebuild:
                self._h.build(_stream())
        return self

    # ---- building blocks of the sharded tree (starr_b200/sharded.py) ----------------------------
    def update_with_deltas_(self, idx: torch.Tensor, vals: torch.Tensor,
                            out: torch.Tensor | None = None) -> torch.Tensor:
        """``update_`` that also returns every update's applied difference d_j (batch order)."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        if out is None:
            out = torch.empty(idx.numel(), dtype=self.dtype, device=self.device)
        d = self._chk(out, self.dtype, "out")
        if d.numel() != idx.numel() or not out.is_contiguous():
            raise TypeError("out must be contiguous and have the size of idx")
        with torch.cuda.device(self.device):
            self._h.update_deltas_device(idx.data_ptr(), idx.numel(), vals.data_ptr(), vals.numel(),
                                         d.data_ptr(), _stream())
        return d

    def update_begin_(self, idx: torch.Tensor, vals: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """First half of ``update_with_deltas_``: through the leaf pass -- ``out`` holds every d_j in stream
        order behind this call, the ancestors are untouched until ``update_finish_``."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        if self._chk(out, self.dtype, "out") is not out or out.numel() < idx.numel():
            raise TypeError("out must be contiguous and at least as long as idx")
        with torch.cuda.device(self.device):
            self._h.update_begin_device(idx.data_ptr(), idx.numel(), vals.data_ptr(), vals.numel(), out.data_ptr(), _stream())
        return out

    def update_finish_(self) -> "DeviceSumTree":
        with torch.cuda.device(self.device):
            self._h.update_finish_device(_stream())
        return self

    def route_compact(self, u: torch.Tensor, rank: int, own_resid: torch.Tensor, own_pos: torch.Tensor,
                      own_count: torch.Tensor) -> None:
        """On a top tree: route every uniform through the tree and append (left-over prefix, batch position) of
        the queries that end in leaf ``rank`` to ``own_resid`` / ``own_pos``; ``own_count`` (uint32[1]) counts them."""
        u = self._chk(u.reshape(-1), torch.float64, "u")
        if own_resid.dtype != self.dtype or own_pos.dtype != torch.int32 or min(own_resid.numel(), own_pos.numel()) < u.numel():
            raise TypeError("own_resid / own_pos must have the tree dtype / int32 and the size of u")
        with torch.cuda.device(self.device):
            self._h.route_compact_device(u.data_ptr(), u.numel(), rank, own_resid.data_ptr(), own_pos.data_ptr(),
                                         own_count.data_ptr(), _stream())

    def search_push(self, prefix_sums: torch.Tensor, count: torch.Tensor, pos: torch.Tensor, index_offset: int, peer,
                    payload_offset: int, span_elems: int) -> None:
        """``search`` over the first ``count[0]`` prefixes; result k (``index_offset`` + leaf) is stored as int64
        at element ``pos[k]`` of the region at ``payload_offset`` of EVERY rank's exchange buffer."""
        vals = self._chk(prefix_sums, self.dtype, "prefix_sums")
        with torch.cuda.device(self.device):
            self._h.search_push_device(vals.data_ptr(), vals.numel(), count.data_ptr(), pos.data_ptr(), index_offset, peer,
                                       payload_offset, span_elems, _stream())

    def search_pairs(self, prefix_sums: torch.Tensor, count: torch.Tensor, pos: torch.Tensor, pairs: torch.Tensor) -> None:
        """``search`` over the first ``count[0]`` prefixes; result k goes to ``pairs[k] = (pos[k], local leaf)`` (int32 [m, 2])."""
        vals = self._chk(prefix_sums, self.dtype, "prefix_sums")
        if pairs.dtype != torch.int32 or not pairs.is_contiguous() or pairs.numel() < 2 * vals.numel() or pos.dtype != torch.int32:
            raise TypeError("pairs must be a contiguous int32 tensor with two entries per prefix, pos an int32 tensor")
        with torch.cuda.device(self.device):
            self._h.search_pairs_device(vals.data_ptr(), vals.numel(), count.data_ptr(), pos.data_ptr(), pairs.data_ptr(), _stream())

    def shard_expand_cat(self, shard: torch.Tensor, slot: torch.Tensor, cat: torch.Tensor, starts: list,
                         out: torch.Tensor | None = None) -> torch.Tensor:
        """``out[j] = cat[starts[shard[j]] + slot[j]]`` (the shards' slabs stored back to back)."""
        cat = self._chk(cat, self.dtype, "cat")
        out = self._chk_out(out, shard.numel(), self.dtype)
        with torch.cuda.device(self.device):
            self._h.shard_expand_cat_device(shard.data_ptr(), slot.data_ptr(), shard.numel(), cat.data_ptr(), starts,
                                            out.data_ptr(), _stream())
        return out[:shard.numel()]

    def apply_deltas_(self, idx: torch.Tensor, deltas: torch.Tensor) -> "DeviceSumTree":
        """For j in batch order add ``deltas[j]`` to every proper ancestor of leaf ``idx[j]``
        (ordered folds for floats); the leaves themselves are not touched."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        deltas = self._chk(deltas.reshape(-1), self.dtype, "deltas")
        if idx.numel() != deltas.numel():
            raise TypeError("idx and deltas must have the same size")
        with torch.cuda.device(self.device):
            self._h.apply_deltas_device(idx.data_ptr(), idx.numel(), deltas.data_ptr(), _stream())
        return self

    def top_fold_record_bytes(self) -> int:
        return self._h.top_fold_record_bytes()

    def top_fold_maps_(self, leaf: torch.Tensor, deltas: torch.Tensor, chunk_lo: int, chunk_hi: int,
                       recs: torch.Tensor) -> None:
        """Phase A of ``apply_deltas_`` on a tiny tree: summaries of the 1024-update chunks
        [chunk_lo, chunk_hi) for every internal node, written into ``recs`` ([chunk][node-1] records).
        ``leaf`` is int64 or uint8."""
        if leaf.dtype not in (torch.int64, torch.uint8) or not leaf.is_contiguous():
            raise TypeError("leaf must be a contiguous int64 or uint8 tensor")
        with torch.cuda.device(self.device):
            self._h.top_fold_maps_device(leaf.data_ptr(), leaf.element_size(), deltas.data_ptr(), leaf.numel(),
                                         chunk_lo, chunk_hi, recs.data_ptr(), _stream())

    def top_fold_apply_(self, leaf: torch.Tensor, deltas: torch.Tensor, recs: torch.Tensor) -> None:
        """Phase B: stitch the chunk summaries of ALL chunks into the internal nodes."""
        if leaf.dtype not in (torch.int64, torch.uint8) or not leaf.is_contiguous():
            raise TypeError("leaf must be a contiguous int64 or uint8 tensor")
        with torch.cuda.device(self.device):
            self._h.top_fold_apply_device(leaf.data_ptr(), leaf.element_size(), deltas.data_ptr(), leaf.numel(),
                                          recs.data_ptr(), _stream())

    def shard_plan(self, ids: torch.Tensor, shift: int, shards: int, rank: int, payload: torch.Tensor,
                   check_negative: bool, want_ids: bool = True, out=None, result_slot: int = 0, wait: bool = True,
                   own_pos: torch.Tensor | None = None):
        """Plan of one replicated batch of a sharded tree (``st_shard_plan_device``): returns
        ``(shard uint8[m], slot int32[m], own_ids int64[>=m] | None, own_payload dtype[>=m], counts)``;
        only the first ``counts[rank]`` entries of the ``own_*`` tensors are meaningful.  Waits for
        the current stream's plan kernels (the counts are needed on the host) and raises
        ``IndexError`` / ``ValueError`` for a bad batch.  ``out``: the four tensors to write into
        (capacity >= m).  ``wait=False`` only enqueues (``counts`` is None): collect the counts with
        ``shard_plan_wait`` after enqueuing whatever needs only the device-side counts."""
        ids = self._chk(ids.reshape(-1), torch.int64, "ids")
        payload = self._chk(payload.reshape(-1), self.dtype, "payload")
        m = ids.numel()
        if out is None:
            out = (torch.empty(m, dtype=torch.uint8, device=self.device),
                   torch.empty(m, dtype=torch.int32, device=self.device),
                   torch.empty(m, dtype=torch.int64, device=self.device) if want_ids else None,
                   torch.empty(m, dtype=self.dtype, device=self.device))
        shard, slot, own_ids, own_payload = out
        if min(shard.numel(), slot.numel(), own_payload.numel()) < m or (want_ids and own_ids.numel() < m):
            raise TypeError("plan buffers are smaller than the batch")
        if own_pos is not None and (own_pos.dtype != torch.int32 or own_pos.numel() < m or not own_pos.is_contiguous()):
            raise TypeError("own_pos must be a contiguous int32 tensor of the batch size")
        with torch.cuda.device(self.device):
            counts = self._h.shard_plan_pos_device(ids.data_ptr(), m, shift, shards, rank, payload.data_ptr(),
                                                   payload.numel(), check_negative, shard.data_ptr(), slot.data_ptr(),
                                                   own_ids.data_ptr() if want_ids else 0, own_payload.data_ptr(),
                                                   own_pos.data_ptr() if own_pos is not None else 0,
                                                   result_slot, _stream(), wait)
        return shard[:m], slot[:m], own_ids, own_payload, counts

    def shard_plan_wait(self, result_slot: int, shards: int) -> list:
        with torch.cuda.device(self.device):
            return self._h.shard_plan_wait(result_slot, shards)

    def search_counted(self, prefix_sums: torch.Tensor, result_slot: int, which: int, out: torch.Tensor) -> None:
        """``search`` over the first ``counts[which]`` entries of ``prefix_sums``, the count being read
        on the device from plan result slot ``result_slot`` (the host need not know it yet)."""
        vals = self._chk(prefix_sums, self.dtype, "prefix_sums")
        if self._chk(out, torch.int64, "out") is not out or out.numel() < vals.numel():
            raise TypeError("out must be a contiguous int64 tensor at least as long as prefix_sums")
        with torch.cuda.device(self.device):
            self._h.search_counted_device(vals.data_ptr(), vals.numel(),
                                          self._h.shard_plan_counts_ptr(result_slot) + 4 * which, out.data_ptr(),
                                          _stream())

    def shard_expand_values(self, shard: torch.Tensor, slot: torch.Tensor, slabs: torch.Tensor) -> torch.Tensor:
        """``out[j] = slabs[shard[j], slot[j]]`` for 2-D ``slabs`` of the tree's dtype."""
        slabs = self._chk(slabs, self.dtype, "slabs")
        out = torch.empty(shard.numel(), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._h.shard_expand_values_device(shard.data_ptr(), slot.data_ptr(), shard.numel(), slabs.data_ptr(),
                                               slabs.shape[1], out.data_ptr(), _stream())
        return out

    def shard_expand_indices(self, shard: torch.Tensor, slot: torch.Tensor, slabs: torch.Tensor,
                             leaves_per_shard: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """``out[j] = shard[j] * leaves_per_shard + slabs[shard[j], slot[j]]`` (int32 slabs, int64 out)."""
        slabs = self._chk(slabs, torch.int32, "slabs")
        if out is None:
            out = torch.empty(shard.numel(), dtype=torch.int64, device=self.device)
        if self._chk(out, torch.int64, "out") is not out or out.numel() != shard.numel():
            raise TypeError("out must be a contiguous int64 tensor of the batch size")
        with torch.cuda.device(self.device):
            self._h.shard_expand_indices_device(shard.data_ptr(), slot.data_ptr(), shard.numel(), slabs.data_ptr(),
                                                slabs.shape[1], leaves_per_shard, out.data_ptr(), _stream())
        return out

    def route_uniform(self, u: torch.Tensor):
        """``sample_uniform`` that also returns the prefix left over at the leaf reached."""
        u = self._chk(u.reshape(-1), torch.float64, "u")
        m = u.numel()
        leaf = torch.empty(m, dtype=torch.int64, device=self.device)
        resid = torch.empty(m, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._h.route_uniform_device(u.data_ptr(), m, leaf.data_ptr(), resid.data_ptr(), _stream())
        return leaf, resid

    def poll(self) -> None:
        """Wait for the current stream and raise the first deferred validation error."""
        with torch.cuda.device(self.device):
            self._h.poll_status(_stream(), clear=True)

    # ---- search / sample --------------------------------------------------------------------
    def _chk_out(self, out: torch.Tensor | None, m: int, dtype=torch.int64, name: str = "out") -> torch.Tensor:
        """A caller-supplied result buffer goes to the kernel as a raw pointer: it must be the right
        device / dtype, contiguous and long enough."""
        if out is None:
            return torch.empty(m, dtype=dtype, device=self.device)
        if self._chk(out, dtype, name) is not out or out.numel() < m:
            raise TypeError(f"{name} must be a contiguous {dtype} tensor on {self.device} with at least {m} elements")
        return out

    def search(self, prefix_sums: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        vals = self._chk(prefix_sums.reshape(-1), self.dtype, "prefix_sums")
        if out is not None:
            self._chk_out(out, vals.numel())
            with torch.cuda.device(self.device):
                self._h.search_device(vals.data_ptr(), vals.numel(), out.data_ptr(), _stream())
            return out
        out = torch.empty(vals.numel(), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._h.search_device(vals.data_ptr(), vals.numel(), out.data_ptr(), _stream())
        return out.reshape(prefix_sums.shape)

    def sample_uniform(self, u: torch.Tensor, out: torch.Tensor | None = None,
                       return_priorities: bool = False):
        """Indices for injected uniforms ``u`` (float64 in [0, 1)): identical to the reference's
        ``get_prefix_sum_id((root * u).astype(dtype))`` (``sumtree_array.py:420-421``).  The root
        is read on the device, so nothing synchronises."""
        u = self._chk(u.reshape(-1), torch.float64, "u")
        m = u.numel()
        out = self._chk_out(out, m)
        prio = torch.empty(m, dtype=self.dtype, device=self.device) if return_priorities else None
        with torch.cuda.device(self.device):
            self._h.sample_uniform_device(u.data_ptr(), m, out.data_ptr(),
                                          prio.data_ptr() if prio is not None else 0, _stream())
        return (out, prio) if return_priorities else out

    @property
    def prob_dtype(self) -> torch.dtype:
        return _NP_TO_TORCH[np.dtype(_lib.PROB_DTYPES[self._h.code])]

    def sample_weights(self, u: torch.Tensor, out=None):
        """``(indices, priorities, probabilities)`` for injected uniforms: ``priorities = leaves[indices]``,
        ``probabilities = priorities / total`` exactly as NumPy's ``x[idx] / x.sum()`` computes it (one
        division in the tree dtype; float64 for integer trees) -- the inputs of the PER importance
        weights ``(N * P(i)) ** -beta`` (reference ``README.md:7``).  ``out``: optional tuple of three buffers."""
        u = self._chk(u.reshape(-1), torch.float64, "u")
        m = u.numel()
        o_idx, o_prio, o_prob = out if out is not None else (None, None, None)
        o_idx = self._chk_out(o_idx, m)
        o_prio = self._chk_out(o_prio, m, self.dtype, "priorities")
        o_prob = self._chk_out(o_prob, m, self.prob_dtype, "probabilities")
        with torch.cuda.device(self.device):
            self._h.sample_weights_device(u.data_ptr(), m, o_idx.data_ptr(), o_prio.data_ptr(), o_prob.data_ptr(), _stream())
        return o_idx, o_prio, o_prob

    def per_step_(self, idx: torch.Tensor, vals: torch.Tensor, u: torch.Tensor, out=None):
        """One fused PER step: apply the priority updates, then draw ``len(u)`` samples
        and return ``(indices, priorities, probabilities)`` -- one library call, nothing synchronises,
        capturable in a CUDA graph after ``reserve``."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        u = self._chk(u.reshape(-1), torch.float64, "u")
        m = u.numel()
        o_idx, o_prio, o_prob = out if out is not None else (None, None, None)
        o_idx = self._chk_out(o_idx, m)
        o_prio = self._chk_out(o_prio, m, self.dtype, "priorities")
        o_prob = self._chk_out(o_prob, m, self.prob_dtype, "probabilities")
        with torch.cuda.device(self.device):
            self._h.per_step_device(idx.data_ptr(), idx.numel(), vals.data_ptr(), vals.numel(), u.data_ptr(), m,
                                    o_idx.data_ptr(), o_prio.data_ptr(), o_prob.data_ptr(), _stream())
        return o_idx, o_prio, o_prob

    def sample(self, nsamples: int, generator: torch.Generator | None = None,
               return_priorities: bool = False):
        u = torch.rand(nsamples, dtype=torch.float64, device=self.device, generator=generator)
        return self.sample_uniform(u, return_priorities=return_priorities)

    def gather(self, idx: torch.Tensor) -> torch.Tensor:
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        out = torch.empty(idx.numel(), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._h.gather_leaves_device(idx.data_ptr(), idx.numel(), out.data_ptr(), _stream())
        return out

    # ---- sums -------------------------------------------------------------------------------
    def strided_sum(self, stride: int) -> torch.Tensor:
        stride = int(stride)
        nout = self.n // stride + (1 if self.n % stride else 0)
        out = torch.empty(nout, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._h.strided_sum_device(stride, out.data_ptr(), nout, _stream())
        return out

    def axis_fold(self, A: int, R: int, C: int) -> torch.Tensor:
        out_np = _lib.FOLD_DTYPES[self._h.code]
        out = torch.empty((A, C), dtype=_NP_TO_TORCH[np.dtype(out_np)], device=self.device)
        with torch.cuda.device(self.device):
            self._h.axis_fold_device(A, R, C, out.data_ptr(), _stream())
        return out

    def sum_over(self, index_from: int, index_to: int):
        torch.cuda.current_stream(self.device).synchronize()
        return self._h.sum_over(index_from, index_to)


class DeviceMinTree:
    """Device-resident MIN tree: the reference's ``MinTree`` / ``CircularMinTree``
    (``starr/experimental/slow.py:53-78, 91-98``) for PER buffers that need ``min(priorities)`` for the
    importance-weight normalisation.  Same heap layout as ``DeviceSumTree``; unset leaves are ``+inf``
    (integer dtypes: the largest value of the type)."""

    def __init__(self, n: int, dtype=torch.float32, device=None):
        if dtype not in _TORCH_TO_NP:
            raise TypeError("No matching signature found")
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceMinTree needs a CUDA device; starr_b200 has no CPU fallback")
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n, self.dtype = int(n), dtype
        assert self.n > 0, "size must be greater than 0"
        self._cursor = 0
        with torch.cuda.device(self.device):
            self.heap = torch.empty(2 * self.n, dtype=dtype, device=self.device)
            torch.cuda.current_stream().synchronize()
            self._h = _lib.TreeHandle(self.n, _TORCH_TO_NP[dtype], self.device.index,
                                      external_heap=self.heap.data_ptr(), kind="min")

    def __len__(self) -> int:
        return self.n

    @property
    def leaves(self) -> torch.Tensor:
        return self.heap[self.n:]

    @property
    def tree(self) -> torch.Tensor:
        return self.heap[:self.n]

    def min(self) -> torch.Tensor:
        """0-d device view of the root (``MinTree.min()``, slow.py:77-78)."""
        return self.heap[1]

    def _chk(self, t, dtype, name):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device):
            raise ValueError(f"{name} must be a CUDA tensor on {self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
        return t.contiguous()

    def update_(self, idx: torch.Tensor, vals: torch.Tensor) -> "DeviceMinTree":
        """``leaves[idx[j]] = vals[j % len(vals)]`` in batch order (the last write to a leaf wins), every
        ancestor recomputed as ``min(left, right)`` (slow.py:62-68).  Asynchronous; ``poll()`` reports
        an out-of-range index (the whole batch is then skipped)."""
        idx = self._chk(idx.reshape(-1), torch.int64, "idx")
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        with torch.cuda.device(self.device):
            self._h.min_update_device(idx.data_ptr(), idx.numel(), vals.data_ptr(), vals.numel(), _stream())
        return self

    def append_(self, vals: torch.Tensor) -> "DeviceMinTree":
        """``CircularMinTree.append`` for a batch (slow.py:96-98)."""
        vals = self._chk(vals.reshape(-1), self.dtype, "vals")
        k = vals.numel()
        if k:
            idx = (torch.arange(k, dtype=torch.int64, device=self.device) + self._cursor) % self.n
            self.update_(idx, vals)
            self._cursor = (self._cursor + k) % self.n
        return self

    def build_(self, leaves: torch.Tensor | None = None) -> "DeviceMinTree":
        if leaves is not None:
            if leaves.numel() != self.n:
                raise TypeError("array and tree must have the same size")
            self.leaves.copy_(leaves.to(device=self.device, dtype=self.dtype).reshape(-1))
        with torch.cuda.device(self.device):
            self._h.min_build(_stream())
        return self

    def poll(self) -> None:
        with torch.cuda.device(self.device):
            self._h.poll_status(_stream(), clear=True)
