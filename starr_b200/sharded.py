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
  plan    every rank computes, for EVERY element of the batch, the owning shard and the stable
          position ("slot") of the element among those of its shard, and compacts its own
          elements in batch order (three streaming kernels, st_shard.cu).  All ranks therefore
          know where each shard's results will sit in that shard's compact slab.
  update  every rank applies the updates it owns (batch order preserved) and writes their
          differences d_j, followed by its new shard root, into its slab; ONE all-gather of the
          slabs moves everything (each d_j crosses NVLink once, no zero-filled all-reduce); the
          d_j are put back into batch order and every rank folds them into the top tree (the
          reference adds EVERY update's d to the top nodes in order, so roots alone are not
          enough for floats).
  sample  every rank routes all queries through the top tree (same compares / subtractions as
          the first g levels of the reference descent), finishes the ones it owns locally with
          the left-over prefix, and the local leaf ids (int32) are all-gathered as slabs and
          expanded to global indices in batch order.

Peer exchange (``peer=True``, the default on CUDA when every rank can reach every other rank's memory):
the three all-gathers and the passes that put slabs back into batch order are replaced by direct stores
into an exchange buffer of EVERY rank (``csrc/st_peer.cu``, CUDA IPC over NVLink / NVSwitch):
  update  the leaf pass of the local update emits the d_j; one coalesced copy per peer (``st_peer_put``) stores the
          rank's slab behind the slabs of the lower ranks in every rank's buffer WHILE the local ancestors are
          still being updated (side stream); shard roots and the rank's 1/G of the top-fold records follow the
          same way; a release flag per rank + a spinning wait kernel replace the collective's implicit barrier.
  sample  routing on the top tree and the selection of the rank's queries are one kernel
          (``st_route_compact_device``: no plan, no host round trip); the local descent writes (batch position,
          local leaf) pairs, which go to every rank as one coalesced copy; a scatter puts them in batch order.
  (Storing every result straight at its batch position on every peer -- ``st_peer_scatter``,
  ``st_search_push_device`` -- was measured first: a 4- or 8-byte remote store costs a whole NVLink packet.)
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _hostlogic


class UpdatePlan:
    """Result of ``ShardedSumTree.plan_update``: who owns which element of the batch."""
    __slots__ = ("ni", "shard", "slot", "own_idx", "own_vals", "counts", "buf")

    def __init__(self, ni, shard, slot, own_idx, own_vals, counts, buf):
        self.ni, self.shard, self.slot, self.own_idx, self.own_vals = ni, shard, slot, own_idx, own_vals
        self.counts, self.buf = counts, buf


def _default_factory(n, dtype, device):
    from .device_tree import DeviceSumTree
    return DeviceSumTree(n, dtype=dtype, device=device)


class ShardedSumTree:
    def __init__(self, n_global: int, dtype=torch.float32, local_leaves: torch.Tensor | None = None,
                 group=None, device=None, tree_factory=None, peer: bool | None = None):
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
        self._marks = None     # diagnostics: list of (label, cuda event) while phase_times() runs
        self._bufs = [None, None, None]
        self._turn = 0
        # peer-memory exchange: created by the first batch (a collective: every rank sees the same sizes)
        if peer is None:     # STARR_B200_PEER=0 keeps the NCCL all-gather exchange (A/B comparisons)
            peer = self.device.type == "cuda" and tree_factory is None and os.environ.get("STARR_B200_PEER", "1") != "0"
        self._peer_wanted = bool(peer)
        self._peer = None
        self._peer_cap = (0, 0)
        self._ustep = self._sstep = 0
        self._side = None
        self._slices = None          # SlicesExchange (starr_b200/_slices.py), created by the first distributed batch
        if local_leaves is not None:
            self.build_(local_leaves)

    @staticmethod
    def _dist():
        return dist

    # ---- helpers --------------------------------------------------------------------------------
    def reserve(self, max_global_batch: int) -> None:
        """Pre-size every workspace a batch of this size needs (update pipeline, plan scratch, top-fold
        records), so that no stream-ordered call reallocates or synchronises later."""
        for tree in (self.local, self.top):
            h = getattr(tree, "handle", None)
            if h is not None and hasattr(h, "reserve_sharded"):
                h.reserve_sharded(max_global_batch, self.world)
            else:
                tree.reserve(max_global_batch)
        if self._peer_wanted:       # collective, like every call on a sharded tree: the exchange area for this batch size
            self._ensure_peer(mu=max_global_batch, ms=max_global_batch)

    def _mark(self, label: str) -> None:
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((label, ev))

    # ---- peer exchange area ---------------------------------------------------------------------
    CH_DELTAS, CH_RECS, CH_SAMPLE = 0, 1, 2

    def _peer_layout(self, mu: int, ms: int):
        """Byte offsets of the double-buffered regions for update batches of <= mu and sample batches of <= ms."""
        es = torch.empty(0, dtype=self.dtype).element_size()
        al = lambda x: (x + 255) // 256 * 256                                   # noqa: E731
        nodes = self.world - 1
        rb = self.top.top_fold_record_bytes() if hasattr(self.top, "top_fold_record_bytes") else 0
        per = ((mu + 1023) // 1024 + self.world - 1) // self.world
        # d: the shards' difference slabs back to back (exactly mu elements); pairs: one slab of (position, local leaf)
        # int32 pairs per rank, each able to hold the whole batch (a skewed tree may route every query to one shard)
        sizes = {"d": al(mu * es), "roots": al(self.world * es), "recs": al(per * self.world * nodes * rb),
                 "pairs": al(ms * 8) * self.world, "cnt": 256}
        off, lay = 0, {}
        for name, size in sizes.items():
            lay[name] = (off, off + size)                                        # parity 0 / parity 1
            off += 2 * size
        lay["bytes"] = off
        return lay

    def _ensure_peer(self, mu: int = 0, ms: int = 0) -> bool:
        """(Re)create the exchange area when a larger batch arrives.  Collective: batches are replicated, so every
        rank gets here with the same sizes at the same call."""
        if not self._peer_wanted:
            return False
        cu, cs = self._peer_cap
        if self._peer is not None and mu <= cu and ms <= cs:
            return True
        from . import _lib
        mu, ms = max(mu, cu, 1024), max(ms, cs, 1024)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)              # nobody is still storing into the old area
        self._drop_peer()
        lay = self._peer_layout(mu, ms)
        area = _lib.PeerArea(self.device.index, self.rank, self.world, lay["bytes"])
        addrs = [None] * self.world
        dist.all_gather_object(addrs, area.address(), group=self.group)
        area.connect(addrs)
        dist.barrier(group=self.group)              # every rank has opened every buffer before the first store
        self._peer, self._peer_cap, self._lay = area, (mu, ms), lay
        self._ustep = self._sstep = 0
        self._rank_pos = torch.tensor([self.rank], dtype=torch.int32, device=self.device)
        self._own_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._pairs = torch.empty(2 * ms, dtype=torch.int32, device=self.device)
        self._d_all = torch.empty(mu, dtype=self.dtype, device=self.device)
        self._pair_slab = (ms * 8 + 255) // 256 * 256
        self._side = self._side or torch.cuda.Stream(device=self.device)
        return True

    def _drop_peer(self) -> None:
        old, self._peer = self._peer, None
        if old is not None:
            old.disconnect()
            dist.barrier(group=self.group)          # an exported buffer is freed only after every importer closed it
            old.close()

    def close(self) -> None:
        """Release the peer exchange area (collective; every rank calls it).  Without it the area is freed when the
        object is collected, which is only safe once every rank has finished with the tree."""
        if self._slices is not None:
            self._slices.close()
            self._slices = None
        if self._peer is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._drop_peer()

    def _region(self, name: str, parity: int, dtype, count: int) -> torch.Tensor:
        """Zero-copy tensor over a region of the LOCAL exchange buffer."""
        off = self._lay[name][parity]
        nbytes = count * torch.empty(0, dtype=dtype).element_size()
        iface = {"shape": (max(nbytes, 1),), "typestr": "|u1", "version": 3, "data": (self._peer.payload_ptr + off, False)}
        holder = type("_Region", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.device)[:nbytes].view(dtype)

    def phase_times(self, next_batch, reps: int = 10, slices: bool = False):
        """Diagnostics (tools/sharded_breakdown.py): median device time in microseconds of every
        phase of ``update_`` and ``sample_uniform`` (``slices``: of ``update_slices_`` and ``sample_slices``);
        ``next_batch()`` returns ``(idx, vals, u)``."""
        import statistics
        acc: dict[str, list[float]] = {}
        for _ in range(reps):
            idx, vals, u = next_batch()
            dist.barrier(group=self.group)
            torch.cuda.synchronize()
            self._marks = []
            self._mark("start")
            if slices:
                self.update_slices_(idx, vals)
                self._mark("update: done")
                self.sample_slices(u)
            else:
                self.update_(idx, vals)
                self._mark("update: done")
                self.sample_uniform(u)
            torch.cuda.synchronize()
            marks, self._marks = self._marks, None
            for (_, a), (label, b) in zip(marks, marks[1:]):
                acc.setdefault(label, []).append(a.elapsed_time(b) * 1e3)
        return [(k, statistics.median(v)) for k, v in acc.items()]

    def poll(self) -> None:
        """Raise the first deferred validation error; both trees' status words are cleared."""
        first = None
        for tree in (self.local, self.top):
            try:
                tree.poll()
            except (ValueError, IndexError) as e:
                first = first or e
        if first is not None:
            raise first
        if self._peer is not None:
            self._peer.status(torch.cuda.current_stream().cuda_stream)

    def _gather_roots(self) -> None:  # build only; updates carry the roots in the delta slabs
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
    @staticmethod
    def _slab_capacity(counts) -> int:
        """Common slab length of one exchange: the largest per-shard count (every rank computes
        the same counts from the replicated batch), rounded up so that slabs stay 256-byte aligned."""
        return max(64, (max(counts) + 63) // 64 * 64)

    def _plan_buffers(self, which: int, m: int):
        """Persistent plan buffers (grow-only).  Sets 0 and 1 alternate between update plans, so
        that the plan of batch k+1 can be computed (on a side stream) while batch k is still
        being applied from the other set; set 2 belongs to ``sample_uniform``."""
        buf = self._bufs[which]
        if buf is None or buf["cap"] < m:
            if buf is not None and self.device.type == "cuda":
                torch.cuda.synchronize(self.device)        # nobody may still be reading the old set
            dev = self.device
            buf = {"cap": m, "free": None,
                   "out": (torch.empty(m, dtype=torch.uint8, device=dev), torch.empty(m, dtype=torch.int32, device=dev),
                           torch.empty(m, dtype=torch.int64, device=dev), torch.empty(m, dtype=self.dtype, device=dev)),
                   "pos": torch.empty(m, dtype=torch.int32, device=dev) if self._peer_wanted else None,
                   "d_own": torch.empty(m, dtype=self.dtype, device=dev) if self._peer_wanted else None}
            self._bufs[which] = buf
        return buf

    def plan_update(self, idx: torch.Tensor, vals: torch.Tensor, stream=None) -> "UpdatePlan":
        """Owner, slot and compaction of one update batch (it reads nothing of the tree), ahead of
        ``update_(idx, vals, plan=...)``.  With ``stream`` (a side stream) the plan runs there and
        overlaps with whatever the current stream is still doing; the caller then guarantees that
        ``idx`` and ``vals`` are complete on the device (not produced by work still pending on the
        current stream).  Raises ``IndexError`` / ``ValueError`` for a bad batch -- on every rank
        alike, before anything is modified -- and ``ZeroDivisionError`` for empty ``vals``."""
        idx = idx.reshape(-1)
        vals = vals.reshape(-1)
        ni, nv = idx.numel(), vals.numel()
        if ni and nv == 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        if ni == 0:
            if nv and bool((vals < 0).any()):              # the reference scans vals even then (pyx:37-39)
                raise ValueError("vals must be non-negative")
            return UpdatePlan(0, None, None, None, None, None, None)
        which = self._turn
        self._turn ^= 1
        buf = self._plan_buffers(which, ni)
        on_cuda = self.device.type == "cuda"
        extra = {"own_pos": buf["pos"]} if buf["pos"] is not None else {}
        if on_cuda and stream is not None:
            if buf["free"] is not None:
                stream.wait_event(buf["free"])             # the update that read this set has finished
            with torch.cuda.stream(stream):
                res = self.local.shard_plan(idx, self.shift, self.world, self.rank, vals, True, out=buf["out"],
                                            result_slot=which, **extra)
        else:
            res = self.local.shard_plan(idx, self.shift, self.world, self.rank, vals, True,
                                        **({"out": buf["out"], "result_slot": which, **extra} if on_cuda else {}))
        return UpdatePlan(ni, *res, buf)

    def update_(self, idx: torch.Tensor | None = None, vals: torch.Tensor | None = None, check: bool = False,
                plan: "UpdatePlan | None" = None) -> "ShardedSumTree":
        """Replicated batch: ``idx`` are GLOBAL leaf ids, ``vals`` cycles over the global batch
        position (reference pyx:44).  A bad batch raises at once on every rank (the plan validates
        the whole batch and synchronises anyway); ``check`` is kept for compatibility."""
        if plan is None:
            plan = self.plan_update(idx, vals)
        if plan.ni == 0:
            return self
        shard, slot, own_idx, own_vals, counts = plan.shard, plan.slot, plan.own_idx, plan.own_vals, plan.counts
        self._mark("update: plan (owner, slot, compaction)")
        if plan.buf is not None and plan.buf.get("pos") is not None and self._ensure_peer(mu=plan.ni):
            return self._update_peer(plan)
        dev = shard.device
        c, cap = counts[self.rank], self._slab_capacity(counts)
        # slab of this rank: its differences d_j in batch order, then (entry `cap`) its new shard root;
        # the slab length stays a multiple of 64 elements so that every rank's part of the all-gather
        # is 16-byte aligned (an odd length makes NCCL fall back to its slow unaligned protocol)
        mine = torch.empty(cap + 64, dtype=self.dtype, device=dev)
        if c:
            self.local.update_with_deltas_(own_idx[:c], own_vals[:c], out=mine[:c])
        mine[cap:cap + 1].copy_(self.local.heap[1:2])
        self._mark("update: local update")
        slabs = torch.empty((self.world, cap + 64), dtype=self.dtype, device=dev)
        dist.all_gather_into_tensor(slabs.view(-1), mine, group=self.group)
        self._mark("update: exchange deltas + roots")
        d_all = self.local.shard_expand_values(shard, slot, slabs)      # batch order again
        self._mark("update: deltas to batch order")
        self._fold_top(shard, d_all)                                    # ordered fold, all ranks alike
        self.top.leaves.copy_(slabs[:, cap])
        if plan.buf is not None and dev.type == "cuda":
            ev = torch.cuda.Event()
            ev.record()
            plan.buf["free"] = ev
        self._mark("update: top fold")
        return self

    def _update_peer(self, plan: "UpdatePlan") -> "ShardedSumTree":
        """``update_`` over the peer exchange area (module docstring): no collective, no slab, no expansion."""
        peer, ni, buf = self._peer, plan.ni, plan.buf
        c = plan.counts[self.rank]
        self._ustep += 1
        step, par = self._ustep, self._ustep & 1
        es = self.local.heap.element_size()
        main = torch.cuda.current_stream()
        d_own = buf["d_own"]
        starts = [0] * self.world                  # slab of shard q inside every rank's "d" region (all ranks know all counts)
        for q in range(1, self.world):
            starts[q] = starts[q - 1] + plan.counts[q - 1]
        if c:
            # leaf pass first: the differences exist, and travel (side stream, one coalesced copy per peer) while the
            # ancestors are updated
            self.local.update_begin_(plan.own_idx[:c], plan.own_vals[:c], out=d_own)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                peer.put(d_own.data_ptr(), c, es, self._lay["d"][par] + starts[self.rank] * es, stream=self._side.cuda_stream)
            self.local.update_finish_()
            main.wait_stream(self._side)
        self._mark("update: local update (+ deltas to every rank)")
        peer.scatter(self.local.heap[1:2].data_ptr(), self._rank_pos.data_ptr(), 1, es, self._lay["roots"][par], self.world,
                     stream=main.cuda_stream)
        peer.signal(self.CH_DELTAS, step, main.cuda_stream)
        peer.wait(self.CH_DELTAS, step, main.cuda_stream)
        self._mark("update: wait for all ranks' deltas + roots")
        d_all = self.local.shard_expand_cat(plan.shard, plan.slot, self._region("d", par, self.dtype, ni), starts,
                                            out=self._d_all)
        self._mark("update: deltas to batch order")
        self._fold_top(plan.shard, d_all, peer_step=(step, par))
        self.top.leaves.copy_(self._region("roots", par, self.dtype, self.world))
        ev = torch.cuda.Event()
        ev.record()
        buf["free"] = ev
        self._mark("update: top fold")
        return self

    def _fold_top(self, shard: torch.Tensor, d_all: torch.Tensor, peer_step=None) -> None:
        """Ordered fold of ALL differences into the replicated top tree.  With the CUDA tree the
        chunk summaries (the parallel part) are split over the ranks and exchanged with one small
        all-gather; the stitch is replicated.  Results are identical on every rank."""
        top = self.top
        ni = shard.numel()
        if not hasattr(top, "top_fold_maps_") or ni <= 4096 or not self.dtype.is_floating_point:
            top.apply_deltas_(shard.to(torch.int64), d_all)
            return
        nodes = self.world - 1
        rb = top.top_fold_record_bytes()
        nchunks = (ni + 1023) // 1024
        per = (nchunks + self.world - 1) // self.world
        lo = self.rank * per
        if peer_step is not None:
            # this rank's 1/G of the chunk summaries goes straight into every rank's record array
            step, par = peer_step
            recs = self._region("recs", par, torch.uint8, per * self.world * nodes * rb)
            top.top_fold_maps_(shard, d_all, lo, lo + per, recs)
            self._mark("update: top fold summaries")
            s = torch.cuda.current_stream().cuda_stream
            span = per * nodes * rb
            self._peer.broadcast(self._lay["recs"][par] + lo * nodes * rb, span, s)
            self._peer.signal(self.CH_RECS, step, s)
            self._peer.wait(self.CH_RECS, step, s)
            self._mark("update: top fold records to every rank")
            top.top_fold_apply_(shard, d_all, recs)
            return
        recs = torch.empty(per * self.world * nodes * rb, dtype=torch.uint8, device=shard.device)
        top.top_fold_maps_(shard, d_all, lo, lo + per, recs)
        self._mark("update: top fold summaries")
        slab = recs[lo * nodes * rb:(lo + per) * nodes * rb].clone()
        dist.all_gather_into_tensor(recs, slab, group=self.group)
        self._mark("update: top fold all-gather")
        top.top_fold_apply_(shard, d_all, recs)

    # ---- sample ---------------------------------------------------------------------------------
    def sample_uniform(self, u: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Global leaf indices for replicated uniforms ``u`` (float64), identical to the
        reference's ``get_prefix_sum_id((root * u).astype(dtype))`` on the whole tree."""
        u = u.reshape(-1)
        m = u.numel()
        if out is None:
            out = torch.empty(m, dtype=torch.int64, device=u.device)
        if m == 0:
            return out
        if self._ensure_peer(ms=m):
            return self._sample_peer(u, out)
        owner, resid = self.top.route_uniform(u)
        self._mark("sample: route")
        if hasattr(self.local, "search_counted"):
            # CUDA tree: enqueue the plan, then the local search with the count read on the device, and
            # only then wait for the plan's counts -- the host round trip hides behind the search
            bufs = self._plan_buffers(2, m)["out"]
            shard, slot, _, own_resid, _ = self.local.shard_plan(
                owner, 0, self.world, self.rank, resid, False, want_ids=False, out=bufs, result_slot=2, wait=False)
            found = bufs[2]                                              # int64[m], unused by this plan
            self.local.search_counted(own_resid, 2, self.rank, found)
            counts = self.local.shard_plan_wait(2, self.world)
            self._mark("sample: plan + local search")
            c, cap = counts[self.rank], self._slab_capacity(counts)
            mine = torch.empty(cap, dtype=torch.int32, device=u.device)
            if c:
                mine[:c].copy_(found[:c])                                # local leaf ids fit 31 bits
        else:
            shard, slot, _, own_resid, counts = self.local.shard_plan(
                owner, 0, self.world, self.rank, resid, False, want_ids=False)
            c, cap = counts[self.rank], self._slab_capacity(counts)
            mine = torch.empty(cap, dtype=torch.int32, device=u.device)
            if c:
                mine[:c].copy_(self.local.search(own_resid[:c]))
        self._mark("sample: local leaf ids")
        slabs = torch.empty((self.world, cap), dtype=torch.int32, device=u.device)
        dist.all_gather_into_tensor(slabs.view(-1), mine, group=self.group)
        self._mark("sample: exchange indices")
        self.local.shard_expand_indices(shard, slot, slabs, self.n_local, out=out)
        self._mark("sample: indices to batch order")
        return out

    def _sample_peer(self, u: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``sample_uniform`` over the peer exchange area: route + select in one kernel, the local descent stores
        global indices at their batch positions on every rank; no plan, no host round trip, no collective."""
        m = u.numel()
        self._sstep += 1
        step, par = self._sstep, self._sstep & 1
        bufs = self._plan_buffers(2, m)
        own_resid, own_pos = bufs["out"][3], bufs["pos"]
        self.top.route_compact(u, self.rank, own_resid, own_pos, self._own_count)
        self._mark("sample: route + select")
        self.local.search_pairs(own_resid, self._own_count, own_pos, self._pairs)
        self._mark("sample: local search")
        s = torch.cuda.current_stream().cuda_stream
        self._peer.put(self._pairs.data_ptr(), m, 8, self._lay["pairs"][par] + self.rank * self._pair_slab,
                       count_dev_ptr=self._own_count.data_ptr(), count_offset=self._lay["cnt"][par] + 4 * self.rank, stream=s)
        self._peer.signal(self.CH_SAMPLE, step, s)
        self._peer.wait(self.CH_SAMPLE, step, s)
        self._mark("sample: pairs to every rank + wait")
        self._peer.scatter_pairs(self._lay["pairs"][par], self._pair_slab, self._lay["cnt"][par], self.n_local, out.data_ptr(), m, s)
        self._mark("sample: indices to batch order")
        return out

    # ---- slices layout: the batch is distributed over the ranks (starr_b200/_slices.py) ---------------------------
    def _slices_exchange(self):
        if not self._peer_wanted:
            return None
        if self._slices is None:
            from ._slices import SlicesExchange
            self._slices = SlicesExchange(self)
        return self._slices

    def reserve_slices(self, max_slice: int, expected_local: int | None = None) -> None:
        """Pre-size the exchange area for slices of up to ``max_slice`` elements per rank and the local update
        workspace for ``expected_local`` owned elements per batch (default: twice the slice).  Collective."""
        # update workspace + the partition's scratch area (local tree: update plans, top tree: sample plans), so that no
        # stream-ordered call reallocates (= waits for the whole device) later
        for tree, batch in ((self.local, int(expected_local or 2 * max_slice)), (self.top, int(max_slice))):
            h = getattr(tree, "handle", None)
            if h is not None and hasattr(h, "reserve_sharded"):
                h.reserve_sharded(max(batch, int(max_slice)), self.world)
            else:
                tree.reserve(batch)
        sx = self._slices_exchange()
        if sx is not None:
            sx.ensure(int(max_slice))

    def plan_update_slices(self, idx: torch.Tensor, vals: torch.Tensor, stream=None):
        """First half of ``update_slices_``: owner / slot of every element of this rank's slice and the exchange of
        the slices (ids and values travel to their owners).  Reads nothing of the tree, so with ``stream`` (a side
        stream; ``idx`` / ``vals`` must be complete on the device) it overlaps the batch that is being applied."""
        sx = self._slices_exchange()
        if sx is None:
            return ("gathered", idx, vals)
        return sx.plan_update(idx, vals, stream)

    def update_slices_(self, idx: torch.Tensor | None = None, vals: torch.Tensor | None = None, plan=None) -> "ShardedSumTree":
        """Distributed batch: rank r supplies elements ``[r m, (r+1) m)`` of the global batch -- ``idx`` are GLOBAL
        leaf ids, ``vals`` one value per index, the same ``m`` on every rank.  Bit-identical to ONE reference tree
        applying the concatenated batch in rank order (pyx:41-49).  A bad slice on any rank raises on every rank,
        before anything is modified.  Without peer memory the slices are all-gathered and applied as a replicated batch."""
        if plan is None:
            plan = self.plan_update_slices(idx, vals)
        if isinstance(plan, tuple):
            _, idx, vals = plan
            idx, vals = idx.reshape(-1).contiguous(), vals.reshape(-1).contiguous()
            if vals.numel() != idx.numel():
                raise TypeError("slices layout: vals must have one entry per index of the slice")
            all_idx = torch.empty(self.world * idx.numel(), dtype=idx.dtype, device=idx.device)
            all_vals = torch.empty(self.world * vals.numel(), dtype=vals.dtype, device=vals.device)
            dist.all_gather_into_tensor(all_idx, idx, group=self.group)
            dist.all_gather_into_tensor(all_vals, vals, group=self.group)
            return self.update_(all_idx, all_vals)
        self._slices.update(plan)
        return self

    def sample_slices(self, u: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Global leaf indices for THIS rank's uniforms (float64; the same count on every rank): what the reference's
        ``get_prefix_sum_id((root * u).astype(dtype))`` returns for them on the whole tree."""
        u = u.reshape(-1)
        m = u.numel()
        if out is None:
            out = torch.empty(m, dtype=torch.int64, device=u.device)
        sx = self._slices_exchange()
        if sx is None:
            all_u = torch.empty(self.world * m, dtype=u.dtype, device=u.device)
            dist.all_gather_into_tensor(all_u, u.contiguous(), group=self.group)
            out.copy_(self.sample_uniform(all_u)[self.rank * m:(self.rank + 1) * m])
            return out
        if u.dtype != torch.float64 or not u.is_contiguous() or out.dtype != torch.int64 or not out.is_contiguous() or out.numel() < m:
            raise TypeError("u must be a contiguous float64 tensor and out a contiguous int64 tensor of the same size")
        return self._slices.sample(u, out)

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
