"""Slices layout of ``ShardedSumTree``: the batch is DISTRIBUTED over the ranks.

Rank r supplies elements ``[r m, (r+1) m)`` of the global batch (the way data-parallel learners produce it:
every rank updates the priorities of, and samples, its own 1/G of the batch), so no rank ever touches
O(M_global) data -- the replicated layout (``ShardedSumTree.update_`` / ``sample_uniform``) makes every
rank plan, expand and fold the whole batch.  The result is the same bit for bit: ONE reference tree
(``starr/src/_sumtree_array.pyx:22-122``) that applies the concatenated batch in rank-major order.

Transport: the peer-memory exchange area (``csrc/st_peer.cu``, CUDA IPC over NVLink / NVSwitch).
Per update (all stream-ordered, one host read of the count matrix):

  plan     owner + stable slot of every slice element, slice partitioned by owner (``st_shard_partition_device``);
           the row of per-owner counts goes to every rank (``st_slices_rows``), the partitioned ids / values go to
           their owners (``st_slices_send``), each owner receiving its elements in global batch order.  Reads nothing
           of the tree: ``plan_update`` may run on a side stream while the previous batch is still being applied.
  update   leaf pass of the local update (``update_begin_``) -> the differences d_j travel back to the rank that
           supplied the element (``st_slices_return``) while the local ancestors are updated (``update_finish_``);
           every rank summarises ITS slice for the top tree (``st_top_fold_maps_slice_device``), the records (a few
           hundred KiB) go to everybody, and the stitch reads other ranks' slices -- through peer memory -- only for
           chunks that must be folded one add at a time.  Shard roots travel as before.
  sample   route on the top tree, partition the left-over prefixes by owner, send, local descent
           (``st_search_leaf32_device``, count read on the device), 32-bit leaves back, global indices in slice order.
           No host round trip at all.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib

_FLAG_NEG, _FLAG_INDEX = 1, 4      # STF_NEG_VALS / STF_INDEX (csrc/st_internal.h)


class SlicesPlan:
    __slots__ = ("step", "par", "m", "event")

    def __init__(self, step, par, m, event):
        self.step, self.par, self.m, self.event = step, par, m, event


class SlicesExchange:
    CH_UROW, CH_UDATA, CH_DBACK, CH_RECS, CH_ROOTS, CH_SROW, CH_SDATA, CH_SBACK = range(8)

    def __init__(self, tree):
        # weak: the tree owns this object; a reference cycle would leave the trees' device memory to the cyclic collector,
        # whose cudaFree (a device-wide wait) could then land in the middle of somebody's step
        self.t = weakref.proxy(tree)
        self.area = None
        self.cap = 0
        self._pstep = self._ustep = self._sstep = 0
        self._free = [None, None]           # event: the update that used this parity's plan buffers has finished
        self._side = None
        # STARR_B200_SLICES_SERIAL=1: the top fold runs on the caller's stream in front of the local ancestors instead of
        # beside them (A/B timing; and the thread-rank test at 8 ranks, where 8 x 3 streams with spinning wait kernels
        # would share the device's hardware queues)
        self.serial = os.environ.get("STARR_B200_SLICES_SERIAL", "0") == "1"
        # STARR_B200_SLICES_SAMPLE=matrix: the first version of sample (stable partition + count matrix, three flag round
        # trips) instead of the direct push (two); A/B timing only
        self.sample_matrix = os.environ.get("STARR_B200_SLICES_SAMPLE", "push") == "matrix"

    # ---- exchange area ------------------------------------------------------------------------------------------
    def _layout(self, cap: int) -> dict:
        t = self.t
        w, es = t.world, torch.empty(0, dtype=t.dtype).element_size()
        al = lambda x: (max(x, 1) + 255) // 256 * 256                            # noqa: E731
        span = (cap + 1023) // 1024 * 1024
        rb = t.top.top_fold_record_bytes() if t.dtype.is_floating_point else 0
        sizes = {
            "mat_u": al(w * (w + 1) * 4), "mat_s": al(w * (w + 1) * 4),
            "recv_idx": al(w * cap * 8), "recv_val": al(w * cap * es),           # worst case: every element to one owner
            "dback": al(cap * es), "dslice": al(span * es), "shard": al(span),
            "roots": al(w * es), "recs": al(w * (span // 1024) * (w - 1) * rb),
            # first version of sample (count matrix; A/B timing only)
            "recv_q": al(w * cap * es) if self.sample_matrix else 256, "back_q": al(cap * 4) if self.sample_matrix else 256,
            # sample without a count matrix: one receive region per source rank, results come back as pairs
            "rq_resid": al(w * cap * es), "rq_pos": al(w * cap * 4), "rq_cnt": 256, "sent_cnt": 256, "back": al(w * cap * 8),
        }
        off, lay = 0, {}
        for name, size in sizes.items():
            lay[name] = (off, off + size)                                         # parity 0 / parity 1
            off += 2 * size
        lay["bytes"] = off
        return lay

    def ensure(self, cap: int) -> None:
        """(Re)create the exchange area for slices of up to ``cap`` elements.  Collective."""
        if self.area is not None and cap <= self.cap:
            return
        t, dist = self.t, self.t._dist()
        cap = max(cap, self.cap, 1024)
        torch.cuda.synchronize(t.device)
        dist.barrier(group=t.group)
        self.close(collective=False)
        lay = self._layout(cap)
        area = _lib.PeerArea(t.device.index, t.rank, t.world, lay["bytes"])
        addrs = [None] * t.world
        dist.all_gather_object(addrs, area.address(), group=t.group)
        area.connect(addrs)
        for par in (0, 1):
            area.slices_table(par, lay["dslice"][par], lay["shard"][par])
        dist.barrier(group=t.group)
        self.area, self.cap, self.lay = area, cap, lay
        self._pstep = self._ustep = self._sstep = 0
        self._free = [None, None]
        dev, w = t.device, t.world
        mk = lambda n, dt: torch.empty(n, dtype=dt, device=dev)                    # noqa: E731
        self.shard = [mk(cap, torch.uint8) for _ in range(2)]
        self.slot = [mk(cap, torch.int32) for _ in range(2)]
        self.pidx = [mk(cap, torch.int64) for _ in range(2)]
        self.pval = [mk(cap, t.dtype) for _ in range(2)]
        self.row = [torch.zeros(w + 1, dtype=torch.int32, device=dev) for _ in range(2)]
        self.d_own = mk(w * cap, t.dtype)
        self.qshard, self.qslot, self.presid = mk(cap, torch.uint8), mk(cap, torch.int32), mk(cap, t.dtype)
        self.qowner, self.qresid = mk(cap, torch.int64), mk(cap, t.dtype)
        self.qrow = torch.zeros(w + 1, dtype=torch.int32, device=dev)
        self.qtotal = torch.zeros(1, dtype=torch.int32, device=dev)
        self.qcnt = torch.zeros(w, dtype=torch.int32, device=dev)
        self.found = mk(w * cap if self.sample_matrix else 1, torch.int32)
        self.rank_pos = torch.tensor([t.rank], dtype=torch.int32, device=dev)
        self._side = self._side or torch.cuda.Stream(device=dev)

    def close(self, collective: bool = True) -> None:
        old, self.area = self.area, None
        if old is None:
            return
        t = self.t
        if collective:
            torch.cuda.synchronize(t.device)
            t._dist().barrier(group=t.group)
        old.disconnect()
        t._dist().barrier(group=t.group)        # an exported buffer is freed only after every importer closed it
        old.close()

    def _ptr(self, name: str, par: int) -> int:
        return self.area.payload_ptr + self.lay[name][par]

    def _view(self, name: str, par: int, dtype, count: int) -> torch.Tensor:
        nbytes = count * torch.empty(0, dtype=dtype).element_size()
        iface = {"shape": (max(nbytes, 1),), "typestr": "|u1", "version": 3, "data": (self._ptr(name, par), False)}
        holder = type("_Region", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.t.device)[:nbytes].view(dtype)

    # ---- update -------------------------------------------------------------------------------------------------
    def plan_update(self, idx: torch.Tensor, vals: torch.Tensor, stream=None) -> SlicesPlan:
        t = self.t
        idx, vals = idx.reshape(-1), vals.reshape(-1)
        m = idx.numel()
        if vals.numel() != m:
            raise TypeError("slices layout: vals must have one entry per index of the slice")
        if idx.dtype != torch.int64 or vals.dtype != t.dtype or not idx.is_contiguous() or not vals.is_contiguous():
            raise TypeError("idx must be a contiguous int64 tensor and vals a contiguous tensor of the tree dtype")
        self.ensure(m)
        self._pstep += 1
        p, par = self._pstep, self._pstep & 1
        area, lay = self.area, self.lay
        es = vals.element_size()
        run_on = stream if stream is not None else torch.cuda.current_stream()
        if self._free[par] is not None:
            run_on.wait_event(self._free[par])          # the update that read this parity's buffers has finished
        with torch.cuda.stream(run_on):
            s = run_on.cuda_stream
            if m:
                t.local.handle.shard_partition_device(idx.data_ptr(), m, t.shift, t.world, vals.data_ptr(), True,
                                                      self.shard[par].data_ptr(), self.slot[par].data_ptr(),
                                                      self.pidx[par].data_ptr(), self.pval[par].data_ptr(),
                                                      self.row[par].data_ptr(), s)
            else:
                self.row[par].zero_()
            area.slices_rows(self.row[par].data_ptr(), lay["mat_u"][par], self.CH_UROW, p, s)
            area.wait(self.CH_UROW, p, s)
            area.slices_matrix_fetch(lay["mat_u"][par], s)
            area.slices_send(lay["mat_u"][par], self.pidx[par].data_ptr(), 8, lay["recv_idx"][par], m,
                             src1=self.pval[par].data_ptr(), esize1=es, dst1=lay["recv_val"][par], stream=s)
            area.signal(self.CH_UDATA, p, s)
            ev = torch.cuda.Event()
            ev.record(run_on)
        return SlicesPlan(p, par, m, ev)

    def update(self, plan: SlicesPlan) -> None:
        t, area, lay = self.t, self.area, self.lay
        if plan.step != self._ustep + 1:
            raise RuntimeError("slices plans must be applied in the order they were made")
        self._ustep = plan.step
        p, par, m, w = plan.step, plan.par, plan.m, t.world
        main = torch.cuda.current_stream()
        main.wait_event(plan.event)
        mat = area.slices_matrix_wait()                       # the one host round trip: the local update needs its count
        flags = 0
        for r in range(w):
            flags |= mat[r][w]
        if flags:                                             # every rank sees the same flags: all raise, nothing is modified
            area.wait(self.CH_UDATA, p, main.cuda_stream)     # the step is consumed on every rank alike
            self._mark_free(par, main)
            if flags & _FLAG_INDEX:
                raise IndexError("Out of bounds on buffer access (axis 0)")
            raise ValueError("vals must be non-negative")
        if any(sum(mat[r][:w]) != m for r in range(w)):       # the chunk numbering of the top fold assumes equal slices
            area.wait(self.CH_UDATA, p, main.cuda_stream)
            self._mark_free(par, main)
            raise ValueError("slices layout: every rank must supply the same number of elements per batch")
        c = sum(mat[r][t.rank] for r in range(w))
        es = t.local.heap.element_size()
        s = main.cuda_stream
        area.wait(self.CH_UDATA, p, s)
        t._mark("update: plan + exchange of the slices")
        if c:
            t.local.update_begin_(self._view("recv_idx", par, torch.int64, c), self._view("recv_val", par, t.dtype, c),
                                  out=self.d_own)
        t._mark("update: leaf pass")
        is_float = t.dtype.is_floating_point
        side = main if self.serial else self._side
        if is_float:
            span = max(1024, (m + 1023) // 1024 * 1024)
            per, nodes, rb = span // 1024, w - 1, t.top.top_fold_record_bytes()
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                ss = side.cuda_stream
                area.slices_return(lay["mat_u"][par], self.d_own.data_ptr(), es, lay["dback"][par], c, ss)
                area.signal(self.CH_DBACK, p, ss)
                area.wait(self.CH_DBACK, p, ss)
                area.slices_unpermute(lay["mat_u"][par], self.shard[par].data_ptr(), self.slot[par].data_ptr(), m,
                                      lay["dback"][par], es, self._ptr("dslice", par), shard_copy_offset=lay["shard"][par],
                                      stream=ss)
                recs = self._ptr("recs", par)
                t.top.handle.top_fold_maps_slice_device(self._ptr("shard", par), self._ptr("dslice", par), m, t.rank * per, per,
                                                        recs, ss)
                area.broadcast(lay["recs"][par] + t.rank * per * nodes * rb, per * nodes * rb, ss)
                area.signal(self.CH_RECS, p, ss)
                area.wait(self.CH_RECS, p, ss)
                t.top.handle.top_fold_apply_slices_device(area, par, span, m, recs, ss)
        if c:
            t.local.update_finish_()
        t._mark("update: local ancestors (top fold beside them)")
        area.scatter(t.local.heap[1:2].data_ptr(), self.rank_pos.data_ptr(), 1, es, lay["roots"][par], w, stream=s)
        area.signal(self.CH_ROOTS, p, s)
        area.wait(self.CH_ROOTS, p, s)
        t.top.leaves.copy_(self._view("roots", par, t.dtype, w))
        if is_float:
            if side is not main:
                main.wait_stream(side)
        else:
            t.top.build_()          # integer adds are modular: the pairwise build IS the reference's result
        self._mark_free(par, main)
        t._mark("update: roots + join")

    def _mark_free(self, par: int, stream) -> None:
        ev = torch.cuda.Event()
        ev.record(stream)
        self._free[par] = ev

    # ---- sample -------------------------------------------------------------------------------------------------
    def sample(self, u: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if self.sample_matrix:
            return self._sample_matrix(u, out)
        t = self.t
        m = u.numel()
        self.ensure(m)
        area, lay, cap = self.area, self.lay, self.cap
        self._sstep += 1
        p, par = self._sstep, self._sstep & 1
        s = torch.cuda.current_stream().cuda_stream
        # every prefix goes straight into its owner's receive region of this source, the counts follow with the flag
        t.top.handle.route_push_device(u.data_ptr(), m, area, lay["rq_resid"][par], lay["rq_pos"][par], cap, self.qcnt.data_ptr(), s)
        area.slices_counts(self.qcnt.data_ptr(), lay["rq_cnt"][par], lay["sent_cnt"][par], self.CH_SDATA, p, s)
        t._mark("sample: route + prefixes to their owners")
        area.wait(self.CH_SDATA, p, s)
        t._mark("sample: wait for every rank's prefixes")
        # local descent over the segments that arrived; (slice position, leaf) pairs go straight back to their sources
        t.local.handle.search_segments_push_device(area, lay["rq_resid"][par], lay["rq_pos"][par], lay["rq_cnt"][par], cap,
                                                   lay["back"][par], s)
        area.signal(self.CH_SBACK, p, s)
        t._mark("sample: local search (+ results to the querying ranks)")
        area.wait(self.CH_SBACK, p, s)
        t._mark("sample: wait for every rank's results")
        if m:
            area.scatter_pairs(lay["back"][par], cap * 8, lay["sent_cnt"][par], t.n_local, out.data_ptr(), m, s)
        t._mark("sample: indices to slice order")
        return out

    def _sample_matrix(self, u: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        t = self.t
        m = u.numel()
        self.ensure(m)
        area, lay = self.area, self.lay
        self._sstep += 1
        p, par, w = self._sstep, self._sstep & 1, t.world
        s = torch.cuda.current_stream().cuda_stream
        es = t.local.heap.element_size()
        if m:
            t.top.handle.route_uniform_device(u.data_ptr(), m, self.qowner.data_ptr(), self.qresid.data_ptr(), s)
            t.top.handle.shard_partition_device(self.qowner.data_ptr(), m, 0, w, self.qresid.data_ptr(), False,
                                                self.qshard.data_ptr(), self.qslot.data_ptr(), 0, self.presid.data_ptr(),
                                                self.qrow.data_ptr(), s)
        else:
            self.qrow.zero_()
        t._mark("sample: route + partition")
        area.slices_rows(self.qrow.data_ptr(), lay["mat_s"][par], self.CH_SROW, p, s)
        area.wait(self.CH_SROW, p, s)
        area.slices_send(lay["mat_s"][par], self.presid.data_ptr(), es, lay["recv_q"][par], m,
                         total_out_ptr=self.qtotal.data_ptr(), stream=s)
        area.signal(self.CH_SDATA, p, s)
        area.wait(self.CH_SDATA, p, s)
        t._mark("sample: prefixes to their owners")
        bound = w * m                      # the host does not know how many queries arrived: the device-side count trims
        if bound:
            t.local.handle.search_leaf32_device(self._ptr("recv_q", par), bound, self.qtotal.data_ptr(), self.found.data_ptr(), s)
        t._mark("sample: local search")
        area.slices_return(lay["mat_s"][par], self.found.data_ptr(), 4, lay["back_q"][par], bound, s)
        area.signal(self.CH_SBACK, p, s)
        area.wait(self.CH_SBACK, p, s)
        t._mark("sample: leaves back to the querying rank")
        if m:
            area.slices_unpermute(lay["mat_s"][par], self.qshard.data_ptr(), self.qslot.data_ptr(), m, lay["back_q"][par], 4,
                                  out.data_ptr(), leaves_per_shard=t.n_local, stream=s)
        t._mark("sample: indices to slice order")
        return out
