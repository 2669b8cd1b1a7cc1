"""Multi-GPU parity test of ShardedSumTree over NCCL (skipped on boxes with fewer than 2 GPUs):
the sharded tree must be bit-identical to ONE oracle tree over all the leaves."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, dt_name, logn, q, peer=True):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle
        from starr_b200 import ShardedSumTree
        dt = np.dtype(dt_name)
        tdt = getattr(torch, dt_name)
        n = 1 << logn
        rng = np.random.default_rng(7)
        draw = (lambda k: np.exp(rng.uniform(-8, 8, k)).astype(dt)) if dt.kind == "f" else \
            (lambda k: rng.integers(0, 1000, k).astype(dt))
        leaves = draw(n)
        ref_leaves, ref_tree = leaves.copy(), np.zeros(n, dt)
        oracle.c.build_sumtree_from_array(ref_leaves, ref_tree)
        nl = n // world
        dev = torch.device("cuda", rank)
        st = ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[rank * nl:(rank + 1) * nl].copy()).to(dev), peer=peer)
        ok = st.gather_sumtree().cpu().numpy().tobytes() == ref_tree.tobytes()
        for m, nv in ((1, 1), (4096, 4096), (100_000, 777), (1 << 18, 1 << 18)):
            idx = rng.integers(0, n, m).astype(np.int64)
            vals = draw(nv)
            oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
            st.update_(torch.from_numpy(idx).to(dev), torch.from_numpy(vals).to(dev), check=True)
            ok &= st.gather_sumtree().cpu().numpy().tobytes() == ref_tree.tobytes()
            ok &= st.gather_leaves().cpu().numpy().tobytes() == ref_leaves.tobytes()
            u = rng.random(50_000)
            want = np.zeros(u.size, np.intp)
            oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
            got = st.sample_uniform(torch.from_numpy(u).to(dev)).cpu().numpy()
            ok &= bool(np.array_equal(got, want))
        st.poll()
        ok &= (st._peer is not None) == bool(peer)
        # slices layout: every rank supplies its 1/world of the same batches (without peer memory the slices are
        # all-gathered and applied as a replicated batch)
        for ms, nq in ((1, 2), (3000, 1000), (1 << 16, 20_000)):
            idx = rng.integers(0, n, ms * world).astype(np.int64)
            vals = draw(ms * world)
            oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
            st.update_slices_(torch.from_numpy(idx[rank * ms:(rank + 1) * ms]).to(dev),
                              torch.from_numpy(vals[rank * ms:(rank + 1) * ms]).to(dev))
            ok &= st.gather_sumtree().cpu().numpy().tobytes() == ref_tree.tobytes()
            ok &= st.gather_leaves().cpu().numpy().tobytes() == ref_leaves.tobytes()
            u = rng.random(nq * world)
            want = np.zeros(u.size, np.intp)
            oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
            got = st.sample_slices(torch.from_numpy(u[rank * nq:(rank + 1) * nq]).to(dev)).cpu().numpy()
            ok &= bool(np.array_equal(got, want[rank * nq:(rank + 1) * nq]))
        st.poll()
        st.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer", [True, False], ids=["peer", "allgather"])
@pytest.mark.parametrize("dt_name", ["float32", "int32", "float64"])
def test_sharded_nccl_equals_single_tree(dt_name, peer):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (world.bit_length() - 1)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + len(dt_name)) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port + (500 if peer else 0), dt_name, 22, q, peer)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok in results), results


# ---- single-GPU coverage of the sharded building blocks --------------------------------------
def _plan_reference(ids, shift, shards, rank, payload):
    owner = ids >> shift
    slot = np.full(ids.size, -1, np.int32)
    counts = []
    for s in range(shards):
        pos = np.flatnonzero(owner == s)
        slot[pos] = np.arange(pos.size, dtype=np.int32)
        counts.append(int(pos.size))
    mine = np.flatnonzero(owner == rank)
    return owner.astype(np.uint8), slot, ids[mine] - (rank << shift), payload[mine % payload.size], counts


@pytest.mark.parametrize("shards,rank,m,npay,skew", [
    (2, 1, 1, 1, False), (8, 3, 2047, 2047, False), (8, 0, 2049, 5, False), (64, 63, 100_003, 100_003, False),
    (4, 2, 1 << 20, 1 << 20, True), (8, 7, 300_000, 1000, True)])
def test_shard_plan_matches_numpy(shards, rank, m, npay, skew):
    from starr_b200 import DeviceSumTree
    dev = torch.device("cuda", 0)
    shift = 10
    rng = np.random.default_rng(m + shards)
    ids = rng.integers(0, shards << shift, m).astype(np.int64)
    if skew:                                             # most of the batch lands in one shard
        ids[rng.random(m) < 0.9] = (rank << shift) + 5
    payload = rng.random(npay).astype(np.float32)
    tree = DeviceSumTree(1 << shift, dtype=torch.float32, device=dev)
    shard, slot, own_ids, own_pay, counts = tree.shard_plan(
        torch.from_numpy(ids).to(dev), shift, shards, rank, torch.from_numpy(payload).to(dev), True)
    w_shard, w_slot, w_ids, w_pay, w_counts = _plan_reference(ids, shift, shards, rank, payload)
    assert counts == w_counts
    c = counts[rank]
    assert np.array_equal(shard.cpu().numpy(), w_shard)
    assert np.array_equal(slot.cpu().numpy(), w_slot)
    assert np.array_equal(own_ids[:c].cpu().numpy(), w_ids)
    assert own_pay[:c].cpu().numpy().tobytes() == w_pay.tobytes()
    tree.poll()
    # expansion of per-shard slabs back to batch order
    cap = max(counts) + 3
    slabs = torch.from_numpy(rng.random((shards, cap)).astype(np.float32)).to(dev)
    got = tree.shard_expand_values(shard, slot, slabs).cpu().numpy()
    assert got.tobytes() == slabs.cpu().numpy()[w_shard, w_slot].tobytes()
    islabs = torch.from_numpy(rng.integers(0, 1 << shift, (shards, cap)).astype(np.int32)).to(dev)
    goti = tree.shard_expand_indices(shard, slot, islabs, 1 << shift).cpu().numpy()
    assert np.array_equal(goti, w_shard.astype(np.int64) * (1 << shift) + islabs.cpu().numpy()[w_shard, w_slot])


def test_async_plan_and_device_counted_search(oracle_mod=None):
    """wait=False only enqueues the plan; the local search then reads its count on the device, and the
    host collects the counts afterwards (what ShardedSumTree.sample_uniform does)."""
    import oracle
    from starr_b200 import DeviceSumTree
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(31)
    n, m, shards, rank = 1 << 12, 50_000, 8, 5
    leaves = rng.random(n).astype(np.float32)
    tree_np = np.zeros(n, np.float32)
    oracle.c.build_sumtree_from_array(leaves, tree_np)
    tree = DeviceSumTree(torch.from_numpy(leaves).to(dev))
    owner = rng.integers(0, shards, m).astype(np.int64)
    resid = (rng.random(m) * tree_np[1]).astype(np.float32)
    bufs = (torch.empty(m, dtype=torch.uint8, device=dev), torch.empty(m, dtype=torch.int32, device=dev),
            torch.full((m,), -7, dtype=torch.int64, device=dev), torch.empty(m, dtype=torch.float32, device=dev))
    shard, slot, _, own, counts = tree.shard_plan(torch.from_numpy(owner).to(dev), 0, shards, rank,
                                                  torch.from_numpy(resid).to(dev), False, want_ids=False, out=bufs,
                                                  result_slot=2, wait=False)
    assert counts is None
    tree.search_counted(own, 2, rank, bufs[2])
    counts = tree.shard_plan_wait(2, shards)
    mine = np.flatnonzero(owner == rank)
    assert counts == [int((owner == s).sum()) for s in range(shards)]
    want = np.zeros(mine.size, np.intp)
    oracle.c.get_prefix_sum_idx(want, resid[mine], leaves, tree_np)
    got = bufs[2].cpu().numpy()
    assert np.array_equal(got[:mine.size], want)
    assert (got[mine.size:] == -7).all()                     # nothing beyond the device-side count was touched


def test_shard_plan_rejects_bad_batches():
    """The plan is synchronous, so a bad batch raises at once (the same verdict on every rank,
    since the batch is replicated) and nothing has been modified yet."""
    from starr_b200 import DeviceSumTree
    dev = torch.device("cuda", 0)
    local = DeviceSumTree(1 << 10, dtype=torch.float32, device=dev)
    ids = torch.arange(5000, dtype=torch.int64, device=dev) % (4 << 10)
    pay = torch.ones(5000, dtype=torch.float32, device=dev)
    for bad_value in (4 << 10, -1, 1 << 40):
        bad_ids = ids.clone()
        bad_ids[4321] = bad_value
        with pytest.raises(IndexError):
            local.shard_plan(bad_ids, 10, 4, 1, pay, True)
    bad_pay = pay.clone()
    bad_pay[77] = -1.0
    with pytest.raises(ValueError):
        local.shard_plan(ids, 10, 4, 1, bad_pay, True)
    with pytest.raises(ValueError):                           # unused entries of vals count too (pyx:37-39)
        local.shard_plan(ids[:10], 10, 4, 1, bad_pay, True)
    local.shard_plan(ids, 10, 4, 1, bad_pay, False)           # residuals are not vals: no sign check
    local.poll()                                              # and nothing was flagged on the tree


@pytest.mark.parametrize("dt_name", ["float32", "float64"])
def test_top_fold_uint8_ids_equal_int64_ids(dt_name):
    from starr_b200 import DeviceSumTree
    dev = torch.device("cuda", 0)
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(3)
    m, g = 50_000, 8
    roots = np.exp(rng.uniform(-2, 12, g)).astype(dt)
    shard = rng.integers(0, g, m)
    d = (np.exp(rng.uniform(-8, 8, m)) * rng.choice([-1, 1], m)).astype(dt)
    heaps = []
    for ids in (torch.from_numpy(shard.astype(np.int64)), torch.from_numpy(shard.astype(np.uint8))):
        top = DeviceSumTree(torch.from_numpy(roots).to(dev))
        recs = torch.empty(((m + 1023) // 1024) * (g - 1) * top.top_fold_record_bytes(), dtype=torch.uint8, device=dev)
        top.top_fold_maps_(ids.to(dev), torch.from_numpy(d).to(dev), 0, (m + 1023) // 1024, recs)
        top.top_fold_apply_(ids.to(dev), torch.from_numpy(d).to(dev), recs)
        top.poll()
        heaps.append(top.heap.cpu().numpy().tobytes())
    assert heaps[0] == heaps[1]
    # and both equal the sequential statement: every ancestor += d in batch order
    tree = np.zeros(g, dt)
    oracle_leaves = roots.copy()
    import oracle
    oracle.c.build_sumtree_from_array(oracle_leaves, tree)
    for j in range(m):
        h = (int(shard[j]) + g) // 2
        while h > 0:
            tree[h] = tree[h] + d[j]
            h //= 2
    assert np.frombuffer(heaps[0], dt)[:g].tobytes() == tree.tobytes()


@pytest.mark.parametrize("peer", [False, True], ids=["allgather", "peer"])
@pytest.mark.parametrize("world,dt_name", [(2, "float32"), (8, "float32"), (4, "float64"), (4, "int32")])
def test_sharded_cuda_path_on_one_gpu(world, dt_name, peer, monkeypatch):
    """All ranks as threads on cuda:0 (tests/_thread_group.py): the complete CUDA path of the
    sharded tree against ONE oracle tree, on a box with a single GPU.  ``peer``: the exchange goes through
    the peer-memory area (st_peer.cu; ranks of one process adopt each other's buffers by pointer) with every
    rank on a stream of its own, as separate processes would be."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from _thread_group import ThreadGroup

    import starr_b200.sharded as sharded
    group = ThreadGroup(world)
    monkeypatch.setattr(sharded, "dist", group)
    dev = torch.device("cuda", 0)
    dt, tdt = np.dtype(dt_name), getattr(torch, dt_name)
    n = 1 << 16
    rng = np.random.default_rng(11)
    draw = (lambda k: np.exp(rng.uniform(-8, 8, k)).astype(dt)) if dt.kind == "f" else \
        (lambda k: rng.integers(0, 1000, k).astype(dt))
    leaves = draw(n)
    ref_leaves, ref_tree = leaves.copy(), np.zeros(n, dt)
    oracle.c.build_sumtree_from_array(ref_leaves, ref_tree)
    nl = n // world
    batches = []
    for m, nv in ((1, 1), (4096, 4096), (20_000, 333), (70_000, 70_000)):
        idx = rng.integers(0, n, m).astype(np.int64)
        if m == 20_000:
            idx[rng.random(m) < 0.7] = rng.integers(0, 64, 1)[0]        # skewed: duplicates in one shard
        vals, u = draw(nv), rng.random(10_000)
        oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
        want = np.zeros(u.size, np.intp)
        oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
        batches.append((idx, vals, u, ref_tree.copy(), ref_leaves.copy(), want))

    if peer:
        # Every kernel the steps use runs once BEFORE the ranks start (first-use work of the CUDA runtime -- module
        # loading, function attributes -- may wait for the whole device, and in this harness the device then holds
        # another rank's spinning wait kernel; separate processes on separate GPUs cannot get into that state).
        from starr_b200 import DeviceSumTree
        warm = DeviceSumTree(torch.from_numpy(leaves[:nl].copy()).to(dev))
        warm.reserve(70_000)
        for idx, vals, u, *_ in batches:
            w_idx = torch.from_numpy(idx % nl).to(dev)
            warm.update_(w_idx, torch.from_numpy(vals).to(dev))
            warm.update_begin_(w_idx, torch.from_numpy(np.resize(vals, idx.size)).to(dev), torch.empty(idx.size, dtype=tdt, device=dev))
            warm.update_finish_()
            warm.sample_uniform(torch.from_numpy(u).to(dev))
        warm.poll()
        torch.cuda.synchronize()
        del warm

    import gc
    gc.collect()   # device memory of earlier tests is freed NOW (cudaFree waits for the whole device), not under a spinning wait kernel

    def rank_body(r):
        torch.cuda.set_device(0)
        if peer:        # a spinning wait kernel must not sit in front of another rank's signal on a shared stream
            with torch.cuda.stream(torch.cuda.Stream()):
                try:
                    return rank_steps(r)
                finally:
                    torch.cuda.current_stream().synchronize()
        return rank_steps(r)

    def rank_steps(r):
        st = sharded.ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[r * nl:(r + 1) * nl].copy()).to(dev),
                                    peer=peer)
        st.reserve(70_000)   # no workspace growth (cudaFree = device-wide wait) while another rank's wait kernel spins
        bad_steps = []
        side = torch.cuda.Stream()
        for b, (idx, vals, u, tree_after, leaves_after, want) in enumerate(batches):
            d_idx, d_vals = torch.from_numpy(idx).to(dev), torch.from_numpy(vals).to(dev)
            if b % 2:                      # plan ahead on a side stream (inputs are complete: .to() above is synchronous)
                torch.cuda.current_stream().synchronize()
                st.update_(plan=st.plan_update(d_idx, d_vals, stream=side))
            else:
                st.update_(d_idx, d_vals)
            if st.gather_sumtree().cpu().numpy().tobytes() != tree_after.tobytes():
                bad_steps.append((b, "tree"))
            if st.gather_leaves().cpu().numpy().tobytes() != leaves_after.tobytes():
                bad_steps.append((b, "leaves"))
            if not np.array_equal(st.sample_uniform(torch.from_numpy(u).to(dev)).cpu().numpy(), want):
                bad_steps.append((b, "sample"))
        st.poll()
        # a bad batch raises on every rank alike, before anything is modified
        before = st.gather_sumtree().cpu().numpy().tobytes()
        bad = torch.from_numpy(batches[1][0]).to(dev).clone()
        bad[17] = n
        try:
            st.update_(bad, torch.from_numpy(batches[1][1]).to(dev))
            bad_steps.append(("bad batch", "not reported"))
        except IndexError:
            pass
        st.poll()
        if st.gather_sumtree().cpu().numpy().tobytes() != before:
            bad_steps.append(("bad batch", "tree modified"))
        if peer and st._peer is None:
            bad_steps.append(("peer", "exchange area was not used"))
        st.close()
        return bad_steps

    assert group.run(rank_body) == [[]] * world


@pytest.mark.parametrize("world,dt_name", [(2, "float32"), (8, "float32"), (4, "float64"), (4, "int32")])
def test_sharded_slices_on_one_gpu(world, dt_name, monkeypatch):
    """Slices layout (starr_b200/_slices.py): every rank supplies 1/G of the batch; the tree, the leaves and the sampled
    indices must equal ONE oracle tree that applies the concatenated batch.  Ranks are threads on cuda:0, each on a stream
    of its own; odd batches are planned ahead on a side stream."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from _thread_group import ThreadGroup

    import starr_b200.sharded as sharded
    group = ThreadGroup(world)
    monkeypatch.setattr(sharded, "dist", group)
    # 8 ranks on ONE device: one stream per rank (three each would share the device's hardware queues, and a spinning
    # wait kernel at the head of a queue holds up whatever else was mapped to it -- separate processes on separate GPUs
    # have the queues to themselves); the side-stream schedule is covered by the 2- and 4-rank cases
    one_stream = world > 4
    if one_stream:
        monkeypatch.setenv("STARR_B200_SLICES_SERIAL", "1")
    dev = torch.device("cuda", 0)
    dt, tdt = np.dtype(dt_name), getattr(torch, dt_name)
    n = 1 << 16
    rng = np.random.default_rng(12)
    draw = (lambda k: np.exp(rng.uniform(-8, 8, k)).astype(dt)) if dt.kind == "f" else \
        (lambda k: rng.integers(0, 1000, k).astype(dt))
    leaves = draw(n)
    ref_leaves, ref_tree = leaves.copy(), np.zeros(n, dt)
    oracle.c.build_sumtree_from_array(ref_leaves, ref_tree)
    nl = n // world
    batches = []
    for m, mq in ((1, 3), (512, 100), (2500, 2500), (9000, 5000)):       # per rank
        idx = rng.integers(0, n, m * world).astype(np.int64)
        if m == 2500:
            idx[rng.random(idx.size) < 0.7] = rng.integers(0, 64, 1)[0]    # skewed: most of the batch is ONE leaf of shard 0
        if m == 9000:
            idx[: m] = rng.integers(n - nl, n, m)                          # the first rank's slice all belongs to the last shard
        vals, u = draw(m * world), rng.random(mq * world)
        oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
        want = np.zeros(u.size, np.intp)
        oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
        batches.append((m, mq, idx, vals, u, ref_tree.copy(), ref_leaves.copy(), want))

    # every kernel runs once BEFORE the ranks start (see test_sharded_cuda_path_on_one_gpu)
    from starr_b200 import DeviceSumTree
    warm = DeviceSumTree(torch.from_numpy(leaves[:nl].copy()).to(dev))
    warm.reserve(9000 * world)
    for m, mq, idx, vals, u, *_ in batches:
        w_idx, w_vals = torch.from_numpy(idx % nl).to(dev), torch.from_numpy(vals).to(dev)
        warm.update_begin_(w_idx, w_vals, torch.empty(idx.size, dtype=tdt, device=dev))
        warm.update_finish_()
        warm.sample_uniform(torch.from_numpy(u).to(dev))
    warm.poll()
    torch.cuda.synchronize()
    del warm

    import gc
    gc.collect()   # device memory of earlier tests is freed NOW (cudaFree waits for the whole device), not under a spinning wait kernel

    def rank_body(r):
        torch.cuda.set_device(0)
        with torch.cuda.stream(torch.cuda.Stream()):
            try:
                return rank_steps(r)
            finally:
                torch.cuda.current_stream().synchronize()

    def rank_steps(r):
        st = sharded.ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[r * nl:(r + 1) * nl].copy()).to(dev), peer=True)
        st.reserve_slices(9000, expected_local=9000 * world)   # no workspace growth while another rank's wait kernel spins
        bad_steps = []
        side = torch.cuda.Stream()
        for b, (m, mq, idx, vals, u, tree_after, leaves_after, want) in enumerate(batches):
            d_idx = torch.from_numpy(idx[r * m:(r + 1) * m]).to(dev)
            d_vals = torch.from_numpy(vals[r * m:(r + 1) * m]).to(dev)
            if b % 2 and not one_stream:
                torch.cuda.current_stream().synchronize()
                st.update_slices_(plan=st.plan_update_slices(d_idx, d_vals, stream=side))
            else:
                st.update_slices_(d_idx, d_vals)
            if st.gather_sumtree().cpu().numpy().tobytes() != tree_after.tobytes():
                bad_steps.append((b, "tree"))
            if st.gather_leaves().cpu().numpy().tobytes() != leaves_after.tobytes():
                bad_steps.append((b, "leaves"))
            got = st.sample_slices(torch.from_numpy(u[r * mq:(r + 1) * mq]).to(dev)).cpu().numpy()
            if not np.array_equal(got, want[r * mq:(r + 1) * mq]):
                bad_steps.append((b, "sample"))
        # a bad slice on ONE rank raises on every rank, before anything is modified
        before = st.gather_sumtree().cpu().numpy().tobytes()
        m, _, idx, vals = batches[1][:4]
        bad = torch.from_numpy(idx[r * m:(r + 1) * m]).to(dev).clone()
        if r == world - 1:
            bad[17] = n
        try:
            st.update_slices_(bad, torch.from_numpy(vals[r * m:(r + 1) * m]).to(dev))
            bad_steps.append(("bad batch", "not reported"))
        except IndexError:
            pass
        if st.gather_sumtree().cpu().numpy().tobytes() != before:
            bad_steps.append(("bad batch", "tree modified"))
        # an empty batch and an empty query set are legal (every rank still takes part in the exchange)
        st.update_slices_(torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=tdt, device=dev))
        if st.sample_slices(torch.zeros(0, dtype=torch.float64, device=dev)).numel() != 0:
            bad_steps.append(("empty", "sample"))
        if st.gather_sumtree().cpu().numpy().tobytes() != before:
            bad_steps.append(("empty batch", "tree modified"))
        # and the tree keeps working afterwards
        m, mq, idx, vals, u, *_ = batches[0]
        st.update_slices_(torch.from_numpy(idx[r * m:(r + 1) * m]).to(dev), torch.from_numpy(vals[r * m:(r + 1) * m]).to(dev))
        if st._slices is None or st._slices.area is None:
            bad_steps.append(("slices", "exchange area was not used"))
        st.close()
        return bad_steps

    assert group.run(rank_body) == [[]] * world
