"""World-size-2 (and 4) gloo tests of the subtree-sharded tree's host logic, on CPU.

The local tree is played by an oracle-backed stand-in (tests/_cpu_tree.py) injected through
ShardedSumTree's tree_factory; what is under test is everything ShardedSumTree itself does:
ownership filtering in batch order, `vals` cycling over the GLOBAL batch position, the exact
delta exchange, the ordered top-tree fold, query routing and result combination.  The bar is the
same as on one GPU: bit-identical to ONE oracle tree over all the leaves."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, dt_name, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from _cpu_tree import CpuTree

        from starr_b200.sharded import ShardedSumTree
        dt = np.dtype(dt_name)
        tdt = getattr(torch, dt_name)
        rng = np.random.default_rng(42)                      # replicated inputs
        draw = (lambda k: np.exp(rng.uniform(-8, 8, k)).astype(dt)) if dt.kind == "f" else \
            (lambda k: rng.integers(0, 1000, k).astype(dt))
        leaves = draw(n)
        ref_leaves, ref_tree = leaves.copy(), np.zeros(n, dt)
        oracle.c.build_sumtree_from_array(ref_leaves, ref_tree)
        nl = n // world
        st = ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[rank * nl:(rank + 1) * nl].copy()),
                            tree_factory=lambda n_, d_, dev: CpuTree(n_, d_))
        ok = st.gather_sumtree().numpy().tobytes() == ref_tree.tobytes()
        for b, (m, nv) in enumerate(((1, 1), (257, 257), (3000, 7), (5000, 5000))):
            idx = rng.integers(0, n if b != 3 else n // 8, m).astype(np.int64)     # batch 3: one shard only
            vals = draw(nv)
            oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
            st.update_(torch.from_numpy(idx), torch.from_numpy(vals), check=True)
            ok &= st.gather_sumtree().numpy().tobytes() == ref_tree.tobytes()
            ok &= st.gather_leaves().numpy().tobytes() == ref_leaves.tobytes()
            u = rng.random(2000)
            want = np.zeros(u.size, np.intp)
            oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
            got = st.sample_uniform(torch.from_numpy(u)).numpy()
            ok &= bool(np.array_equal(got, want))
        # edge cases of the host logic: empty batch, empty vals, empty query set, an explicit plan
        before = st.gather_sumtree().numpy().tobytes()
        st.update_(torch.zeros(0, dtype=torch.int64), torch.from_numpy(draw(3)))
        ok &= st.gather_sumtree().numpy().tobytes() == before
        try:
            st.update_(torch.tensor([1, 2]), torch.from_numpy(draw(0)))
            ok = False
        except ZeroDivisionError:
            pass
        ok &= st.sample_uniform(torch.zeros(0, dtype=torch.float64)).numel() == 0
        idx = rng.integers(0, n, 100).astype(np.int64)
        vals = draw(100)
        oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
        plan = st.plan_update(torch.from_numpy(idx), torch.from_numpy(vals))
        ok &= sum(plan.counts) == 100 and plan.ni == 100
        st.update_(plan=plan)
        ok &= st.gather_sumtree().numpy().tobytes() == ref_tree.tobytes()
        ok &= float(st.total()) == float(ref_tree[1])
        # slices layout (every rank supplies its 1/world of the batch); without peer memory -- here -- the slices are
        # all-gathered and applied as one replicated batch in rank order
        for ms, nq in ((1, 1), (300, 50)):
            idx = rng.integers(0, n, ms * world).astype(np.int64)
            vals = draw(ms * world)
            oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_leaves, ref_tree)
            st.update_slices_(torch.from_numpy(idx[rank * ms:(rank + 1) * ms]), torch.from_numpy(vals[rank * ms:(rank + 1) * ms]))
            ok &= st.gather_sumtree().numpy().tobytes() == ref_tree.tobytes()
            ok &= st.gather_leaves().numpy().tobytes() == ref_leaves.tobytes()
            u = rng.random(nq * world)
            want = np.zeros(u.size, np.intp)
            oracle.c.get_prefix_sum_idx(want, (ref_tree[1] * u).astype(dt), ref_leaves, ref_tree)
            got = st.sample_slices(torch.from_numpy(u[rank * nq:(rank + 1) * nq])).numpy()
            ok &= bool(np.array_equal(got, want[rank * nq:(rank + 1) * nq]))
        try:
            st.update_(torch.tensor([0]), torch.from_numpy(np.array([1], dt)) * -1 if dt.kind != "u" else torch.tensor([n]), check=True)
            raised = dt.kind == "u"
        except (ValueError, IndexError, TypeError):
            raised = True
        q.put((rank, bool(ok), raised))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dt_name", [(2, "float32"), (2, "int32"), (4, "float32"), (2, "float64")])
def test_sharded_equals_single_tree(world, dt_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + len(dt_name)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, dt_name, 1 << 10, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _ in results), results
    assert all(r for _, _, r in results), results
