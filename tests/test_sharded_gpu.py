"""Multi-GPU parity test of ShardedSumTree over NCCL (skipped on boxes with fewer than 2 GPUs):
the sharded tree must be bit-identical to ONE oracle tree over all the leaves."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, dt_name, logn, q):
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
        st = ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[rank * nl:(rank + 1) * nl].copy()).to(dev))
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
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dt_name", ["float32", "int32", "float64"])
def test_sharded_nccl_equals_single_tree(dt_name):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (world.bit_length() - 1)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + len(dt_name)) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, dt_name, 22, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok in results), results
