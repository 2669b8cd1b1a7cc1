"""GPU tests of the batched torch entry point (DeviceSumTree): PER-loop parity against the
oracle at BASELINE config #2, deferred error reporting, and the full-size config #3
(2^27 float32 leaves, 2^20 updates + 2^20 samples) checked bit-for-bit against the oracle and
through size-independent properties."""
import numpy as np
import pytest
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("dt_name", ["float32", "float64", "int32", "int64"])
def test_per_loop_parity(torch_mod, oracle_mod, dt_name):
    """BASELINE config #2: 2^20 leaves, 4096 updates + 4096 samples per iteration, where the
    update indices are the previous step's samples (skewed, many duplicates)."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    dt = np.dtype(dt_name)
    tdt = getattr(torch, dt_name)
    n, m = 1 << 20, 4096
    rng = np.random.default_rng(0)
    leaves = rng.random(n).astype(dt) if dt.kind == "f" else rng.integers(0, 1000, n).astype(dt)
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
    assert_bits_equal(_np(dev.tree), tree, "build")
    idx = rng.integers(0, n, m).astype(np.intp)
    for it in range(6):
        vals = (rng.random(m) ** 4 * 3).astype(dt) if dt.kind == "f" else rng.integers(0, 1000, m).astype(dt)
        C.update_sumtree(idx, vals, leaves, tree)
        dev.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda())
        u = rng.random(m)
        q = (tree[1] * u).astype(dt)
        want = np.zeros(m, np.intp)
        C.get_prefix_sum_idx(want, q, leaves, tree)
        got, prio = dev.sample_uniform(torch.from_numpy(u).cuda(), return_priorities=True)
        assert np.array_equal(_np(got), want), f"iteration {it}"
        assert_bits_equal(_np(prio), leaves[want], "sampled priorities")
        idx = want.copy()       # next step updates what was just sampled
    dev.poll()
    assert_bits_equal(_np(dev.tree), tree, "tree after PER loop")
    assert_bits_equal(_np(dev.leaves), leaves, "leaves after PER loop")


def test_deferred_errors(torch_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    dev = DeviceSumTree(torch.arange(16, dtype=torch.float32).cuda())
    before = dev.heap.clone()
    dev.update_(torch.tensor([1, 2], device="cuda"), torch.tensor([1.0, -2.0], device="cuda"))
    with pytest.raises(ValueError, match="vals must be non-negative"):
        dev.poll()
    assert torch.equal(dev.heap, before)          # the failed batch was skipped as a whole
    with pytest.raises(IndexError):
        dev.update_(torch.tensor([1, 16], device="cuda"), torch.tensor([1.0], device="cuda"), check=True)
    assert torch.equal(dev.heap, before)
    dev.update_(torch.tensor([3], device="cuda"), torch.tensor([10.0], device="cuda"), check=True)
    assert dev.total().item() == 120 + 7
    with pytest.raises(ValueError, match="array elements must be non-negative"):
        dev.build_(-torch.ones(16, device="cuda"))
    with pytest.raises(TypeError):
        dev.update_(torch.tensor([3], device="cuda"), torch.tensor([10.0], device="cuda", dtype=torch.float64))


def test_sums_on_device(torch_mod, oracle_mod):
    """BASELINE config #5 shapes at reduced size plus the real 4096x4096 aligned case."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    rng = np.random.default_rng(2)
    for dt, tdt in ((np.float32, torch.float32), (np.int32, torch.int32)):
        x = rng.random((4096, 4096)).astype(dt) if dt == np.float32 else rng.integers(0, 1000, (4096, 4096)).astype(dt)
        dev = DeviceSumTree(torch.from_numpy(x).cuda())
        tree = _np(dev.tree)
        flat = x.ravel()
        # axis=1: aligned stride -> exactly the nodes 4096..8191
        assert_bits_equal(_np(dev.strided_sum(4096)), tree[4096:8192], "axis=1 aligned")
        assert_bits_equal(_np(dev.strided_sum(4096)), C.strided_sum(flat, tree, 4096), "axis=1 vs oracle")
        assert_bits_equal(_np(dev.strided_sum(1000)), C.strided_sum(flat, tree, 1000), "ragged stride")
        # axis=0: sequential row fold == ndarray.sum(axis=0) bit for bit (int32 -> int64)
        assert_bits_equal(_np(dev.axis_fold(1, 4096, 4096)).ravel(), x.sum(axis=0), "axis=0")
        # nd-index sampling: flat indices unravel like the reference (sumtree_array.py:360-363)
        u = rng.random(1 << 16)
        got = _np(dev.sample_uniform(torch.from_numpy(u).cuda()))
        want = np.zeros(u.size, np.intp)
        C.get_prefix_sum_idx(want, (tree[1] * u).astype(dt), flat, tree)
        assert np.array_equal(got, want)


def test_full_size_config3(torch_mod, oracle_mod):
    """2^27 float32 leaves, 2^20 updates + 2^20 samples: bit-exact tree / leaves / indices vs the
    oracle, plus structural properties that do not need the oracle."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    n, m = 1 << 27, 1 << 20
    rng = np.random.default_rng(0)
    leaves = rng.random(n, dtype=np.float32)
    tree = np.zeros(n, np.float32)
    C.build_sumtree_from_array(leaves, tree)
    dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
    # property: every node is exactly left + right after a build
    heap = dev.heap
    assert torch.equal(heap[1:n], heap[2:2 * n:2] + heap[3:2 * n:2])
    assert_bits_equal(_np(dev.tree), tree, "build 2^27")
    idx = rng.integers(0, n, m).astype(np.intp)
    vals = rng.random(m, dtype=np.float32)
    C.update_sumtree(idx, vals, leaves, tree)
    dev.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda(), check=True)
    assert_bits_equal(_np(dev.leaves), leaves, "leaves after 2^20 updates")
    assert_bits_equal(_np(dev.tree), tree, "tree after 2^20 updates")
    u = np.sort(rng.random(m))
    got = _np(dev.sample_uniform(torch.from_numpy(u).cuda()))
    want = np.zeros(m, np.intp)
    C.get_prefix_sum_idx(want, (tree[1] * u).astype(np.float32), leaves, tree)
    assert np.array_equal(got, want)
    # property: sorted queries land on non-decreasing leaves (power-of-two tree: no rotation)
    assert (np.diff(got) >= 0).all()
    # idempotence: re-applying the same batch changes nothing that a second oracle pass would not
    C.update_sumtree(idx, vals, leaves, tree)
    dev.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda())
    assert_bits_equal(_np(dev.tree), tree, "tree after re-applying the batch")


@pytest.mark.parametrize("dt_name", ["float32", "float64"])
def test_apply_deltas_tiny_tree(torch_mod, dt_name):
    """The sharded layout's replicated top tree: every internal node of a tiny tree folds ALL the
    differences below it in batch order (masked ordered fold, no sort / partition)."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(11)
    for G in (2, 4, 8):
        for m in (100, 5000, 300_000):
            t = DeviceSumTree(torch.from_numpy(np.exp(rng.uniform(-3, 12, G)).astype(dt)).cuda())
            heap0 = _np(t.heap).copy()
            leaf = rng.integers(0, G, m).astype(np.int64)
            d = (np.exp(rng.uniform(-12, 6, m)) * rng.choice([-1.0, 1.0], m)).astype(dt)
            t.apply_deltas_(torch.from_numpy(leaf).cuda(), torch.from_numpy(d).cuda())
            want = heap0.copy()
            for h in range(1, G):
                lvl = h.bit_length() - 1
                member = ((leaf + G) >> (int(np.log2(G)) - lvl)) == h
                seq = np.concatenate([[heap0[h]], d[member]]).astype(dt)
                want[h] = np.add.accumulate(seq, dtype=dt)[-1]        # sequential fold in dtype
            assert_bits_equal(_np(t.heap), want, f"G={G} m={m}")


@pytest.mark.parametrize("dt_name", ["int32", "int64", "uint8", "int16"])
def test_apply_deltas_integer_trees_unsorted_ids(torch_mod, dt_name):
    """Integer ``apply_deltas_``: ids arrive in batch order, so equal ancestors are NOT adjacent."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(19)
    for n in (2, 4, 8, 64, 1000):
        for m in (1, 37, 5000, 100_000):
            leaves = rng.integers(0, 50, n).astype(dt)
            t = DeviceSumTree(torch.from_numpy(leaves).cuda())
            want = _np(t.heap).copy()
            leaf = rng.integers(0, n, m).astype(np.int64)
            d = rng.integers(-3 if dt.kind == "i" else 0, 4, m).astype(dt)
            t.apply_deltas_(torch.from_numpy(leaf).cuda(), torch.from_numpy(d).cuda())
            for j in range(m):                       # wrap-around arithmetic in the dtype, like C
                h = (int(leaf[j]) + n) // 2
                while h > 0:
                    want[h] = (want[h].astype(np.int64) + d[j].astype(np.int64)).astype(dt) if dt.itemsize < 8 \
                        else want[h] + d[j]
                    h //= 2
            assert_bits_equal(_np(t.heap), want, f"{dt_name} n={n} m={m}")


def test_per_step_cuda_graph(torch_mod, oracle_mod):
    """The PER step (update -> sample with priorities) captured ONCE in a CUDA graph and replayed with
    new inputs in static buffers: same bits as the oracle, no launches from Python per step."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    n, m = 1 << 20, 4096
    rng = np.random.default_rng(5)
    leaves = rng.random(n).astype(np.float32)
    tree = np.zeros(n, np.float32)
    C.build_sumtree_from_array(leaves, tree)
    dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
    dev.reserve(m)
    idx_b = torch.zeros(m, dtype=torch.int64, device="cuda")
    val_b = torch.zeros(m, dtype=torch.float32, device="cuda")
    u_b = torch.zeros(m, dtype=torch.float64, device="cuda")
    out_b = torch.zeros(m, dtype=torch.int64, device="cuda")

    def feed(idx, val, u):
        idx_b.copy_(torch.from_numpy(idx)); val_b.copy_(torch.from_numpy(val)); u_b.copy_(torch.from_numpy(u))

    # warm-up outside capture (first-use attribute setting, workspace), on a side stream as torch requires
    idx = rng.integers(0, n, m).astype(np.int64); val = rng.random(m).astype(np.float32); u = rng.random(m)
    feed(idx, val, u)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dev.update_(idx_b, val_b)
        _, prio_w = dev.sample_uniform(u_b, out=out_b, return_priorities=True)
    torch.cuda.current_stream().wait_stream(s)
    C.update_sumtree(idx.astype(np.intp), val, leaves, tree)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        dev.update_(idx_b, val_b)
        _, prio_b = dev.sample_uniform(u_b, out=out_b, return_priorities=True)
    # the capture itself did not run anything; replay with fresh inputs several times
    for it in range(4):
        idx = rng.integers(0, n, m).astype(np.int64) if it % 2 == 0 else want.astype(np.int64)
        val = (rng.random(m) ** 3).astype(np.float32); u = rng.random(m)
        feed(idx, val, u)
        graph.replay()
        C.update_sumtree(idx.astype(np.intp), val, leaves, tree)
        want = np.zeros(m, np.intp)
        C.get_prefix_sum_idx(want, (tree[1] * u).astype(np.float32), leaves, tree)
        assert np.array_equal(_np(out_b), want), f"replay {it}"
        assert_bits_equal(_np(prio_b), leaves[want], "priorities")
    dev.poll()
    assert_bits_equal(_np(dev.tree), tree, "tree after graph replays")
