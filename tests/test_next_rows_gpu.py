"""GPU tests of the SURVEY 8(f) rows: fused PER step with priorities / probabilities (f1), device-side
in-place arithmetic + rebuild (f2), MinTree + ring append (f3), the starr.experimental shim (f4), the
pipelined host step of the C ABI, and the per-batch scoping of the validation status."""
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


def _mk(rng, n, dt):
    dt = np.dtype(dt)
    return rng.random(n).astype(dt) if dt.kind == "f" else rng.integers(0, 1000, n).astype(dt)


# ---- f1: fused PER step ---------------------------------------------------------------------------
@pytest.mark.parametrize("dt_name", ["float32", "float64", "int32"])
@pytest.mark.parametrize("n,m", [(1 << 16, 4096), (100003, 1000), (1 << 20, 1 << 17)])
def test_per_step_matches_oracle_and_numpy(torch_mod, oracle_mod, dt_name, n, m):
    """update -> sample -> priorities -> p_i / total in one call == oracle update + oracle search +
    NumPy's x[idx] / x.sum() (reference README.md:7; sumtree_array.py:181-182, :503)."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(n + m)
    leaves = _mk(rng, n, dt)
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
    dev.reserve(m)
    idx = rng.integers(0, n, m).astype(np.intp)
    for it in range(3):
        vals = _mk(rng, m, dt)
        u = rng.random(m)
        C.update_sumtree(idx, vals, leaves, tree)
        want = np.zeros(m, np.intp)
        C.get_prefix_sum_idx(want, (tree[1] * u).astype(dt), leaves, tree)
        got, prio, prob = dev.per_step_(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda(),
                                        torch.from_numpy(u).cuda())
        assert np.array_equal(_np(got), want), f"indices, step {it}"
        assert_bits_equal(_np(prio), leaves[want], "priorities")
        assert_bits_equal(_np(prob), leaves[want] / tree[1], "probabilities = x[idx] / x.sum()")
        idx = want.copy()
    dev.poll()
    assert_bits_equal(_np(dev.tree), tree, "tree")
    # sample_weights alone, into caller buffers
    u = rng.random(257)
    bufs = (torch.empty(300, dtype=torch.int64, device="cuda"), torch.empty(300, dtype=dev.dtype, device="cuda"),
            torch.empty(300, dtype=dev.prob_dtype, device="cuda"))
    i2, p2, w2 = dev.sample_weights(torch.from_numpy(u).cuda(), out=bufs)
    want = np.zeros(257, np.intp)
    C.get_prefix_sum_idx(want, (tree[1] * u).astype(dt), leaves, tree)
    assert np.array_equal(_np(i2)[:257], want)
    assert_bits_equal(_np(w2)[:257], leaves[want] / tree[1], "probabilities")


def test_per_step_cuda_graph_library_call(torch_mod, oracle_mod):
    """The library-level step is capturable: replaying the graph with fresh inputs stays bit-exact."""
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    n, m = 1 << 20, 4096
    rng = np.random.default_rng(3)
    leaves = rng.random(n, dtype=np.float32)
    tree = np.zeros(n, np.float32)
    C.build_sumtree_from_array(leaves, tree)
    dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
    dev.reserve(m)
    d_idx = torch.zeros(m, dtype=torch.int64, device="cuda")
    d_val = torch.zeros(m, dtype=torch.float32, device="cuda")
    d_u = torch.zeros(m, dtype=torch.float64, device="cuda")
    outs = (torch.empty(m, dtype=torch.int64, device="cuda"), torch.empty(m, dtype=torch.float32, device="cuda"),
            torch.empty(m, dtype=torch.float32, device="cuda"))

    def feed():
        idx = rng.integers(0, n, m).astype(np.intp); val = rng.random(m, dtype=np.float32); u = rng.random(m)
        d_idx.copy_(torch.from_numpy(idx)); d_val.copy_(torch.from_numpy(val)); d_u.copy_(torch.from_numpy(u))
        C.update_sumtree(idx, val, leaves, tree)
        want = np.zeros(m, np.intp)
        C.get_prefix_sum_idx(want, (tree[1] * u).astype(np.float32), leaves, tree)
        return want

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    want = feed()
    with torch.cuda.stream(side):
        dev.per_step_(d_idx, d_val, d_u, out=outs)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert np.array_equal(_np(outs[0]), want)
    want = feed()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        dev.per_step_(d_idx, d_val, d_u, out=outs)
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(_np(outs[0]), want)
    for _ in range(3):
        want = feed()
        graph.replay()
        torch.cuda.synchronize()
        assert np.array_equal(_np(outs[0]), want)
        assert_bits_equal(_np(outs[2]), leaves[want] / tree[1], "probabilities")
    assert_bits_equal(_np(dev.tree), tree, "tree after graph replays")


@pytest.mark.parametrize("NS", [2, 4])
def test_host_pipeline_step(oracle_mod, NS):
    """st_per_step_host_submit / _wait: pinned host buffers in and out, NS slots in flight."""
    from starr_b200 import _lib
    C = oracle_mod.c
    n, m = 1 << 18, 1 << 14
    rng = np.random.default_rng(9)
    leaves = rng.random(n, dtype=np.float32)
    tree = np.zeros(n, np.float32)
    C.build_sumtree_from_array(leaves, tree)
    h = _lib.TreeHandle(n, np.float32)
    h.build_from_host(leaves)
    slots = []
    for _ in range(NS):
        pb = _lib.PinnedBuffer(m * (8 + 4 + 8 + 8 + 4 + 4))
        off = 0
        arrs = []
        for dt in (np.int64, np.float32, np.float64, np.int64, np.float32, np.float32):
            arrs.append(pb.array(dt, m, off)); off += m * np.dtype(dt).itemsize
        slots.append((pb, arrs))
    wants = {}
    steps = 7
    for k in range(steps + 1):
        if k < steps:
            _, (idx, val, u, o_idx, o_prio, o_prob) = slots[k % NS]
            if k >= NS:
                h.per_step_host_wait(k % NS)
                w_idx, w_prio, w_prob = wants.pop(k - NS)
                assert np.array_equal(o_idx, w_idx), f"step {k - NS}"
                assert_bits_equal(o_prio, w_prio, "priorities"); assert_bits_equal(o_prob, w_prob, "probabilities")
            idx[:] = rng.integers(0, n, m); val[:] = rng.random(m, dtype=np.float32); u[:] = rng.random(m)
            C.update_sumtree(idx.astype(np.intp), val, leaves, tree)
            want = np.zeros(m, np.intp)
            C.get_prefix_sum_idx(want, (tree[1] * u).astype(np.float32), leaves, tree)
            wants[k] = (want, leaves[want].copy(), leaves[want] / tree[1])
            h.per_step_host_submit(k % NS, idx, val, u, o_idx, o_prio, o_prob)
    for k in range(steps - NS, steps):
        h.per_step_host_wait(k % NS)
        _, (_, _, _, o_idx, o_prio, o_prob) = slots[k % NS]
        assert np.array_equal(o_idx, wants[k][0]), f"step {k}"
        assert_bits_equal(o_prob, wants[k][2], "probabilities")
    t = np.zeros(n, np.float32)
    h.store_host(None, t)
    assert_bits_equal(t, tree, "tree after pipelined host steps")
    # a bad batch is reported by wait() and skipped; the next one is applied
    _, (idx, val, u, o_idx, o_prio, o_prob) = slots[0]
    val[:] = 1; val[5] = -1
    h.per_step_host_submit(0, idx, val, u, o_idx, None, None)
    with pytest.raises(ValueError, match="vals must be non-negative"):
        h.per_step_host_wait(0)
    h.store_host(None, t)
    assert_bits_equal(t, tree, "failed batch wrote nothing")
    h.close()


# ---- status scoping (the validation flag must not be sticky across batches) -----------------
def test_failed_batch_does_not_block_later_batches(torch_mod, oracle_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    for n, m in ((1 << 12, 64), (1 << 18, 1 << 15)):      # single-CTA path and the multi-kernel path
        rng = np.random.default_rng(n)
        leaves = rng.random(n, dtype=np.float32)
        tree = np.zeros(n, np.float32)
        C.build_sumtree_from_array(leaves, tree)
        dev = DeviceSumTree(torch.from_numpy(leaves).cuda())
        idx = rng.integers(0, n, m).astype(np.intp)
        bad = rng.random(m, dtype=np.float32); bad[m // 2] = -3
        good = rng.random(m, dtype=np.float32)
        dev.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(bad).cuda())        # skipped as a whole
        dev.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(good).cuda())       # applied
        C.update_sumtree(idx, good, leaves, tree)
        got = dev.sample_uniform(torch.from_numpy(np.array([0.5])).cuda())
        with pytest.raises(ValueError, match="vals must be non-negative"):
            dev.poll()
        assert_bits_equal(_np(dev.tree), tree, "the batch behind the failed one was applied")
        assert_bits_equal(_np(dev.leaves), leaves, "leaves")
        dev.poll()                                                                          # cleared
        oob = idx.copy(); oob[3] = n
        dev.update_(torch.from_numpy(oob).cuda(), torch.from_numpy(good).cuda())
        dev.build_()                                    # must build (and not raise the stale index error)
        with pytest.raises(IndexError):
            dev.poll()
        del got


def test_out_buffers_are_validated(torch_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    dev = DeviceSumTree(torch.arange(64, dtype=torch.float32).cuda())
    u = torch.rand(16, dtype=torch.float64, device="cuda")
    for bad in (torch.empty(8, dtype=torch.int64, device="cuda"), torch.empty(16, dtype=torch.int32, device="cuda"),
                torch.empty(32, dtype=torch.int64, device="cuda")[::2], torch.empty(16, dtype=torch.int64)):
        with pytest.raises((TypeError, ValueError)):
            dev.sample_uniform(u, out=bad)
        with pytest.raises((TypeError, ValueError)):
            dev.search(torch.rand(16, device="cuda"), out=bad)
    ok = torch.empty(16, dtype=torch.int64, device="cuda")
    assert dev.sample_uniform(u, out=ok) is ok


# ---- f2: in-place arithmetic on the device -----------------------------------------------------------
@pytest.mark.parametrize("dt_name", ["float32", "float64", "int32", "uint8", "int16", "int64"])
def test_inplace_ops_match_numpy_then_oracle_build(oracle_mod, dt_name):
    """x op= c / x.fill(c) through SumTreeArray: leaves == NumPy's ufunc result, tree == oracle build of them
    (reference sumtree_array.py:150-173, :202-205)."""
    from starr_b200 import SumTreeArray
    C = oracle_mod.c
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(17)
    for shape in ((1000,), (37, 29), (4, 8, 16)):
        base = (rng.random(shape) * 5).astype(dt) if dt.kind == "f" else rng.integers(0, 9, shape).astype(dt)
        x = SumTreeArray(base)
        ref = base.copy()

        def check(what):
            tree = np.zeros(ref.size, dt)
            C.build_sumtree_from_array(np.ascontiguousarray(ref).ravel(), tree)
            assert_bits_equal(np.asarray(x[...]), ref, f"{what}: leaves")
            assert_bits_equal(x.sumtree(), tree, f"{what}: tree")
            assert x.sum() == tree[1]

        x += 3; ref += 3; check("+= int")
        x *= 2; ref *= 2; check("*= int")
        x -= 1; ref -= 1; check("-= int")
        x += dt.type(2); ref += dt.type(2); check("+= numpy scalar of the same dtype")
        if dt.kind == "f":
            x += 0.1; ref += 0.1; check("+= float")
            x *= 1.7; ref *= 1.7; check("*= float")
            x /= 3; ref /= 3; check("/= int")
            x /= 0.3; ref /= 0.3; check("/= float")
        x.fill(7); ref.fill(7); check("fill")
        # operands the device does not take go through the host path and still agree
        x += np.ones(shape, dt); ref += np.ones(shape, dt); check("+= array (host path)")
        if dt.kind == "f":
            np.sqrt(x, out=x); np.sqrt(ref, out=ref); check("sqrt (host path)")


def test_inplace_negative_result_raises_and_keeps_tree(oracle_mod):
    from starr_b200 import SumTreeArray
    x = SumTreeArray(np.arange(8, dtype=np.float32))
    before = x.sumtree()
    with pytest.raises(ValueError, match="array elements must be non-negative"):
        x -= 100                                   # leaves change, the tree does not (pyx:58-59)
    assert np.array_equal(np.asarray(x[...]), np.arange(8, dtype=np.float32) - 100)
    assert_bits_equal(x.sumtree(), before, "tree untouched")
    with pytest.raises(Exception):
        y = SumTreeArray(np.arange(8, dtype=np.int32)); y /= 2      # NumPy: no same-kind cast to int32


def test_device_tree_inplace_tensor_operand(torch_mod, oracle_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    C = oracle_mod.c
    rng = np.random.default_rng(2)
    for n in (4097, 1 << 20):
        for dt, tdt in ((np.float32, torch.float32), (np.int32, torch.int32)):
            a = _mk(rng, n, dt); b = _mk(rng, n, dt) + dt(1)
            dev = DeviceSumTree(torch.from_numpy(a).cuda())
            for op, f in (("add", np.add), ("multiply", np.multiply), ("subtract", None)):
                if f is None:
                    dev.inplace_("add", torch.from_numpy(b).cuda()); a = a + b
                    dev.inplace_("subtract", torch.from_numpy(b).cuda()); a = a - b
                else:
                    dev.inplace_(op, torch.from_numpy(b).cuda()); a = f(a, b)
                tree = np.zeros(n, dt)
                C.build_sumtree_from_array(a, tree)
                assert_bits_equal(_np(dev.leaves), a, op); assert_bits_equal(_np(dev.tree), tree, op)
            if dt == np.float32:
                dev.inplace_("true_divide", torch.from_numpy(b).cuda()); a = a / b
                assert_bits_equal(_np(dev.leaves), a, "divide")


# ---- f3: MinTree + ring append ----------------------------------------------------------------------
MIN_SIZES = (2, 3, 5, 6, 7, 13, 64, 100, 1000)


@pytest.mark.parametrize("n", MIN_SIZES)
def test_min_tree_matches_reference_fixtures(torch_mod, golden_next, n):
    """DeviceMinTree / append_ vs fixtures from the reference's slow.MinTree, CircularMinTree and
    CircularSumTree (starr/experimental/slow.py:53-98; tests/test_sumtree_list.py:80-117)."""
    torch = torch_mod
    from starr_b200 import DeviceMinTree, DeviceSumTree
    g = golden_next
    mt = DeviceMinTree(n, torch.float64)
    assert float(mt.min()) == np.inf
    for b in range(3):
        mt.update_(torch.from_numpy(g[f"min_n{n}/b{b}/idx"]).cuda(), torch.from_numpy(g[f"min_n{n}/b{b}/vals"]).cuda())
        want = g[f"min_n{n}/b{b}/data"]
        assert_bits_equal(_np(mt.heap)[1:], want[1:], f"min tree after batch {b}")
        assert float(mt.min()) == want[1]
    ring_min, ring_sum = DeviceMinTree(n, torch.float64), DeviceSumTree(n, torch.float64)
    for b in range(3):
        v = torch.from_numpy(g[f"ring_n{n}/b{b}/vals"]).cuda()
        ring_min.append_(v); ring_sum.append_(v)
        assert_bits_equal(_np(ring_min.heap)[1:], g[f"ring_n{n}/b{b}/min_data"][1:], f"ring min after batch {b}")
        assert_bits_equal(_np(ring_sum.heap), g[f"ring_n{n}/b{b}/sum_data"], f"ring sum after batch {b}")
    mt.poll()


@pytest.mark.parametrize("dt_name", ["float32", "float64", "int32"])
def test_min_tree_large_batches_and_errors(torch_mod, oracle_mod, dt_name):
    torch = torch_mod
    from starr_b200 import DeviceMinTree
    dt = np.dtype(dt_name)
    rng = np.random.default_rng(1)
    for n in (100003, 1 << 20):
        mt = DeviceMinTree(n, getattr(torch, dt_name))
        data = np.full(2 * n, np.inf)
        for m in (1 << 17, 5000, 70000):
            idx = rng.integers(0, n, m).astype(np.int64)
            if m == 5000:
                idx[:] = rng.integers(0, 50, m)                     # heavy duplicates: the last write wins
            vals = _mk(rng, m, dt)
            mt.update_(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda())
            oracle_mod.min_tree_set(data, idx, vals.astype(np.float64))
        got = _np(mt.heap).astype(np.float64)
        if dt.kind != "f":
            data_cmp = np.where(np.isinf(data), float(np.iinfo(dt).max), data)
        else:
            data_cmp = data
        assert np.array_equal(got[1:], data_cmp[1:])
        before = mt.heap.clone()
        bad = torch.from_numpy(np.array([1, n, 2], dtype=np.int64)).cuda()
        mt.update_(bad, torch.ones(3, dtype=mt.dtype, device="cuda"))
        with pytest.raises(IndexError):
            mt.poll()
        assert torch.equal(mt.heap, before)
        mt.build_()
        assert torch.equal(mt.heap[1:], before[1:])


# ---- f4: the experimental shim ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", (2, 3, 5, 37, 1000, 4096))
def test_experimental_shim_matches_reference_fixtures(golden_next, n):
    from starr_b200 import experimental as X
    g = golden_next
    a, t = g[f"exp_n{n}/array0"].copy(), g[f"exp_n{n}/tree0"].copy()
    for b in range(3):
        X.update_tree_multi(g[f"exp_n{n}/b{b}/idx"], g[f"exp_n{n}/b{b}/vals"], a, t)
        assert_bits_equal(a, g[f"exp_n{n}/b{b}/array"], f"array after batch {b}")
        assert_bits_equal(t, g[f"exp_n{n}/b{b}/tree"], f"tree after batch {b}")
    out = np.ones(200, np.int32)
    assert X.get_prefix_sum_idx_multi(out, g[f"exp_n{n}/queries"], a, t) is out
    assert np.array_equal(out, g[f"exp_n{n}/found"])


def test_experimental_shim_large_and_errors(torch_mod, oracle_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    from starr_b200 import experimental as X
    rng = np.random.default_rng(4)
    n, m = 1 << 20, 1 << 18
    a = np.exp(rng.uniform(-8, 8, n)); t = np.zeros(n)
    oracle_mod.c.build_sumtree_from_array(a, t)
    dev = DeviceSumTree(torch.from_numpy(a).cuda())
    idx = rng.integers(0, n, m).astype(np.int32); v = np.exp(rng.uniform(-8, 8, m))
    oracle_mod.exp.update_tree_multi(idx, v, a, t)
    dev.update_(torch.from_numpy(idx.astype(np.int64)).cuda(), torch.from_numpy(v).cuda(), assign_leaf=True)
    assert_bits_equal(_np(dev.leaves), a, "assigned leaves"); assert_bits_equal(_np(dev.tree), t, "tree")
    with pytest.raises(ValueError, match="non-negative"):
        X.update_tree_multi(np.array([0], np.int32), np.array([-1.0]), np.zeros(4), np.zeros(4))
    with pytest.raises(ValueError):
        X.update_tree_multi(np.array([0], np.int64), np.array([1.0]), np.zeros(4), np.zeros(4))   # int32 ids only


def test_zero_delta_counter(torch_mod):
    torch = torch_mod
    from starr_b200 import DeviceSumTree
    n, m = 1 << 20, 1 << 16
    dev = DeviceSumTree(torch.rand(n, device="cuda"))
    idx = torch.randperm(n, device="cuda")[:m]
    val = torch.rand(m, device="cuda")
    dev.handle.update_counters(reset=True)
    dev.update_(idx, val)
    c = dev.handle.update_counters(torch.cuda.current_stream().cuda_stream, reset=True)
    assert c["updates"] == m and c["zero_delta"] < m // 100
    dev.update_(idx, val)                      # the same batch again: nothing changes
    c = dev.handle.update_counters(torch.cuda.current_stream().cuda_stream, reset=True)
    assert c["updates"] == m and c["zero_delta"] > m * 0.7   # fl(leaf + fl(val - leaf)) is not always val
