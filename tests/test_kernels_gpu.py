"""GPU parity tests of the five kernel entry points (through the C ABI) against the oracle.

Bit-exact for every dtype: tree contents and leaves after build / update batches, sampled
indices, range and strided sums.  Mirrors reference tests/test_cython_sumtree_array.py and adds
randomised differential cases (duplicates, cycling vals, non-power-of-two n, fresh zero tree,
log-uniform magnitudes that exercise float rounding order).
"""
import numpy as np
import pytest
from cases import KERNEL_BATCHES, KERNEL_SPECS, KERNEL_STRIDES
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import starr_b200
    return starr_b200


def _state(h):
    leaves = np.empty(h.n, h.dtype)
    tree = np.empty(h.n, h.dtype)
    h.store_host(leaves, tree)
    return leaves, tree


# ---- the reference's own kernel tests, run against the CUDA path ------------------------------
def test_sum_tree_reference_vectors(sb):
    # reference tests/test_cython_sumtree_array.py:11-79
    idx = np.array([0, 1, 2, 3, 4], dtype=np.intp)
    values = np.array([3, 4, 5, 6, 7], dtype=float)
    base = np.zeros_like(values)
    tree = np.zeros_like(values)
    sb.update_sumtree(idx, np.random.randint(1, 10, 5).astype(float), base, tree)
    sb.update_sumtree(idx, values, base, tree)
    assert list(base) == list(values)
    assert tree[0] == 0 and tree[1] == values.sum()
    assert tree[1] == tree[2] + tree[3] and tree[2] == tree[4] + values[0]
    assert tree[3] == values[1] + values[2] and tree[4] == values[3] + values[4]
    q = np.array([0, 5, 6, 12, 13, 15, 16, 19, 20, 24, 25], dtype=float)
    out = np.zeros_like(q).astype(int)
    sb.get_prefix_sum_idx(out, q, base, tree)
    assert list(out) == [3, 3, 4, 4, 0, 0, 1, 1, 2, 2, 2]
    for i in range(5):
        for j in range(i, 6):
            assert sb.sum_over(base, tree, i, j) == values[i:j].sum()
    assert list(sb.strided_sum(base, tree, 1)) == list(values)
    assert list(sb.strided_sum(base, tree, 2)) == [7, 11, 7]
    assert list(sb.strided_sum(base, tree, 3)) == [12, 13]
    assert list(sb.strided_sum(base, tree, 4)) == [18, 7]
    assert list(sb.strided_sum(base, tree, 5)) == [25]


def test_valid_types(sb):
    # reference :81-108 (float128 has no GPU representation and is rejected instead)
    N, K = 100, 10
    idx = np.random.choice(N, size=K).astype(np.intp)
    vals = np.random.choice(10, size=K).astype(float)
    for at in (np.int16, np.int32, np.int64, np.float32, np.float64, np.uint8, np.uint16, np.uint32, np.uint64):
        base = np.zeros(N).astype(at)
        tree = np.zeros(N).astype(at)
        sb.update_sumtree(idx, vals.astype(at), base, tree)
        assert tree[1] == base.sum()
        out = np.zeros(K).astype(np.intp)
        sb.get_prefix_sum_idx(out, np.random.choice(int(vals.sum()), size=K).astype(at), base, tree)
    with pytest.raises(TypeError):
        b = np.zeros(N, np.float128)
        sb.update_sumtree(idx, vals.astype(np.float128), b, b.copy())


def test_build_equals_updates(sb):
    # reference :166-184
    for n in range(2, 16):
        vals = np.arange(1, n + 1).astype(float)
        a1, t1, t2 = np.zeros(n), np.zeros(n), np.zeros(n)
        sb.update_sumtree(np.arange(n).astype(np.intp), vals, a1, t1)
        sb.build_sumtree_from_array(vals, t2)
        assert t1[1] == vals.sum() and t2[1] == vals.sum()
        assert np.abs(t1 - t2).max() == 0


def test_error_contract(sb):
    # reference :137-164, :186-191 -- and nothing may be written when validation fails
    base, tree = np.arange(10.0), np.zeros(10)
    sb.build_sumtree_from_array(base, tree)
    b0, t0 = base.copy(), tree.copy()
    with pytest.raises(ValueError, match="vals must be non-negative"):
        sb.update_sumtree(np.arange(4, dtype=np.intp), -np.arange(4.0), base, tree)
    with pytest.raises(ValueError, match="vals must be non-negative"):   # unused trailing val still counts
        sb.update_sumtree(np.arange(2, dtype=np.intp), np.array([1.0, 2.0, -1.0]), base, tree)
    assert np.array_equal(base, b0) and np.array_equal(tree, t0)
    with pytest.raises(ValueError, match="array elements must be non-negative"):
        sb.build_sumtree_from_array(-np.arange(10.0), tree)
    assert np.array_equal(tree, t0)
    with pytest.raises(IndexError):
        sb.update_sumtree(np.array([3, 10], dtype=np.intp), np.ones(2), base, tree)
    assert np.array_equal(base, b0) and np.array_equal(tree, t0)
    with pytest.raises(ZeroDivisionError):
        sb.update_sumtree(np.arange(4, dtype=np.intp), np.zeros(0), base, tree)
    # a failed call must not poison the next one
    sb.update_sumtree(np.array([1], dtype=np.intp), np.array([5.0]), base, tree)
    assert base[1] == 5 and tree[1] == b0.sum() + 4


# ---- golden fixtures generated from the unmodified reference ----------------------------------
@pytest.mark.parametrize("spec", KERNEL_SPECS, ids=[s[0] for s in KERNEL_SPECS])
def test_golden_fixtures(sb, golden_kernel, spec):
    from starr_b200 import _lib
    g = golden_kernel
    name, dt, n, _ = spec
    h = _lib.TreeHandle(n, dt)
    h.build_from_host(g[f"{name}/leaves0"])
    leaves, tree = _state(h)
    assert_bits_equal(tree, g[f"{name}/tree0"], "build")
    for b in range(KERNEL_BATCHES):
        h.update_host(g[f"{name}/b{b}/idx"], g[f"{name}/b{b}/vals"])
        leaves, tree = _state(h)
        assert_bits_equal(leaves, g[f"{name}/b{b}/leaves"], f"leaves after batch {b}")
        assert_bits_equal(tree, g[f"{name}/b{b}/tree"], f"tree after batch {b}")
    found = np.zeros(512, np.intp)
    h.search_host(g[f"{name}/queries"], found)
    assert np.array_equal(found, g[f"{name}/found"])
    # sample(): (root * u).astype(dtype) evaluated on the device == NumPy's host arithmetic
    u = g[f"{name}/u"]
    q_host = (tree[1] * u).astype(dt)
    want = np.zeros(512, np.intp)
    h.search_host(q_host, want)
    got = np.zeros(512, np.intp)
    h.sample_uniform_host(u, got)
    assert np.array_equal(got, want)
    for s in KERNEL_STRIDES + (n,):
        assert_bits_equal(h.strided_sum_host(s), g[f"{name}/strided{s}"], f"strided {s}")
    sums = np.array([h.sum_over(int(a), int(b)) for a, b in g[f"{name}/ranges"]], dtype=dt)
    assert_bits_equal(sums, g[f"{name}/range_sums"], "sum_over")
    h.close()


# ---- randomised differential tests against the oracle -----------------------------------------
def _draw(rng, k, dt, kind):
    if np.issubdtype(dt, np.floating):
        if kind == "log":
            return np.exp(rng.uniform(-8, 8, k)).astype(dt)
        if kind == "mixed":   # zeros, tiny, huge: absorption and binade crossings
            v = np.exp(rng.uniform(-20, 20, k))
            v[rng.random(k) < 0.2] = 0
            return v.astype(dt)
        return rng.random(k).astype(dt)
    return rng.integers(0, 1000, k).astype(dt)


CASES = [
    (np.float32, 2, "uni"), (np.float32, 3, "log"), (np.float32, 6, "log"), (np.float32, 37, "log"),
    (np.float32, 1000, "uni"), (np.float32, 1000, "mixed"), (np.float32, 1024, "log"),
    (np.float32, 65536, "uni"), (np.float32, 100003, "log"), (np.float32, 1 << 20, "uni"),
    (np.float64, 5, "log"), (np.float64, 1000, "mixed"), (np.float64, 40000, "uni"),
    (np.int32, 7, "uni"), (np.int32, 1000, "uni"), (np.int32, 1 << 18, "uni"),
    (np.int64, 999, "uni"), (np.int16, 300, "uni"), (np.uint8, 77, "uni"), (np.uint16, 513, "uni"),
    (np.uint32, 4097, "uni"), (np.uint64, 1025, "uni"),
]


@pytest.mark.parametrize("dt,n,kind", CASES, ids=[f"{np.dtype(c[0]).name}-{c[1]}-{c[2]}" for c in CASES])
def test_differential_vs_oracle(oracle_mod, dt, n, kind):
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(n * 31 + np.dtype(dt).itemsize)
    leaves = _draw(rng, n, dt, kind)
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    h = _lib.TreeHandle(n, dt)
    h.build_from_host(leaves)
    gl, gt = _state(h)
    assert_bits_equal(gt, tree, "build")
    batch_sizes = [1, 2, 33, min(4096, 4 * n), min(1 << 17, 8 * n)]
    for bi, m in enumerate(batch_sizes):
        nv = m if bi % 2 == 0 else int(rng.integers(1, m + 1))
        if bi == 3:     # skewed: many duplicates
            idx = rng.integers(0, max(1, n // 8), m).astype(np.intp)
        elif bi == 4 and n > 64:   # a contiguous run plus random
            idx = np.concatenate([np.arange(m // 2) % n, rng.integers(0, n, m - m // 2)]).astype(np.intp)
        else:
            idx = rng.integers(0, n, m).astype(np.intp)
        vals = _draw(rng, nv, dt, kind)
        C.update_sumtree(idx, vals, leaves, tree)
        h.update_host(idx, vals)
        gl, gt = _state(h)
        assert_bits_equal(gl, leaves, f"leaves after batch {bi} (m={m}, nv={nv})")
        assert_bits_equal(gt, tree, f"tree after batch {bi} (m={m}, nv={nv})")
    m = 5000
    q = (rng.random(m) * float(tree[1]) * 1.05).astype(dt)
    want = np.zeros(m, np.intp)
    C.get_prefix_sum_idx(want, q, leaves, tree)
    got = np.zeros(m, np.intp)
    h.search_host(q, got)
    assert np.array_equal(got, want)
    for s in (1, 2, 3, 5, 8, 100, 4096, n):
        assert_bits_equal(h.strided_sum_host(s), C.strided_sum(leaves, tree, s), f"strided {s}")
    h.close()


def test_update_from_fresh_zero_tree(oracle_mod):
    """PER start-up: all-zero tree filled by updates (every node starts at 0: no float shortcut
    applies, everything goes through the ordered folds)."""
    from starr_b200 import _lib
    C = oracle_mod.c
    for dt in (np.float32, np.float64):
        n = 50000
        rng = np.random.default_rng(9)
        leaves, tree = np.zeros(n, dt), np.zeros(n, dt)
        h = _lib.TreeHandle(n, dt)
        for m in (100, 20000, 100000):
            idx = rng.integers(0, n, m).astype(np.intp)
            vals = np.exp(rng.uniform(-6, 6, m)).astype(dt)
            C.update_sumtree(idx, vals, leaves, tree)
            h.update_host(idx, vals)
            gl, gt = _state(h)
            assert_bits_equal(gl, leaves, "leaves")
            assert_bits_equal(gt, tree, "tree")
        h.close()


def test_batch_order_matters_and_is_respected(oracle_mod):
    """Same multiset of updates, different order -> different float tree; we must follow each."""
    from starr_b200 import _lib
    C = oracle_mod.c
    n, m = 1 << 14, 8192
    rng = np.random.default_rng(4)
    base = rng.random(n).astype(np.float32)
    idx = rng.integers(0, n, m).astype(np.intp)
    vals = np.exp(rng.uniform(-8, 8, m)).astype(np.float32)
    trees = []
    for perm in (np.arange(m), rng.permutation(m)):
        leaves, tree = base.copy(), np.zeros(n, np.float32)
        C.build_sumtree_from_array(leaves, tree)
        C.update_sumtree(idx[perm], vals[perm], leaves, tree)
        h = _lib.TreeHandle(n, np.float32)
        h.build_from_host(base)
        h.update_host(idx[perm], vals[perm])
        gl, gt = _state(h)
        assert_bits_equal(gt, tree, "tree")
        trees.append(tree)
        h.close()
    assert (trees[0].view(np.uint32) != trees[1].view(np.uint32)).any()


def test_axis_fold_matches_numpy(oracle_mod):
    from starr_b200 import _lib
    rng = np.random.default_rng(8)
    for dt in (np.float32, np.float64, np.int32, np.uint8, np.int64):
        for shape, lo, hi in (((64, 48), 0, 0), ((5, 40, 7), 1, 1), ((3, 4, 5, 6), 0, 1), ((300, 257), 0, 0)):
            x = _draw(rng, int(np.prod(shape)), dt, "log").reshape(shape)
            if np.dtype(dt).itemsize == 1:
                x = (x % 5).astype(dt)
            A = int(np.prod(shape[:lo])); R = int(np.prod(shape[lo:hi + 1])); Cc = int(np.prod(shape[hi + 1:]))
            h = _lib.TreeHandle(x.size, dt)
            h.build_from_host(x.ravel())
            got = h.axis_fold_host(A, R, Cc)
            want = x.sum(axis=tuple(range(lo, hi + 1)))
            assert_bits_equal(got.reshape(want.shape), want, f"{np.dtype(dt).name} {shape}")
            h.close()


def test_axis_fold_tma_strips_match_numpy(oracle_mod):
    """Shapes that take the TMA strip kernel (st_axis_fold.cu): ragged column strips, row counts that are no
    multiple of the tile, leading dims, integer row splits, a leading -0.0 row, every strip width."""
    from starr_b200 import _lib
    rng = np.random.default_rng(9)
    cases = [(np.float32, (1, 1000, 1024)), (np.float32, (3, 200, 260)), (np.float64, (2, 777, 64)),
             (np.int32, (1, 5000, 512)), (np.uint8, (1, 300, 4096)), (np.int16, (2, 1031, 136)),
             (np.float32, (1, 4097, 36)), (np.uint64, (1, 2048, 40)), (np.float32, (1, 65, 6000)),
             (np.float64, (1, 3000, 4098))]
    for dt, (A, R, Cc) in cases:
        x = _draw(rng, A * R * Cc, dt, "log").reshape(A, R, Cc)
        if np.dtype(dt).itemsize == 1:
            x = (x % 5).astype(dt)
        if np.dtype(dt).kind == "f":
            x[:, 0, ::3] = -0.0
            x[:, 1:, 0] = 0.0          # column 0 stays -0.0 only if the fold starts from row 0 and adds +0.0 ... = +0.0
            x[:, 1:, 3] = -0.0         # column 3: (-0.0) + (-0.0) stays -0.0
        h = _lib.TreeHandle(x.size, dt)
        h.build_from_host(x.ravel())
        got = h.axis_fold_host(A, R, Cc)
        want = x.sum(axis=1)
        assert_bits_equal(got.reshape(want.shape), want, f"{np.dtype(dt).name} {(A, R, Cc)}")
        h.close()


# ---- the exact parallel ordered fold (st_ordered_fold.cuh) under adversarial inputs ------------
def _adversarial_batches(rng, n, m, dt, kind):
    idx = rng.integers(0, n, m).astype(np.intp)
    if kind == "ties":            # multiples of a small power of two: exact half-ulp ties everywhere
        vals = (rng.integers(0, 1 << 12, m) * 2.0 ** -9).astype(dt)
    elif kind == "boundary":      # total hovers around a power of two and hops across it
        vals = np.where(rng.random(m) < 0.5, 1.0, 1.0 + 2.0 ** -10).astype(dt)
        idx = rng.integers(0, 4, m).astype(np.intp)
    elif kind == "range":         # 60 orders of magnitude, zeros and subnormals
        vals = (10.0 ** rng.uniform(-44 if dt == np.float32 else -320, 30, m)).astype(dt)
        vals[rng.random(m) < 0.1] = 0
    elif kind == "grow":          # monotone growth from zero: a binade crossing per doubling
        vals = (np.arange(m) % 1000 + rng.random(m)).astype(dt)
    else:                         # "same": one leaf hammered
        vals = rng.random(m).astype(dt)
        idx[:] = n // 3
    return idx, vals


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["ties", "boundary", "range", "grow", "same"])
def test_ordered_fold_adversarial(oracle_mod, dt, kind):
    """Few leaves, many updates: every node sees a long ordered segment, so the result depends
    on the parity-map scan, its restarts at binade crossings and the sequential fallbacks."""
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(hash(kind) % 1000)
    for n, m in ((64, 200_000), (5, 40_000), (4096, 300_000)):
        leaves = np.zeros(n, dt) if kind in ("grow", "range") else np.full(n, (1 << 20) / n, dt)
        tree = np.zeros(n, dt)
        C.build_sumtree_from_array(leaves, tree)
        h = _lib.TreeHandle(n, dt)
        h.build_from_host(leaves)
        for rep in range(2):
            idx, vals = _adversarial_batches(rng, n, m, dt, kind)
            C.update_sumtree(idx, vals, leaves, tree)
            h.update_host(idx, vals)
            gl, gt = _state(h)
            assert_bits_equal(gl, leaves, f"{kind} n={n} rep={rep} leaves")
            assert_bits_equal(gt, tree, f"{kind} n={n} rep={rep} tree")
        h.close()


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_node_sitting_on_a_power_of_two(oracle_mod, dt):
    """Upper nodes EXACTLY on a power of two whose updates are all a fraction of an ulp (uniform priorities on a
    power-of-two tree: bench.py's fixed-2^30 leg had top node 3 == 2^28).  Tiny negative updates put the exact sum just
    below the binade, where the finer grid still rounds to 2^k: the chunk summaries must cover that (no chunk folded
    one add at a time), and the result stays bit-identical."""
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(5)
    n, m = 1 << 12, 200_000
    big = 2.0 ** (30 if dt == np.float32 else 60)
    leaves = np.full(n, big / n, dt)                      # every node is a power of two; ulp(root) = 2^7
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    h = _lib.TreeHandle(n, dt)
    h.build_from_host(leaves)
    before = h.update_stats()
    for rep in range(2):
        idx = rng.integers(0, n, m).astype(np.intp)
        # differences of a few ulps of the LEAF = 2^-12 ulps of the root: absorbed by every node that sees a huge segment
        vals = (leaves[idx] + rng.integers(-2, 3, m).astype(dt) * dt(np.spacing(dt(big / n)))).astype(dt)
        C.update_sumtree(idx, vals, leaves, tree)
        h.update_host(idx, vals)
        gl, gt = _state(h)
        assert_bits_equal(gl, leaves, f"rep={rep} leaves")
        assert_bits_equal(gt, tree, f"rep={rep} tree")
    after = h.update_stats()
    assert after["huge_segments_maps_only"] > before["huge_segments_maps_only"]
    assert after["chunks_folded_exactly"] == before["chunks_folded_exactly"], after
    h.close()


@pytest.mark.parametrize("dt", [np.float32, np.int32], ids=["f32", "i32"])
def test_batches_larger_than_one_chunk(oracle_mod, dt):
    """Batches above 2^21 updates are applied chunk by chunk (2^20 + 1 is now ONE ragged chunk); `vals` keeps cycling over the position
    in the WHOLE batch (pyx:44), duplicates may straddle chunk boundaries."""
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(21)
    n = 1 << 22
    leaves = _draw(rng, n, dt, "uni")
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    h = _lib.TreeHandle(n, dt)
    h.build_from_host(leaves)
    for m, nv in (((1 << 21) + 12345, 1000003), ((1 << 20) + 1, (1 << 20) + 1), (5 << 20, 7)):
        idx = rng.integers(0, n, m).astype(np.intp)
        idx[::5] = idx[0]                      # one leaf hit in every chunk
        vals = _draw(rng, nv, dt, "log")
        C.update_sumtree(idx, vals, leaves, tree)
        h.update_host(idx, vals)
        gl, gt = _state(h)
        assert_bits_equal(gl, leaves, f"leaves m={m} nv={nv}")
        assert_bits_equal(gt, tree, f"tree m={m} nv={nv}")
    h.close()


@pytest.mark.parametrize("dt,n", [(np.float32, 1 << 22), (np.float64, 1 << 22), (np.float32, (1 << 22) + 12345),
                                  (np.float32, 3_000_001)], ids=["f32", "f64", "f32-odd", "f32-3M"])
def test_sparse_float_batches_sort_fewer_bits(oracle_mod, dt, n):
    """Sparse float batches leave the low key bits unsorted (one radix pass less): a group then holds
    the updates of up to 128 neighbouring leaves, interleaved in batch order.  Crowd a few groups with
    duplicates of several neighbouring leaves; everything else stays sparse."""
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(23)
    leaves = _draw(rng, n, dt, "log")
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    h = _lib.TreeHandle(n, dt)
    h.build_from_host(leaves)
    for m in (5000, 20_000, 32_768):
        idx = rng.integers(0, n, m).astype(np.intp)
        for base in rng.integers(0, n - 256, 6):              # crowded neighbourhoods, duplicates interleaved
            where = rng.integers(0, m, 300)
            idx[where] = base + rng.integers(0, 200, 300)
        idx[rng.integers(0, m, 50)] = n - 1                   # and the last leaf, repeatedly
        vals = _draw(rng, m // 3, dt, "log")                  # cycling vals
        C.update_sumtree(idx, vals, leaves, tree)
        h.update_host(idx, vals)
        gl, gt = _state(h)
        assert_bits_equal(gl, leaves, f"leaves n={n} m={m}")
        assert_bits_equal(gt, tree, f"tree n={n} m={m}")
    h.close()


def test_non_power_of_two_large(oracle_mod):
    """A large non-power-of-two tree (leaves at two depths, rotated order) through the big-batch path."""
    from starr_b200 import _lib
    C = oracle_mod.c
    rng = np.random.default_rng(22)
    for n in (3_000_001, (1 << 22) - 1, (1 << 21) + 1):
        leaves = rng.random(n).astype(np.float32)
        tree = np.zeros(n, np.float32)
        C.build_sumtree_from_array(leaves, tree)
        h = _lib.TreeHandle(n, np.float32)
        h.build_from_host(leaves)
        gl, gt = _state(h)
        assert_bits_equal(gt, tree, f"build n={n}")
        m = 400_000
        idx = rng.integers(0, n, m).astype(np.intp)
        vals = np.exp(rng.uniform(-6, 6, m)).astype(np.float32)
        C.update_sumtree(idx, vals, leaves, tree)
        h.update_host(idx, vals)
        gl, gt = _state(h)
        assert_bits_equal(gl, leaves, f"leaves n={n}")
        assert_bits_equal(gt, tree, f"tree n={n}")
        q = (rng.random(100_000) * float(tree[1])).astype(np.float32)
        want = np.zeros(q.size, np.intp)
        C.get_prefix_sum_idx(want, q, leaves, tree)
        got = np.zeros(q.size, np.intp)
        h.search_host(q, got)
        assert np.array_equal(got, want)
        for s in (1000, 4096, 3):
            assert_bits_equal(h.strided_sum_host(s), C.strided_sum(leaves, tree, s), f"strided {s} n={n}")
        h.close()
