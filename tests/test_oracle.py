"""CPU tests: the oracle (plain-C restatement) is pinned against
  * the reference's own golden vectors (tests/test_cython_sumtree_array.py:11-79, README.md:46-47,
    starr/sumtree_array.py:332-333),
  * fixtures generated from the unmodified reference (tests/golden/*.npz),
  * the compiled reference module in oracle/_ref, bit for bit, when it is present.
"""
import numpy as np
import pytest
from cases import KERNEL_BATCHES, KERNEL_SPECS, KERNEL_STRIDES
from conftest import assert_bits_equal


def test_reference_n5_tree_and_search_vector(oracle_mod):
    # reference tests/test_cython_sumtree_array.py:11-79
    C = oracle_mod.c
    idx = np.arange(5, dtype=np.intp)
    values = np.array([3, 4, 5, 6, 7], dtype=float)
    base, tree = np.zeros(5), np.zeros(5)
    C.update_sumtree(idx, np.random.randint(1, 10, 5).astype(float), base, tree)
    C.update_sumtree(idx, values, base, tree)
    assert list(base) == [3, 4, 5, 6, 7]
    assert list(tree) == [0, 25, 16, 9, 13]
    q = np.array([0, 5, 6, 12, 13, 15, 16, 19, 20, 24, 25], dtype=float)
    out = np.zeros(q.size, dtype=np.intp)
    C.get_prefix_sum_idx(out, q, base, tree)
    assert list(out) == [3, 3, 4, 4, 0, 0, 1, 1, 2, 2, 2]
    for i in range(5):
        for j in range(i, 6):
            assert C.sum_over(base, tree, i, j) == values[i:j].sum()
    assert list(C.strided_sum(base, tree, 1)) == [3, 4, 5, 6, 7]
    assert list(C.strided_sum(base, tree, 2)) == [7, 11, 7]
    assert list(C.strided_sum(base, tree, 3)) == [12, 13]
    assert list(C.strided_sum(base, tree, 4)) == [18, 7]
    assert list(C.strided_sum(base, tree, 5)) == [25]


def test_reference_doc_values(oracle_mod):
    C = oracle_mod.c
    a = np.array([1, 2, 3, 4], dtype=np.float32)
    t = np.zeros_like(a)
    C.build_sumtree_from_array(a, t)
    assert list(t) == [0, 10, 3, 7]                       # README.md:46-47
    out = np.zeros(10, dtype=np.intp)
    C.get_prefix_sum_idx(out, np.arange(10, dtype=np.float32), a, t)
    assert list(out) == [0, 1, 1, 2, 2, 2, 3, 3, 3, 3]    # sumtree_array.py:332-333


def test_build_equals_updates_from_zero(oracle_mod):
    # reference tests/test_cython_sumtree_array.py:166-184
    C = oracle_mod.c
    for n in range(2, 16):
        vals = np.arange(1, n + 1).astype(float)
        a1, t1, t2 = np.zeros(n), np.zeros(n), np.zeros(n)
        C.update_sumtree(np.arange(n, dtype=np.intp), vals, a1, t1)
        C.build_sumtree_from_array(vals, t2)
        assert t1[1] == vals.sum() and np.array_equal(t1, t2)


def test_error_contract(oracle_mod):
    C = oracle_mod.c
    with pytest.raises(ValueError):
        C.update_sumtree(np.arange(4, dtype=np.intp), -np.arange(4.0), np.zeros(10), np.zeros(10))
    with pytest.raises(ValueError):
        C.build_sumtree_from_array(-np.arange(10.0), np.zeros(10))
    with pytest.raises(TypeError):
        C.update_sumtree(np.arange(4, dtype=np.intp), np.arange(4.0), np.zeros(9), np.zeros(10))
    with pytest.raises(TypeError):
        C.get_prefix_sum_idx(np.zeros(11, np.intp), np.arange(4.0), np.zeros(10), np.zeros(10))
    with pytest.raises(TypeError):
        C.update_sumtree(np.arange(4, dtype=np.intp), np.arange(4).astype(np.complex64),
                         np.zeros(10, np.complex64), np.zeros(10, np.complex64))


@pytest.mark.parametrize("spec", KERNEL_SPECS, ids=[s[0] for s in KERNEL_SPECS])
def test_oracle_matches_golden_fixtures(oracle_mod, golden_kernel, spec):
    C, g = oracle_mod.c, golden_kernel
    name, dt, n, _ = spec
    leaves = g[f"{name}/leaves0"].copy()
    tree = np.zeros(n, dt)
    C.build_sumtree_from_array(leaves, tree)
    assert_bits_equal(tree, g[f"{name}/tree0"], "build")
    for b in range(KERNEL_BATCHES):
        C.update_sumtree(g[f"{name}/b{b}/idx"], g[f"{name}/b{b}/vals"], leaves, tree)
        assert_bits_equal(leaves, g[f"{name}/b{b}/leaves"], f"leaves after batch {b}")
        assert_bits_equal(tree, g[f"{name}/b{b}/tree"], f"tree after batch {b}")
    found = np.zeros(512, np.intp)
    C.get_prefix_sum_idx(found, g[f"{name}/queries"], leaves, tree)
    assert np.array_equal(found, g[f"{name}/found"])
    for s in KERNEL_STRIDES + (n,):
        assert_bits_equal(C.strided_sum(leaves, tree, s), g[f"{name}/strided{s}"], f"strided {s}")
    sums = np.array([C.sum_over(leaves, tree, int(a), int(b)) for a, b in g[f"{name}/ranges"]], dtype=dt)
    assert_bits_equal(sums, g[f"{name}/range_sums"], "sum_over")


def test_oracle_matches_compiled_reference(oracle_mod):
    R = oracle_mod.ref()
    if R is None:
        pytest.skip("oracle/_ref not built (needs the reference tree: oracle/build_ref.sh)")
    C = oracle_mod.c
    rng = np.random.default_rng(5)
    for dt in (np.float32, np.float64, np.int32, np.int16, np.uint8, np.uint64):
        for n in (2, 3, 6, 37, 1000, 1024):
            f = np.issubdtype(dt, np.floating)
            mk = (lambda k: np.exp(rng.uniform(-8, 8, k)).astype(dt)) if f else (lambda k: rng.integers(0, 50, k).astype(dt))
            a1 = mk(n); a2 = a1.copy()
            t1 = np.zeros(n, dt); t2 = np.zeros(n, dt)
            R.build_sumtree_from_array(a1, t1); C.build_sumtree_from_array(a2, t2)
            for _ in range(3):
                m = int(rng.integers(1, 4 * n)); nv = int(rng.integers(1, m + 1))
                idx = rng.integers(0, n, m).astype(np.intp); v = mk(nv)
                R.update_sumtree(idx, v, a1, t1); C.update_sumtree(idx, v, a2, t2)
                assert_bits_equal(a2, a1, "leaves"); assert_bits_equal(t2, t1, "tree")
            q = (rng.random(300) * float(t1[1]) * 1.1).astype(dt)
            o1 = np.zeros(300, np.intp); o2 = np.zeros(300, np.intp)
            R.get_prefix_sum_idx(o1, q, a1, t1); C.get_prefix_sum_idx(o2, q, a2, t2)
            assert np.array_equal(o1, o2)
            for s in (1, 2, 3, 5, 8, n):
                assert_bits_equal(C.strided_sum(a2, t2, s), R.strided_sum(a1, t1, s), f"strided {s}")


def test_axis_fold_is_numpy_order(oracle_mod):
    # the reference's non-contiguous sum is ndarray.sum (sumtree_array.py:514): sequential rows
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64, np.int32, np.uint8):
        for shape, lo, hi in (((64, 48), 0, 0), ((5, 40, 7), 1, 1), ((3, 4, 5, 6), 0, 1), ((3, 4, 5, 6), 1, 2)):
            x = (np.exp(rng.uniform(-8, 8, shape)) if np.issubdtype(dt, np.floating) else rng.integers(0, 90, shape)).astype(dt)
            if np.issubdtype(dt, np.floating):
                x[..., 0] = -0.0      # add.reduce starts from +0: an all -0.0 column sums to +0.0, not -0.0
            A = int(np.prod(shape[:lo])); R = int(np.prod(shape[lo:hi + 1])); Cc = int(np.prod(shape[hi + 1:]))
            want = x.sum(axis=tuple(range(lo, hi + 1)))
            got = oracle_mod.c.axis_fold(x.reshape(A, R, Cc)).reshape(want.shape)
            assert_bits_equal(got, want, f"{np.dtype(dt).name} {shape} axes {lo}..{hi}")


# ---- experimental twin and MinTree restatements ---------------------------------
EXP_SIZES = (2, 3, 5, 37, 1000, 4096)
MIN_SIZES = (2, 3, 5, 6, 7, 13, 64, 100, 1000)


@pytest.mark.parametrize("n", EXP_SIZES)
def test_experimental_oracle_matches_golden(oracle_mod, golden_next, n):
    """orc_exp_* vs fixtures generated from the unmodified starr.experimental._cython."""
    g, X = golden_next, oracle_mod.exp
    a, t = g[f"exp_n{n}/array0"].copy(), g[f"exp_n{n}/tree0"].copy()
    for b in range(3):
        X.update_tree_multi(g[f"exp_n{n}/b{b}/idx"], g[f"exp_n{n}/b{b}/vals"], a, t)
        assert_bits_equal(a, g[f"exp_n{n}/b{b}/array"], f"array after batch {b}")
        assert_bits_equal(t, g[f"exp_n{n}/b{b}/tree"], f"tree after batch {b}")
    out = np.ones(200, np.int32)
    X.get_prefix_sum_idx_multi(out, g[f"exp_n{n}/queries"], a, t)
    assert np.array_equal(out, g[f"exp_n{n}/found"])


def test_experimental_oracle_matches_compiled_reference(oracle_mod):
    R = oracle_mod.ref_exp()
    if R is None:
        pytest.skip("oracle/_ref/_xcython not built (needs the reference tree)")
    rng = np.random.default_rng(11)
    for n in (2, 3, 6, 100, 1023):
        a1 = rng.random(n); t1 = np.zeros(n)
        oracle_mod.c.build_sumtree_from_array(a1, t1)
        a2, t2 = a1.copy(), t1.copy()
        for _ in range(3):
            m = int(rng.integers(1, 4 * n))
            idx = rng.integers(0, n, m).astype(np.int32); v = np.exp(rng.uniform(-8, 8, int(rng.integers(1, m + 1))))
            R.update_tree_multi(idx, v, a1, t1); oracle_mod.exp.update_tree_multi(idx, v, a2, t2)
            assert_bits_equal(a2, a1, "array"); assert_bits_equal(t2, t1, "tree")


def test_experimental_assign_differs_from_main_path(oracle_mod):
    """The twin ASSIGNS the leaf (cpp:23); the main module stores fl(old + fl(val - old)) (pyx:43-45)."""
    a1 = np.array([1e8, 1.0]); t1 = np.zeros(2); oracle_mod.c.build_sumtree_from_array(a1, t1)
    a2, t2 = a1.copy(), t1.copy()
    v = np.array([0.1])
    oracle_mod.c.update_sumtree(np.array([0], np.intp), v, a1, t1)
    oracle_mod.exp.update_tree_multi(np.array([0], np.int32), v, a2, t2)
    assert a2[0] == 0.1 and a1[0] != 0.1 and t1[1] == t2[1]


@pytest.mark.parametrize("n", MIN_SIZES)
def test_min_oracle_matches_golden(oracle_mod, golden_next, n):
    """orc_min_set_f64 vs fixtures from the reference's slow.MinTree / CircularMinTree."""
    g = golden_next
    data = np.full(2 * n, np.inf)
    for b in range(3):
        oracle_mod.min_tree_set(data, g[f"min_n{n}/b{b}/idx"], g[f"min_n{n}/b{b}/vals"])
        assert_bits_equal(data, g[f"min_n{n}/b{b}/data"], f"min tree after batch {b}")
    ring, cursor = np.full(2 * n, np.inf), 0
    for b in range(3):
        v = g[f"ring_n{n}/b{b}/vals"]
        oracle_mod.min_tree_set(ring, (np.arange(v.size, dtype=np.int64) + cursor) % n, v)
        cursor = (cursor + v.size) % n
        assert_bits_equal(ring, g[f"ring_n{n}/b{b}/min_data"], f"ring min tree after batch {b}")
