"""GPU tests of the SumTreeArray drop-in: the behaviours reference tests/test_sumtree_array.py
pins (constructor forms, aliasing across views, guarded writes, set/put forms, ufuncs, sums,
sampling output shapes, error types), plus replay of the API script against fixtures produced
by the unmodified reference (tests/golden/api_cases.npz)."""
import numpy as np
import pytest
from cases import API_SPECS, api_script
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def STA():
    from starr_b200 import SumTreeArray
    return SumTreeArray


def test_creation_forms(STA):
    x = STA(2)
    assert isinstance(x, STA) and x.size == 2 and x.min() == 0 and x.max() == 0
    assert x.dtype == float and not x.flags["WRITEABLE"]
    assert STA(2, dtype=int).dtype == int
    src = np.array([0, 1]).astype("int32")
    x = STA(src)
    src[0] = 99
    assert x[0] == 0 and x[1] == 1 and x._sumtree[1] == 1 and x.dtype == np.int32
    x = STA(np.array([0, 1]).astype("int32"), dtype="int64")
    assert x.dtype == np.int64 and x._sumtree[1] == 1
    a = STA(np.array([0, 1]).astype("int32"))
    b = STA(a)
    assert a is b
    a[0] = 99
    assert b[0] == 99 and b._sumtree[1] == 100
    c = STA(a, dtype="int64")
    a[0] = 5
    assert c[0] == 99 and c.dtype == np.int64 and c._sumtree[1] == 100
    for bad in (0, 1, np.array(1), np.array([1])):
        with pytest.raises(ValueError):
            STA(bad)


def test_view_creation_and_guard(STA):
    x = np.array([1, 2, 3])
    y = x.view(STA)
    assert isinstance(y, STA) and np.shares_memory(x, y)
    assert x.flags["WRITEABLE"] and not y.flags["WRITEABLE"] and y._sumtree[1] == 6
    s = STA(np.array([0, 1]).astype("int32"))
    v = s.view(np.ndarray)
    assert not isinstance(v, STA) and np.shares_memory(s, v) and not v.flags["WRITEABLE"]
    o = s.astype("int32")
    assert not isinstance(o, STA) and not np.shares_memory(s, o)
    s._enable_writes(True)
    assert s.flags["WRITEABLE"]
    s._enable_writes(False)
    assert not s.flags["WRITEABLE"] and s.base.flags["WRITEABLE"]


def test_fill_reshape_ufunc(STA):
    s = STA(np.array([0, 1]).astype("int32"))
    s.fill(99)
    assert s[0] == 99 and s[1] == 99 and s._sumtree[1] == 198
    x1 = STA(np.array([0, 1, 2, 3]).astype("int32"))
    x2 = x1.reshape((2, 2))
    assert x1 is not x2 and x1._flat_base is x2._flat_base and x1._sumtree is x2._sumtree
    assert x1._indices.base is x2._indices.base
    x1[0] = 10
    assert x1._sumtree[1] == 16 and x2[0, 0] == 10 and x2._sumtree[1] == 16
    x2[1] = 20
    assert list(x1) == [10, 1, 20, 20] and x1._sumtree[1] == 51 and x2[1, 1] == 20
    x = STA(np.array([0, 1, 2, 3]).astype("int32"))
    y = x + 1
    assert not isinstance(y, STA) and y[0] == 1 and x[0] == 0 and x._sumtree[1] == 6
    z = x
    z += 10
    assert isinstance(z, STA) and x[0] == 10 and x._sumtree[1] == 46
    with pytest.raises(NotImplementedError):
        np.add(x, 1, out=np.zeros(4, np.int32).view(STA))


def test_set_and_put_forms(STA):
    for use_put in (False, True):
        x = STA(np.array([0, 1, 2, 3, 4, 5]).astype("int32").reshape((3, 2)))
        if use_put:
            x.put(1, 10)
        else:
            x[1] = 10
        assert x.tolist() == [[0, 1], [10, 10], [4, 5]] and x._sumtree[1] == 30
        sel = (np.array([0, 1]), np.array([1, 0]))
        if use_put:
            x.put(sel, 20)
        else:
            x[sel] = 20
        assert x.tolist() == [[0, 20], [20, 10], [4, 5]] and x._sumtree[1] == 59
        mask = np.array([True, False, True])
        if use_put:
            x.put(mask, 30)
        else:
            x[mask] = 30
        assert x.tolist() == [[30, 30], [20, 10], [30, 30]] and x._sumtree[1] == 150
    x = STA(np.zeros((4, 3), np.float32))
    x[1:3, :] = [[1], [2]]          # values cycle in C order, NOT broadcasting (reference quirk)
    assert x.tolist()[1:3] == [[1, 2, 1], [2, 1, 2]]


def test_integer_array_indices_fast_path(STA):
    """1-D array indexed by an integer ndarray: the flat ids are the indices themselves (no lookup table); same result
    as NumPy's own assignment semantics with the reference's value cycling -- negatives, 2-D index arrays, duplicates
    (last write wins), unsigned dtypes, and NumPy's IndexError for anything out of range (nothing modified)."""
    rng = np.random.default_rng(4)
    base = rng.integers(0, 50, 1000).astype(np.float32)
    x = STA(base.copy())
    want = base.copy()
    for idx in (np.array([3, -1, 7, -1000, 999]), rng.integers(-1000, 1000, (4, 25)), np.array([5, 5, 5]),
                np.array([1, 2, 3], dtype=np.uint8), np.array([], dtype=np.int64), np.array([7], dtype=np.int32)):
        vals = rng.integers(0, 50, max(idx.size, 1)).astype(np.float32)
        x[idx] = vals
        for j, i in enumerate(idx.ravel()):          # the reference's loop (pyx:41-49): in order, values cycle
            want[int(i)] = vals[j % vals.size]
        assert x.tolist() == want.tolist()
        assert float(x.sum()) == float(np.asarray(x._sumtree)[1])
    before = x.tolist()
    for bad in (np.array([0, 1000]), np.array([-1001])):
        with pytest.raises(IndexError):
            x[bad] = 1.0
        assert x.tolist() == before


def test_get_returns_plain_arrays(STA):
    x = STA(np.array([0, 1, 2, 3, 4, 5]).astype("int32").reshape((3, 2)))
    y = x[1]
    assert not isinstance(y, STA) and np.shares_memory(x, y) and not y.flags["WRITEABLE"] and list(y) == [2, 3]
    y = x[np.array([0, 1]), np.array([1, 0])]
    assert not isinstance(y, STA) and not np.shares_memory(x, y) and list(y) == [1, 2]
    y = x[np.array([True, False, True])]
    assert y.tolist() == [[0, 1], [4, 5]]


def test_prefix_sum_id_shapes(STA):
    x = STA(np.array([1, 2, 3, 4]).astype("int32"))
    y = x.get_prefix_sum_id(1.5)
    assert isinstance(y, np.ndarray) and y.size == 1 and y[0] == 1
    for flat in (False, True):
        y = x.get_prefix_sum_id([[0.5, 1.5], [3.5, 6.5]], flatten_indices=flat)
        assert y.tolist() == [[0, 1], [2, 3]]
    x = STA(np.array([[1, 2], [3, 4]]).astype("int32"))
    y = x.get_prefix_sum_id(1.5)
    assert isinstance(y, tuple) and len(y) == 2 and y[0][0] == 0 and y[1][0] == 1
    assert x.get_prefix_sum_id(1.5, flatten_indices=True)[0] == 1
    y = x.get_prefix_sum_id([0.5, 1.5, 3.5, 6.5])
    assert list(y[0]) == [0, 0, 1, 1] and list(y[1]) == [0, 1, 0, 1]
    y = x.get_prefix_sum_id([[0.5, 1.5], [3.5, 6.5]])
    assert y[0].tolist() == [[0, 0], [1, 1]] and y[1].tolist() == [[0, 1], [0, 1]]
    assert STA(np.array([1, 2, 3, 4], dtype="float32")).get_prefix_sum_id(np.arange(10)).tolist() == \
        [0, 1, 1, 2, 2, 2, 3, 3, 3, 3]                                  # sumtree_array.py:332-333


def test_sample_shapes_and_validation(STA):
    x = STA(np.array([1, 2, 3, 4]).astype("int32"))
    for flat in (False, True):
        idx = x.sample(flatten_indices=flat)
        assert isinstance(idx, np.ndarray) and idx.size == 1 and x[idx].size == 1
        assert x[x.sample(20, flatten_indices=flat)].size == 20
    x = STA(np.array([[1, 2], [3, 4]]).astype("int32"))
    idx = x.sample(20)
    assert isinstance(idx, tuple) and len(idx) == 2 and x[idx].size == 20
    idx = x.sample(20, flatten_indices=True)
    assert isinstance(idx, np.ndarray) and idx.ndim == 1 and len(idx) == 20
    with pytest.raises(ValueError):
        STA(10).sample(10)
    # distribution sanity: frequencies follow the priorities
    x = STA(np.array([1, 2, 3, 4], dtype="float32"))
    freq = np.bincount(x.sample(40000), minlength=4) / 40000
    assert np.abs(freq - np.array([.1, .2, .3, .4])).max() < 0.02


def test_sums(STA):
    x = STA(np.array([[1, 2], [3, 4]]).astype("int32"))
    assert x.sum() == 10
    y = x.sum(keepdims=True)
    assert y.ravel()[0] == 10 and y.size == 1 and y.ndim == 2
    assert list(x.sum(axis=1)) == [3, 7]
    v = np.random.choice(100, size=(100, 100))
    x = STA(v)
    for kd in (False, True):
        for ax in (0, 1):
            got = x.sum(axis=ax, keepdims=kd)
            assert isinstance(got, np.ndarray) and not isinstance(got, STA)
            assert np.array_equal(got, v.sum(axis=ax, keepdims=kd))
    assert x.sum() == v.sum()
    x1 = np.array([[0, 1, 2], [3, 4, 5]])
    x2 = STA(x1)
    assert x1.mean() == x2.mean()
    assert np.array_equal(x1.mean(axis=0), x2.mean(axis=0)) and np.array_equal(x1.mean(axis=1), x2.mean(axis=1))
    x = STA((2, 3, 4))
    assert x._parse_axis_arg(None) is None and list(x._parse_axis_arg((1, -1))) == [1, 2]
    with pytest.raises(IndexError):
        x._parse_axis_arg((1, 1))
    with pytest.raises(IndexError):
        x._parse_axis_arg(3)
    # the reference's flattened-leading-dims quirk on the tree path
    z = STA(np.ones((2, 3, 4, 5)))
    assert z.sum(axis=3).shape == (24,) and z.sum(axis=3, keepdims=True).shape == (24, 1)
    assert z.sum(axis=(0, 1, 2, 3)).shape == (1,)
    assert z.sum(axis=1).shape == (2, 4, 5) and z.sum(axis=(0, 2)).shape == (3, 5)


def test_errors_and_private_hooks(STA):
    with pytest.raises(ValueError):
        STA(np.array([-1, 0, 1]).astype("int32"))
    x = STA(np.array([1, 2, 3, 4]).astype("int32"))
    with pytest.raises(ValueError):
        x[:2] = -1
    assert list(x) == [1, 2, 3, 4] and x._sumtree[1] == 10
    y = x.sumtree()
    y[0] = 100
    assert not isinstance(y, STA) and x._sumtree[0] == 0
    x = STA(np.arange(4))
    x._sumtree.fill(0)
    assert x._sumtree[1] == 0
    x._rebuild_sumtree()
    assert x._sumtree[1] == 6
    y = x.copy()
    assert isinstance(y, STA) and not np.shares_memory(x, y)
    assert not np.shares_memory(x._flat_base, y._flat_base) and not np.shares_memory(x._sumtree, y._sumtree)
    assert not np.shares_memory(x._indices, y._indices)
    assert y.sum() == x.sum() and np.array_equal(x._sumtree, y._sumtree) and np.array_equal(x._indices, y._indices)
    x = STA(np.array([2., 3., 1.]))
    with pytest.raises(ValueError):
        x -= 100           # leaves already changed, rebuild rejects them (reference behaviour)


def test_passthroughs_return_ndarray(STA):
    a = np.array([[0, 1], [2, 3]])
    x = STA(a)
    for name, args in (("diagonal", ()), ("dot", (a,)), ("flatten", ()), ("repeat", (5,)), ("swapaxes", (0, 1)),
                       ("take", ([0, 3],)), ("trace", ()), ("transpose", ()), ("round", (0,))):
        got = getattr(x, name)(*args)
        assert not isinstance(got, STA)
        assert np.array_equal(got, getattr(a, name)(*args))
    assert not isinstance(x.T, STA) and np.array_equal(x.T, a.T)
    assert not isinstance(x.imag, STA) and np.shares_memory(x.real, x)
    c = STA(np.array([0, 1])).choose([np.array([0, 1]), np.array([2, 3])])
    assert list(c) == [0, 3]
    s = STA(np.array([2, 3, 1, 0]))
    s.sort()
    assert list(s) == [0, 1, 2, 3] and s._sumtree[1] == 6
    p = STA(np.array([2, 3, 1, 0]))
    q = np.array([2, 3, 1, 0])
    p.partition(2); q.partition(2)
    assert np.array_equal(p, q)
    z = STA(np.array([2, 3, 1, 0]).reshape([1, 1, 2, 1, 2, 1, 1])).squeeze()
    assert isinstance(z, STA) and z.shape == (2, 2)


@pytest.mark.parametrize("spec", API_SPECS, ids=[s[0] for s in API_SPECS])
def test_api_script_matches_reference_fixtures(STA, golden_api, spec):
    name, dt, shape = spec
    got = api_script(STA, name, dt, shape)
    for k, v in got.items():
        want = golden_api[f"{name}/{k}"]
        assert_bits_equal(v, want, f"{name}/{k}")
