"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/starr_b200.h declares, refuses to work without a GPU (no CPU fallback), and the
pure host-side logic of the Python mirror behaves like the reference's."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "starr_b200.h")).read()
    return sorted(set(re.findall(r"ST_API\s+[\w\s\*]+?\b(st_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from starr_b200 import _build
    return _build.build()


def test_header_symbols_are_exported(built_lib):
    declared = _declared_symbols()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(st_\w+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/starr_b200.h but not exported: {missing}"


def test_ctypes_binding_covers_header(built_lib):
    from starr_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.lib()
    assert lib.st_abi_version() == 4
    assert [lib.st_dtype_size(c) for c in range(9)] == [2, 4, 8, 1, 2, 4, 8, 4, 8]
    assert lib.st_status_string(1) == b"vals must be non-negative"
    assert lib.st_status_string(2) == b"array elements must be non-negative"


def test_library_is_sm100a_with_tma(built_lib):
    """The shipped cubin targets sm_100a and the search kernel really stages through TMA."""
    r = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass   # cp.async.bulk + mbarrier expect_tx


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import starr_b200
    with pytest.raises(RuntimeError, match="no CPU path"):
        starr_b200.SumTreeArray(np.arange(4.0))
    with pytest.raises(RuntimeError):
        starr_b200.update_sumtree(np.arange(2, dtype=np.intp), np.ones(2), np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError):
        starr_b200.DeviceSumTree(16)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "starr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f


def test_argument_validation_before_any_device_work():
    import starr_b200 as sb
    f8 = np.zeros(10)
    with pytest.raises(TypeError):  # size mismatch, reference pyx:28-29
        sb.update_sumtree(np.arange(4, dtype=np.intp), np.arange(4.0), np.zeros(9), f8)
    with pytest.raises(TypeError):
        sb.build_sumtree_from_array(np.zeros(9), f8)
    with pytest.raises(TypeError):  # pyx:72-73
        sb.get_prefix_sum_idx(np.zeros(11, np.intp), np.arange(4.0), f8, f8)
    with pytest.raises(TypeError):  # fused-type dispatch: complex unsupported
        c = np.zeros(10, np.complex64)
        sb.update_sumtree(np.arange(4, dtype=np.intp), np.arange(4).astype(np.complex64), c, c)
    with pytest.raises(TypeError):  # mixed value dtypes
        sb.update_sumtree(np.arange(4, dtype=np.intp), np.arange(4.0).astype(np.float32), f8, f8)


def test_parse_axes_like_reference():
    from starr_b200._hostlogic import parse_axes
    assert parse_axes(3, None) is None
    for axis, want in ((0, [0]), ((0, 1), [0, 1]), ((1, -1), [1, 2]), (-1, [2]), (-2, [1])):
        got = parse_axes(3, axis)
        assert isinstance(got, np.ndarray) and got.ndim == 1 and list(got) == want
    with pytest.raises(IndexError):
        parse_axes(3, (1, 1))
    with pytest.raises(IndexError):
        parse_axes(3, 3)


def test_plan_sum():
    from starr_b200._hostlogic import fold_output_shape, plan_sum
    assert plan_sum((4096, 4096), [1]) == ("strided", 4096)
    assert plan_sum((4096, 4096), [0]) == ("fold", 1, 4096, 4096)
    assert plan_sum((2, 3, 4, 5), [3]) == ("strided", 5)
    assert plan_sum((2, 3, 4, 5), [2, 3]) == ("strided", 20)
    assert plan_sum((2, 3, 4, 5), [3, 2]) == ("strided", 20)
    assert plan_sum((2, 3, 4, 5), [0, 1, 2, 3]) == ("strided", 120)
    assert plan_sum((2, 3, 4, 5), [1]) == ("fold", 2, 3, 20)
    assert plan_sum((2, 3, 4, 5), [1, 2]) == ("fold", 2, 12, 5)
    assert plan_sum((2, 3, 4, 5), [0, 2]) == ("host",)
    assert plan_sum((3, 1000, 1), [1]) == ("host",)   # C == 1: NumPy sums pairwise there
    assert fold_output_shape((2, 3, 4, 5), [1, 2], False) == (2, 5)
    assert fold_output_shape((2, 3, 4, 5), [1, 2], True) == (2, 1, 1, 5)


def test_shard_geometry():
    from starr_b200._hostlogic import shard_geometry
    assert shard_geometry(1 << 30, 8) == (1 << 27, 3)
    assert shard_geometry(1 << 20, 1) == (1 << 20, 0)
    for bad in ((1000, 2), (1 << 20, 3), (4, 4)):
        with pytest.raises(ValueError):
            shard_geometry(*bad)


def test_slices_exchange_layout_is_aligned_and_disjoint():
    """Host logic of the slices layout (starr_b200/_slices.py): every region of the exchange buffer is 256-byte
    aligned, the two parities of all regions are pairwise disjoint, and the per-source receive regions hold the
    worst case (every element of every rank going to one owner)."""
    import types

    import torch

    from starr_b200._slices import SlicesExchange
    for world, dtype, rb in ((2, torch.float32, 32), (8, torch.float32, 32), (4, torch.float64, 56), (4, torch.int32, 0)):
        tree = types.SimpleNamespace(world=world, dtype=dtype, top=types.SimpleNamespace(top_fold_record_bytes=lambda rb=rb: rb))
        sx = SlicesExchange.__new__(SlicesExchange)
        sx.t, sx.sample_matrix = tree, False
        cap = 5000
        lay = sx._layout(cap)
        es = torch.empty(0, dtype=dtype).element_size()
        spans = []
        names = [k for k in lay if k != "bytes"]
        for i, name in enumerate(names):
            a, b = lay[name]
            assert a % 256 == 0 and b % 256 == 0 and a < b
            nxt = lay[names[i + 1]][0] if i + 1 < len(names) else lay["bytes"]
            size = b - a
            assert b + size == nxt, name                   # parity 1 ends where the next region starts
            spans += [(a, a + size), (b, b + size)]
        spans.sort()
        assert all(x[1] <= y[0] for x, y in zip(spans, spans[1:]))
        size_of = lambda name: lay[name][1] - lay[name][0]   # noqa: E731
        assert size_of("recv_idx") >= world * cap * 8 and size_of("recv_val") >= world * cap * es
        assert size_of("rq_resid") >= world * cap * es and size_of("rq_pos") >= world * cap * 4 and size_of("back") >= world * cap * 8
        assert size_of("mat_u") >= world * (world + 1) * 4
        span = (cap + 1023) // 1024 * 1024
        assert size_of("dslice") >= span * es and size_of("shard") >= span
        if rb:
            assert size_of("recs") >= world * (span // 1024) * (world - 1) * rb
