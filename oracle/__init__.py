"""oracle -- TEST INFRASTRUCTURE ONLY (never imported by ``starr_b200``).

Two CPU checkers for the SumTreeArray hot path:

* ``oracle.c``  -- ctypes front end to ``liboracle.so`` (``sumtree_oracle.c``), the plain-C
  restatement of reference ``starr/src/_sumtree_array.pyx``.  Same five function names and
  argument order as ``starr/__init__.py:1-5`` so parity tests read like the reference's.
* ``oracle.ref()`` -- the UNMODIFIED reference kernel module compiled into ``oracle/_ref/``
  by ``oracle/build_ref.sh`` (travels to the GPU box as a prebuilt ``.so``).
* ``oracle.ref_api()`` -- the reference's Python ``SumTreeArray`` class, importable only in
  the build container (needs ``/root/reference``); used by ``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os
import subprocess
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_SUFFIX = {
    np.dtype(np.int16): "i16",
    np.dtype(np.int32): "i32",
    np.dtype(np.int64): "i64",
    np.dtype(np.uint8): "u8",
    np.dtype(np.uint16): "u16",
    np.dtype(np.uint32): "u32",
    np.dtype(np.uint64): "u64",
    np.dtype(np.float32): "f32",
    np.dtype(np.float64): "f64",
}
# NumPy's default accumulator dtype for ndarray.sum (sumtree_array.py:514 fallback)
_FOLD_ACC = {
    "i16": np.int64, "i32": np.int64, "i64": np.int64,
    "u8": np.uint64, "u16": np.uint64, "u32": np.uint64, "u64": np.uint64,
    "f32": np.float32, "f64": np.float64,
}
_CTYPE = {
    "i16": ctypes.c_int16, "i32": ctypes.c_int32, "i64": ctypes.c_int64,
    "u8": ctypes.c_uint8, "u16": ctypes.c_uint16, "u32": ctypes.c_uint32,
    "u64": ctypes.c_uint64, "f32": ctypes.c_float, "f64": ctypes.c_double,
}


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    src = os.path.join(_HERE, "sumtree_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/starr/src/_sumtree_array.pyx") and (force or not _ref_so() or not _ref_exp_so()):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for suf, ct in _CTYPE.items():
            getattr(_lib, f"orc_sum_over_{suf}").restype = ct
    return _lib


def _suffix(a: np.ndarray) -> str:
    try:
        return _SUFFIX[a.dtype]
    except KeyError:
        raise TypeError("No matching signature found") from None


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _same_dtype(*arrays):
    dt = arrays[0].dtype
    for a in arrays[1:]:
        if a.dtype != dt:
            raise TypeError("No matching signature found")


def _c(a):
    if not (isinstance(a, np.ndarray) and a.flags.c_contiguous):
        raise TypeError("oracle expects C-contiguous ndarrays")
    return a


class c:
    """Namespace mirroring ``starr._cython`` over the plain-C restatement."""

    @staticmethod
    def update_sumtree(idxs, vals, array, sumtree):
        if len(array) != len(sumtree):
            raise TypeError("array and sumtree must have the same size")
        _same_dtype(vals, array, sumtree)
        if idxs.dtype != np.intp:
            raise TypeError("idxs must be intp")
        suf = _suffix(array)
        rc = getattr(_load(), f"orc_update_{suf}")(
            _ptr(_c(idxs)), ctypes.c_int64(len(idxs)), _ptr(_c(vals)), ctypes.c_int64(len(vals)),
            _ptr(_c(array)), _ptr(_c(sumtree)), ctypes.c_int64(len(array)))
        if rc == 1:
            raise ValueError("vals must be non-negative")
        if rc == 2:
            raise IndexError("Out of bounds on buffer access (axis 0)")
        if rc == 3:
            raise ZeroDivisionError("integer division or modulo by zero")

    @staticmethod
    def build_sumtree_from_array(array, sumtree):
        if len(array) != len(sumtree):
            raise TypeError("array and sumtree must have the same size")
        _same_dtype(array, sumtree)
        suf = _suffix(array)
        rc = getattr(_load(), f"orc_build_{suf}")(_ptr(_c(array)), _ptr(_c(sumtree)),
                                                  ctypes.c_int64(len(array)))
        if rc == 1:
            raise ValueError("array elements must be non-negative")

    @staticmethod
    def get_prefix_sum_idx(output, vals, array, sumtree):
        if len(output) != len(vals):
            raise TypeError("output and vals must have the same size")
        if len(array) != len(sumtree):
            raise TypeError("array and sumtree must have the same size")
        _same_dtype(vals, array, sumtree)
        suf = _suffix(array)
        getattr(_load(), f"orc_search_{suf}")(
            _ptr(_c(output)), _ptr(_c(vals)), ctypes.c_int64(len(vals)),
            _ptr(_c(array)), _ptr(_c(sumtree)), ctypes.c_int64(len(array)))
        return output

    @staticmethod
    def sum_over(array, sumtree, index_from, index_to):
        _same_dtype(array, sumtree)
        suf = _suffix(array)
        v = getattr(_load(), f"orc_sum_over_{suf}")(
            _ptr(_c(array)), _ptr(_c(sumtree)), ctypes.c_int64(len(array)),
            ctypes.c_int64(index_from), ctypes.c_int64(index_to))
        return array.dtype.type(v)

    @staticmethod
    def strided_sum(array, sumtree, stride):
        _same_dtype(array, sumtree)
        suf = _suffix(array)
        n = len(array)
        nout = n // stride + (1 if n % stride else 0)
        out = np.empty(nout, dtype=array.dtype)
        getattr(_load(), f"orc_strided_sum_{suf}")(
            _ptr(_c(array)), _ptr(_c(sumtree)), ctypes.c_int64(n), ctypes.c_int64(stride),
            _ptr(out), ctypes.c_int64(nout))
        return out

    @staticmethod
    def axis_fold(x3d):
        """out[a, c] = sequential fold over r of x3d[a, r, c] (NumPy's order for a
        non-trailing-axis ``ndarray.sum``; see sumtree_oracle.c)."""
        x3d = np.ascontiguousarray(x3d)
        suf = _suffix(x3d)
        A, R, C = x3d.shape
        out = np.empty((A, C), dtype=_FOLD_ACC[suf])
        getattr(_load(), f"orc_axis_fold_{suf}")(
            _ptr(x3d), ctypes.c_int64(A), ctypes.c_int64(R), ctypes.c_int64(C), _ptr(out))
        return out


class exp:
    """``starr.experimental._cython`` (the C++ twin) restated in plain C: float64 / int32, assign-leaf."""

    @staticmethod
    def update_tree_multi(idxs, vals, array, sumtree):
        assert idxs.dtype == np.int32 and vals.dtype == array.dtype == sumtree.dtype == np.float64
        rc = _load().orc_exp_update_tree_multi(_ptr(_c(idxs)), ctypes.c_int32(len(idxs)), _ptr(_c(vals)),
                                               ctypes.c_int32(len(vals)), _ptr(_c(array)), ctypes.c_int32(len(array)),
                                               _ptr(_c(sumtree)))
        if rc == 1:
            raise ValueError("val must be non-negative")

    @staticmethod
    def get_prefix_sum_idx_multi(output, vals, array, sumtree):
        assert output.dtype == np.int32 and vals.dtype == array.dtype == sumtree.dtype == np.float64
        _load().orc_exp_get_prefix_sum_idx_multi(_ptr(_c(output)), _ptr(_c(vals)), ctypes.c_int32(len(vals)),
                                                 _ptr(_c(array)), ctypes.c_int32(len(sumtree)), _ptr(_c(sumtree)))
        return output


def min_tree_set(data, idxs, vals):
    """Sequential ``MinTree.__setitem__`` over a batch on ``data[2n]`` (float64), slow.py:62-68."""
    assert data.dtype == np.float64 and vals.dtype == np.float64 and idxs.dtype == np.int64
    _load().orc_min_set_f64(_ptr(_c(data)), ctypes.c_int64(len(data) // 2), _ptr(_c(idxs)), ctypes.c_int64(len(idxs)),
                            _ptr(_c(vals)), ctypes.c_int64(len(vals)))
    return data


# ----------------------------------------------------------------------------------------
# the real reference, where available
# ----------------------------------------------------------------------------------------
def _ref_exp_so():
    hits = glob.glob(os.path.join(_HERE, "_ref", "_xcython*.so"))
    return hits[0] if hits else None


_ref_exp_mod = None


def ref_exp():
    """The unmodified reference experimental module (``starr.experimental._cython``), or None."""
    global _ref_exp_mod
    if _ref_exp_mod is None:
        so = _ref_exp_so()
        if so is None:
            return None
        spec = importlib.util.spec_from_file_location("starr.experimental._cython", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_exp_mod = mod
    return _ref_exp_mod


def _ref_so():
    hits = glob.glob(os.path.join(_HERE, "_ref", "_cython*.so"))
    return hits[0] if hits else None


_ref_mod = None


def ref():
    """The unmodified reference kernel module (``starr._cython``) from oracle/_ref, or None."""
    global _ref_mod
    if _ref_mod is None:
        so = _ref_so()
        if so is None:
            return None
        spec = importlib.util.spec_from_file_location("starr._cython", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod


def ref_api(reference_root: str = "/root/reference"):
    """The reference package (``starr``: SumTreeArray + the five kernels) assembled from the
    read-only source tree plus the compiled module in oracle/_ref.  Build container only."""
    pkg_dir = os.path.join(reference_root, "starr")
    if not os.path.isdir(pkg_dir) or ref() is None:
        return None
    if "starr" in sys.modules and getattr(sys.modules["starr"], "__oracle_ref__", False):
        return sys.modules["starr"]
    pkg = types.ModuleType("starr")
    pkg.__path__ = [pkg_dir]
    pkg.__oracle_ref__ = True
    sys.modules["starr"] = pkg
    sys.modules["starr._cython"] = ref()
    spec = importlib.util.spec_from_file_location("starr.sumtree_array",
                                                  os.path.join(pkg_dir, "sumtree_array.py"))
    sta = importlib.util.module_from_spec(spec)
    sys.modules["starr.sumtree_array"] = sta
    spec.loader.exec_module(sta)
    for name in ("get_prefix_sum_idx", "strided_sum", "sum_over", "build_sumtree_from_array",
                 "update_sumtree"):
        setattr(pkg, name, getattr(ref(), name))
    pkg.SumTreeArray = sta.SumTreeArray
    return pkg
