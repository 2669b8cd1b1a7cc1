"""The five kernel entry points of the reference module, same names / argument order / errors.

Reference: ``starr/src/_sumtree_array.pyx`` (``update_sumtree`` :22, ``build_sumtree_from_array``
:51, ``get_prefix_sum_idx`` :66, ``sum_over`` :173, ``strided_sum`` :185), re-exported at
``starr/__init__.py:1-5``.

The reference functions are stateless: they mutate the caller's ``(array, sumtree)`` NumPy
buffers.  These replacements keep that contract by staging the pair into a temporary device
tree, running the CUDA kernel and copying the pair back, so they are drop-ins for the
reference's own tests.  They pay two PCIe copies of the whole tree per call; code that cares
about speed holds a resident tree instead (``SumTreeArray`` / ``DeviceSumTree``).
"""
from __future__ import annotations

import numpy as np

from . import _lib

_SIZE_MSG = "array and sumtree must have the same size"


def _as_1d(a, name: str) -> np.ndarray:
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    return a


def _value_arrays(*arrays):
    """Fused-type dispatch of the reference: all value arrays share one supported dtype."""
    dt = arrays[0].dtype
    _lib.dtype_code(dt)  # TypeError("No matching signature found") for complex, float16, ...
    for a in arrays[1:]:
        if a.dtype != dt:
            raise TypeError("No matching signature found")
    return dt


def _index_array(a, name: str) -> np.ndarray:
    a = _as_1d(a, name)
    if a.dtype != np.intp:
        raise ValueError("Buffer dtype mismatch, expected 'Py_ssize_t' but got %r" % a.dtype.name)
    return a


class _Staged:
    """(array, sumtree) staged on the device for the duration of one call."""

    def __init__(self, array: np.ndarray, sumtree: np.ndarray):
        if array.size < 2:
            raise ValueError("the device tree needs at least 2 leaves")
        self.array, self.sumtree = array, sumtree
        self.h = _lib.TreeHandle(array.size, array.dtype)
        self.h.load_host(np.ascontiguousarray(array), np.ascontiguousarray(sumtree))

    def write_back(self, leaves: bool = True, tree: bool = True) -> None:
        for dst, want, which in ((self.array, leaves, 0), (self.sumtree, tree, 1)):
            if not want:
                continue
            if not dst.flags.writeable:
                raise ValueError("buffer source array is read-only")
            tmp = dst if dst.flags.c_contiguous else np.empty(dst.size, dst.dtype)
            self.h.store_host(tmp if which == 0 else None, tmp if which == 1 else None)
            if tmp is not dst:
                dst[...] = tmp

    def close(self) -> None:
        self.h.close()


def update_sumtree(idxs, vals, array, sumtree) -> None:
    """``array[idxs[i]] = vals[i % len(vals)]`` for i in order, deltas added to every ancestor
    (reference pyx:22-49)."""
    idxs = _index_array(idxs, "idxs")
    vals, array, sumtree = _as_1d(vals, "vals"), _as_1d(array, "array"), _as_1d(sumtree, "sumtree")
    _value_arrays(vals, array, sumtree)
    if len(array) != len(sumtree):
        raise TypeError(_SIZE_MSG)
    st = _Staged(array, sumtree)
    try:
        st.h.update_host(np.ascontiguousarray(idxs), np.ascontiguousarray(vals))
        st.write_back()
    finally:
        st.close()


def build_sumtree_from_array(array, sumtree) -> None:
    """``sumtree[i] = child(2i) + child(2i+1)`` for i = n-1 .. 1 (reference pyx:51-64)."""
    array, sumtree = _as_1d(array, "array"), _as_1d(sumtree, "sumtree")
    _value_arrays(array, sumtree)
    if len(array) != len(sumtree):
        raise TypeError(_SIZE_MSG)
    st = _Staged(array, sumtree)
    try:
        st.h.build()
        st.write_back(leaves=False)
    finally:
        st.close()


def get_prefix_sum_idx(output, vals, array, sumtree):
    """Left-child descent per query; returns ``output`` (reference pyx:66-122)."""
    output = _index_array(output, "output")
    vals, array, sumtree = _as_1d(vals, "vals"), _as_1d(array, "array"), _as_1d(sumtree, "sumtree")
    _value_arrays(vals, array, sumtree)
    if len(output) != len(vals):
        raise TypeError("output and vals must have the same size")
    if len(array) != len(sumtree):
        raise TypeError(_SIZE_MSG)
    st = _Staged(array, sumtree)
    try:
        tmp = output if output.flags.c_contiguous else np.empty(output.size, np.intp)
        st.h.search_host(np.ascontiguousarray(vals), tmp)
        if tmp is not output:
            output[...] = tmp
    finally:
        st.close()
    return output


def sum_over(array, sumtree, index_from, index_to):
    """Sum of ``array[index_from:index_to]`` through the tree (reference pyx:131-174)."""
    array, sumtree = _as_1d(array, "array"), _as_1d(sumtree, "sumtree")
    _value_arrays(array, sumtree)
    st = _Staged(array, sumtree)
    try:
        return st.h.sum_over(index_from, index_to)
    finally:
        st.close()


def strided_sum(array, sumtree, stride):
    """``out[i] = sum_over(stride*i, min(stride*(i+1), n))`` (reference pyx:176-194)."""
    array, sumtree = _as_1d(array, "array"), _as_1d(sumtree, "sumtree")
    _value_arrays(array, sumtree)
    st = _Staged(array, sumtree)
    try:
        return st.h.strided_sum_host(stride)
    finally:
        st.close()
