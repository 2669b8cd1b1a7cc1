"""``starr.experimental`` shim: the two functions of the reference's experimental
C++ twin, same names / argument order / dtypes, on the CUDA kernels.

Reference: ``starr/experimental/src/_sumtree_array.pyx:22-45`` (the Cython wrappers) over
``starr/experimental/src/sumtree_array.cpp:8-85``.  ``float64`` values and ``int32`` indices only.
Two differences from ``starr.update_sumtree`` that are kept:

* the leaf is ASSIGNED ``array[idx] = val`` (``cpp:22-23``) instead of ``array[idx] += val - array[idx]``;
  ancestors still receive ``diff = val - old`` in batch order (``cpp:21-29``);
* a negative value is an error per element (``cpp:15-17``).  The reference throws a C++ exception
  through a ``nogil`` extern without ``except +``, which aborts the process; here the batch is
  checked first and ``ValueError`` is raised before anything is written.

Stateless like the reference (caller-owned NumPy buffers): the pair is staged into a temporary
device tree for the call.  Code that cares about speed holds a ``DeviceSumTree`` and calls
``update_(..., assign_leaf=True)``.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _typed(a, dtype, name):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    if a.dtype != np.dtype(dtype):
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % ({"int32": "int", "float64": "double"}[np.dtype(dtype).name], a.dtype.name))
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return a


def update_tree_multi(idxs, vals, array, sumtree) -> None:
    """``array[idxs[i]] = vals[i % len(vals)]`` for i in order; every ancestor ``+= val - old``
    (reference ``sumtree_array.cpp:32-52``)."""
    idxs = _typed(idxs, np.int32, "idxs")
    vals, array, sumtree = (_typed(a, np.float64, n) for a, n in ((vals, "vals"), (array, "array"), (sumtree, "sumtree")))
    if array.size != sumtree.size:
        raise TypeError("array and sumtree must have the same size")
    if idxs.size == 0:
        return
    if vals.size == 0:
        raise ValueError("vals must not be empty")
    used = vals[: min(vals.size, idxs.size)]
    if (used < 0).any():
        raise ValueError("val must be non-negative")
    h = _lib.TreeHandle(array.size, np.float64)
    try:
        h.load_host(array, sumtree)
        h.update_host(idxs.astype(np.int64), np.ascontiguousarray(used), assign_leaf=True)
        h.store_host(array, sumtree)
    finally:
        h.close()


def get_prefix_sum_idx_multi(output, vals, array, sumtree):
    """Left-child descent per query into ``output`` (int32), which is returned
    (reference ``sumtree_array.cpp:54-85``)."""
    output = _typed(output, np.int32, "output")
    vals, array, sumtree = (_typed(a, np.float64, n) for a, n in ((vals, "vals"), (array, "array"), (sumtree, "sumtree")))
    if array.size != sumtree.size:
        raise TypeError("array and sumtree must have the same size")
    if output.size < vals.size:
        raise TypeError("output and vals must have the same size")
    h = _lib.TreeHandle(array.size, np.float64)
    try:
        h.load_host(array, sumtree)
        found = np.zeros(vals.size, dtype=np.intp)
        h.search_host(vals, found)
        output[: vals.size] = found
    finally:
        h.close()
    return output
