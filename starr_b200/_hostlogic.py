"""Pure host-side decisions of the SumTreeArray API (no device, no library) -- unit-testable
on a machine without a GPU.

Reference behaviour restated: ``_parse_axis_arg`` (``starr/sumtree_array.py:423-432``) and the
tree-vs-NumPy choice in ``sum`` (``starr/sumtree_array.py:498-514``).
"""
from __future__ import annotations

import numpy as np


def parse_axes(ndim: int, axis):
    """None -> None; otherwise a 1-D array of non-negative axes.  Duplicates and out-of-range
    axes raise IndexError like the reference (fancy-indexing ``arange(ndim)``)."""
    if axis is None:
        return None
    axes = np.array(axis).reshape(-1)
    if len(set(axes.tolist())) < len(axes):
        raise IndexError("invalid axis argument: %s contains duplicates" % str(axis))
    return np.arange(ndim)[axes]


def plan_sum(shape, axes):
    """How ``sum(axis=axes)`` is evaluated for an array of ``shape``.

    ("strided", stride)   axes are the trailing ones: C-contiguous blocks of `stride` leaves,
                          one covering-node read (or an O(log stride) sibling walk) per output
                          -- the reference's tree path (sumtree_array.py:505-512).
    ("fold", A, R, C)     axes form one contiguous block that is not trailing: leaves viewed as
                          [A][R][C], sequential fold over R on the device (the order NumPy's
                          own reduction uses; replaces the ndarray.sum fallback, :514).
    ("host",)             scattered axes (e.g. (0, 2) of a 3-d array) or C == 1 with size-1
                          trailing dims: NumPy's pairwise order applies; evaluated by NumPy on
                          the host mirror exactly as the reference does.
    """
    shape = tuple(int(s) for s in shape)
    ndim = len(shape)
    axes = np.sort(np.asarray(axes, dtype=np.intp))
    k = len(axes)
    if k == 0:
        return ("host",)
    if int(axes.min()) == ndim - k:
        return ("strided", int(np.prod(np.array(shape)[axes])))
    lo, hi = int(axes[0]), int(axes[-1])
    if hi - lo + 1 != k:
        return ("host",)
    A = int(np.prod(shape[:lo], dtype=np.int64))
    R = int(np.prod(shape[lo:hi + 1], dtype=np.int64))
    C = int(np.prod(shape[hi + 1:], dtype=np.int64))
    if C <= 1 or A * R * C == 0:
        return ("host",)
    return ("fold", A, R, C)


def fold_output_shape(shape, axes, keepdims: bool):
    """Shape NumPy gives ``ndarray.sum(axis=axes, keepdims=keepdims)``."""
    drop = set(int(a) for a in axes)
    if keepdims:
        return tuple(1 if i in drop else int(s) for i, s in enumerate(shape))
    return tuple(int(s) for i, s in enumerate(shape) if i not in drop)


# ---- subtree sharding ---------------------------------------------------------
def shard_geometry(n_global: int, world: int):
    """Leaves [r*n/G, (r+1)*n/G) belong to rank r = the heap subtree rooted at global node G+r.
    Requires power-of-two G and power-of-two n/G (non-power-of-two trees rotate leaves across
    two depths and do not split into G equal subtrees)."""
    if world < 1 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    if n_global % world:
        raise ValueError("n must be divisible by the world size")
    n_local = n_global // world
    if n_local < 2 or n_local & (n_local - 1):
        raise ValueError("n / world must be a power of two >= 2")
    return n_local, world.bit_length() - 1
