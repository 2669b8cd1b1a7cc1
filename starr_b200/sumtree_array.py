"""``SumTreeArray``: the reference's ndarray subclass, with the sum tree resident on a B200.

Mirrors the public behaviour of reference ``starr/sumtree_array.py:21-514`` (same constructor
forms, guarded writes, ``put`` / ``__setitem__``, in-place ufuncs, ``sample``,
``get_prefix_sum_id``, ``sum(axis, keepdims)``, ``mean``, ``sumtree()``, plain-ndarray results
for everything else), but the four calls the reference makes into its Cython module
(``sumtree_array.py:185, 233, 354, 508``) go to the CUDA kernels behind
``libstarr_b200.so`` instead.  State layout:

* the ndarray's own buffer is the HOST MIRROR of the leaves (NumPy C code reads it directly,
  so it cannot live on the device -- SURVEY H4);
* the device tree (``heap[2n]``) is the truth for tree nodes and for leaves;
* ``_sumtree`` is a host mirror of the internal nodes, refreshed lazily (it stays the same
  ndarray object across views, as the reference's tests require).

All views of one array share a single ``_TreeState``.
"""
from __future__ import annotations

import numpy as np

from . import _hostlogic, _lib

_TOO_SMALL = "input to SumTreeArray must have shape with at least 2 elements"
_DEVICE_UFUNCS = ("add", "subtract", "multiply", "true_divide", "divide")


class _TreeState:
    """Device tree + host mirrors shared by every view of one SumTreeArray."""

    __slots__ = ("handle", "leaves", "tree", "tree_stale", "leaves_stale", "_index_base", "_root", "__weakref__")

    def __init__(self, flat_leaves: np.ndarray, device: int = 0):
        self.leaves = flat_leaves                      # host mirror (a view of the array buffer)
        self.tree = np.zeros_like(flat_leaves)         # host mirror of heap[0:n)
        self.tree_stale = False
        self.leaves_stale = False
        self._index_base = None
        self._root = None                              # cached sumtree[1]; None = unknown
        self.handle = _lib.TreeHandle(flat_leaves.size, flat_leaves.dtype, device)

    # -- mirrors ----------------------------------------------------------------------------
    def host_tree(self) -> np.ndarray:
        if self.tree_stale:
            self.handle.store_host(None, self.tree)
            self.tree_stale = False
        return self.tree

    def sync_leaves(self) -> None:
        """Bring the host leaf mirror up to date after device-side updates."""
        if self.leaves_stale:
            if self.leaves.flags.c_contiguous and self.leaves.flags.writeable:
                self.handle.store_host(self.leaves, None)
            else:
                tmp = np.empty(self.leaves.size, self.leaves.dtype)
                self.handle.store_host(tmp, None)
                self.leaves[...] = tmp
            self.leaves_stale = False

    def root(self):
        if self._root is None:
            self._root = self.handle.root()
        return self._root

    def root_is_zero(self) -> bool:
        return self.root() == 0

    def note_root(self, value) -> None:
        self._root = value

    def invalidate_root(self) -> None:
        self._root = None

    def index_base(self) -> np.ndarray:
        if self._index_base is None:
            self._index_base = np.arange(self.leaves.size, dtype=np.intp)
        return self._index_base

    # -- the four kernel call sites ------------------------------------------------------------
    def rebuild(self) -> None:                       # reference sumtree_array.py:184-185
        self.leaves_stale = False                    # host leaves are the source here
        self._root = None
        try:
            self.handle.build_from_host(np.ascontiguousarray(self.leaves))
        finally:
            self.tree_stale = True

    def update(self, flat_idx: np.ndarray, values: np.ndarray) -> None:   # :233
        self.sync_leaves()
        self._root = None
        try:
            # the device holds fl(old + fl(val - old)), which is not always `val` for floats, so the
            # host mirror is patched with what the device actually stored (same staging round)
            # ... and the new root comes back with it, so that sum() / the zero check of sample() need no call
            written, root = self.handle.update_gather_root_host(flat_idx, values)
            if flat_idx.size:
                self._root = root
        finally:
            self.tree_stale = True
        if flat_idx.size:
            self.leaves[flat_idx] = written

    def search(self, flat_vals: np.ndarray) -> np.ndarray:                # :354
        out = np.zeros(flat_vals.size, dtype=np.intp)
        self.handle.search_host(flat_vals, out)
        return out

    def strided_sum(self, stride: int) -> np.ndarray:                     # :508
        return self.handle.strided_sum_host(stride)

    def inplace(self, op: str, scalar: np.ndarray) -> None:               # :150-173, :202-205
        """In-place leaf arithmetic + rebuild entirely on the device: no host ufunc over
        the n leaves and no n-element upload.  The host leaf mirror goes stale and is refreshed lazily."""
        self._root = None
        try:
            self.handle.inplace_device(op, scalar)
            self.leaves_stale = True
            self.handle.build()          # validates the NEW leaves; on failure the tree stays as it was
        finally:
            self.tree_stale = True


def _guarded(method):
    """Run ``method`` with the array temporarily writeable, then lock it again (the reference
    keeps a SumTreeArray read-only so that only its own API can touch the leaves)."""

    def wrapper(self, *args, **kwargs):
        self._enable_writes(True)
        try:
            return method(self, *args, **kwargs)
        finally:
            self._enable_writes(False)

    wrapper.__name__ = getattr(method, "__name__", "wrapper")
    wrapper.__doc__ = method.__doc__
    return wrapper


class SumTreeArray(np.ndarray):
    """ndarray of non-negative numbers with a device-resident sum tree.

    ``SumTreeArray(shape | ndarray | SumTreeArray, dtype=None)``; fast ``sample``,
    ``get_prefix_sum_id`` and ``sum`` over C-contiguous blocks, O(log n) writes through
    ``x[idx] = v`` / ``put``.  Everything that is not an in-place update returns a plain
    ``numpy.ndarray`` (copy or read-only view) so the tree can never be bypassed.
    """

    # ---- construction (reference :87-137) -----------------------------------------------------
    def __new__(cls, shape_or_array, dtype=None):
        if isinstance(shape_or_array, SumTreeArray):
            if dtype is None:
                return shape_or_array
            return SumTreeArray(shape_or_array._nd(), dtype)
        if isinstance(shape_or_array, np.ndarray):
            if shape_or_array.size <= 1:
                raise ValueError(_TOO_SMALL)
            src_dtype = shape_or_array.dtype if dtype is None else dtype
            return shape_or_array.astype(src_dtype, copy=True).view(SumTreeArray)
        fresh = np.zeros(shape_or_array, dtype=float if dtype is None else dtype)
        if fresh.size <= 1:
            raise ValueError(_TOO_SMALL)
        return fresh.view(SumTreeArray)

    @_guarded
    def __array_finalize__(self, array):
        if array is None or not np.shares_memory(array, self):
            raise NotImplementedError("input array and self must share memory")
        parent = self.base
        if isinstance(parent, SumTreeArray) and parent.dtype == self.dtype:
            if array.size != self.size:
                raise NotImplementedError(
                    "input array and base SumTreeArray must have same number of elements")
            self._state = parent._state
        else:
            _lib.dtype_code(self.dtype)   # TypeError for dtypes the kernels do not have
            state = _TreeState(array.view(np.ndarray).ravel())
            self._state = state
            state.rebuild()

    # names the reference (and its tests) expose
    @property
    def _flat_base(self) -> np.ndarray:
        self._state.sync_leaves()
        return self._state.leaves

    @property
    def _sumtree(self) -> np.ndarray:
        return self._state.host_tree()

    @property
    def _indices(self) -> np.ndarray:
        # flat leaf id of every element; one shared arange (8 B/leaf) created on first use
        return self._state.index_base().reshape(self.shape)

    @property
    def device_tree(self):
        """The ``_lib.TreeHandle`` of the resident tree (batched device entry points)."""
        return self._state.handle

    def mark_device_updated(self) -> None:
        """Tell the host mirrors that the device tree was changed behind their back."""
        self._state.tree_stale = True
        self._state.leaves_stale = True
        self._state.invalidate_root()

    def _nd(self) -> np.ndarray:
        """Plain, read-only ndarray view of the (up-to-date) host leaf mirror."""
        self._state.sync_leaves()
        return self.view(np.ndarray)

    def _enable_writes(self, val):
        self.setflags(write=val)

    def _rebuild_sumtree(self):
        self._state.rebuild()

    # ---- ufuncs: out-of-place -> ndarray, in-place on self -> rebuild (reference :144-173) --------
    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        out = out_arr.view(np.ndarray)
        return out[()] if return_scalar else out

    def _device_scalar(self, ufunc, method, inputs, kwargs):
        """The operand of ``self (op)= scalar`` as a 1-element array of this dtype if the device can
        reproduce NumPy's result bit for bit, else None (-> host ufunc + rebuild, as the reference does).
        Eligible: add / subtract / multiply / true_divide called in place on self with a Python scalar
        or a NumPy scalar of the array's own dtype (then NumPy computes in that dtype, one IEEE op)."""
        if method != "__call__" or ufunc.__name__ not in _DEVICE_UFUNCS or len(inputs) != 2 or inputs[0] is not self:
            return None
        if set(kwargs) - {"out"}:
            return None
        dt = self.dtype
        if ufunc.__name__ in ("true_divide", "divide") and dt.kind != "f":
            return None
        c = inputs[1]
        if isinstance(c, (bool, np.bool_)):
            return None
        if isinstance(c, np.generic):
            return np.array([c]) if c.dtype == dt else None
        if isinstance(c, int):
            if dt.kind == "f":
                try:
                    return np.array([c], dtype=dt)
                except OverflowError:
                    return None
            info = np.iinfo(dt)
            return np.array([c], dtype=dt) if info.min <= c <= info.max else None
        if isinstance(c, float) and dt.kind == "f":
            return np.array([c], dtype=dt)
        return None

    @_guarded
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        out = kwargs.get("out")
        if out is not None and len(out) == 1 and out[0] is self:
            scalar = self._device_scalar(ufunc, method, inputs, kwargs)
            if scalar is not None:
                self._state.inplace(ufunc.__name__, scalar)
                return self
        self._state.sync_leaves()
        plain = tuple(x.view(np.ndarray) if isinstance(x, SumTreeArray) else x for x in inputs)
        if out is None:
            return getattr(ufunc, method)(*plain, **kwargs)
        if len(out) == 1 and out[0] is self:
            kwargs["out"] = (self.view(np.ndarray),)
            getattr(ufunc, method)(*plain, **kwargs)
            self._rebuild_sumtree()
            return self
        raise NotImplementedError

    # ---- writes --------------------------------------------------------------------------------
    def __setitem__(self, idx, val):
        self.put(idx, val)

    def put(self, indices, values):
        """Set ``self[indices] = values``; ``values`` cycles over the selected elements in
        C order (reference :229-233 -- not NumPy broadcasting)."""
        flat = self._flat_ids(indices)
        vals = np.ascontiguousarray(values, dtype=self._state.leaves.dtype).ravel()
        self._state.update(flat, vals)

    def _flat_ids(self, indices) -> np.ndarray:
        """Flat leaf ids selected by an index expression, in C order (reference :231, ``self._indices[indices]``).
        A 1-D array indexed by an integer array needs no lookup table (the ids ARE the indices, negatives wrapped):
        that is the PER case, and it skips a random gather from an 8-byte-per-leaf arange.  Everything else -- and
        any out-of-range index, so that NumPy raises its own IndexError -- goes through the table."""
        if self.ndim == 1 and isinstance(indices, np.ndarray) and indices.dtype.kind in "iu" and indices.ndim >= 1:
            n = self.shape[0]
            ids = indices.astype(np.intp, copy=False).ravel()
            if ids.size == 0:
                return np.ascontiguousarray(ids)
            lo, hi = int(ids.min()), int(ids.max())
            if -n <= lo and hi < n:
                if lo < 0:
                    ids = np.where(ids < 0, ids + n, ids)
                return np.ascontiguousarray(ids)
        return np.ascontiguousarray(self._indices[indices]).ravel()

    @_guarded
    def fill(self, *args, **kwargs):
        if len(args) == 1 and not kwargs and isinstance(args[0], (int, float, np.generic)) \
                and not isinstance(args[0], (bool, np.bool_, complex, np.complexfloating)):
            sc = np.empty(1, dtype=self.dtype)
            sc.fill(args[0])             # NumPy's own conversion of the fill value (or its error)
            self._state.inplace("fill", sc)
            return
        self._state.sync_leaves()
        super().fill(*args, **kwargs)
        self._rebuild_sumtree()

    @_guarded
    def partition(self, *args, **kwargs):
        self._state.sync_leaves()
        super().partition(*args, **kwargs)
        self._rebuild_sumtree()

    @_guarded
    def sort(self, *args, **kwargs):
        self._state.sync_leaves()
        super().sort(*args, **kwargs)
        self._rebuild_sumtree()

    # ---- reads that must not create a second tree ----------------------------------------------
    def __getitem__(self, idx):
        return self._nd()[idx]

    def copy(self, *args, **kwargs):
        return SumTreeArray(self._nd())

    def squeeze(self):
        return self.reshape(self.view(np.ndarray).squeeze().shape)

    @property
    def imag(self):
        return self._nd().imag

    @property
    def real(self):
        return self._nd().real

    @property
    def T(self):
        return self._nd().T

    def mean(self, *args, **kwargs):
        total = self.sum(*args, **kwargs)
        groups = np.array(total).size
        assert self.size % groups == 0
        return total / float(self.size / groups)

    def sumtree(self):
        """A copy of the flat tree: ``[0, root, node 2, ...]`` (reference :253-276)."""
        return np.copy(self._sumtree)

    # ---- sampling (reference :294-421) -----------------------------------------------------------
    def get_prefix_sum_id(self, prefix_sum, flatten_indices=False):
        """Index ``j`` of the leaf at which the running sum of ``self.ravel()`` first exceeds
        each entry of ``prefix_sum`` (ties go right).  Output has the shape of ``prefix_sum``;
        for an n-d array a tuple of per-axis index arrays unless ``flatten_indices``."""
        queries = np.ascontiguousarray(prefix_sum, dtype=self.dtype)
        found = self._state.search(queries.ravel())
        if queries.ndim > 1:
            found = found.reshape(queries.shape)
        if self.ndim <= 1 or flatten_indices:
            return found
        return np.unravel_index(found, self.shape)

    def sample(self, nsamples=1, flatten_indices=False):
        """Draw ``nsamples`` indices with probability ``self / self.sum()``."""
        if self._state.root_is_zero():
            raise ValueError("array must have at least 1 positive value")
        # vals = (total * rand).astype(dtype), evaluated on the device with the same roundings; the
        # call also returns the root it scaled by, which refreshes the cached zero-total check
        u = np.random.rand(nsamples)
        found = np.zeros(u.size, dtype=np.intp)
        root = self._state.handle.sample_root_host(np.ascontiguousarray(u), found)
        self._state.note_root(root)
        if self.ndim <= 1 or flatten_indices:
            return found
        return np.unravel_index(found, self.shape)

    # ---- sums (reference :423-514) ---------------------------------------------------------------
    def _parse_axis_arg(self, axis):
        return _hostlogic.parse_axes(self.ndim, axis)

    def sum(self, axis=None, keepdims=False):
        """``numpy.sum`` semantics; sums over trailing (C-contiguous) axes are read off the tree,
        other contiguous axis blocks are folded on the device in row order."""
        axes = self._parse_axis_arg(axis)
        if axes is None:
            root = self._state.root()
            return np.array([root]).reshape([1] * self.ndim) if keepdims else root
        plan = _hostlogic.plan_sum(self.shape, axes)
        if plan[0] == "strided":
            out = self._state.strided_sum(plan[1])
            # the reference returns the leading dims flattened (a quirk we keep)
            return out.reshape(list(out.shape) + [1] * len(axes)) if keepdims else out
        if plan[0] == "fold":
            _, A, R, C = plan
            out = self._state.handle.axis_fold_host(A, R, C)
            return out.reshape(_hostlogic.fold_output_shape(self.shape, axes, keepdims))
        return self._nd().sum(axis=axis, keepdims=keepdims)


def _passthrough(name):
    def method(self, *args, **kwargs):
        return getattr(self._nd(), name)(*args, **kwargs)

    method.__name__ = name
    method.__doc__ = "``numpy.ndarray.%s`` of the underlying array; returns a plain ndarray." % name
    return method


# reference :181-292: these never return a SumTreeArray (no second tree is ever built)
for _name in ("astype", "choose", "diagonal", "dot", "flatten", "repeat", "round", "swapaxes",
              "take", "trace", "transpose"):
    setattr(SumTreeArray, _name, _passthrough(_name))
if hasattr(np.ndarray, "newbyteorder"):  # removed in NumPy 2.0
    SumTreeArray.newbyteorder = _passthrough("newbyteorder")
del _name
