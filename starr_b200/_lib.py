"""ctypes binding of ``libstarr_b200.so`` (C ABI declared in ``include/starr_b200.h``).

This module is the only place that touches the shared library.  There is no CPU fallback:
if the library is missing it raises at import of the first symbol, and if there is no CUDA
device ``st_create`` fails with ``ST_ERR_CUDA`` and we raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STARR_B200_LIB") or os.path.join(_HERE, "libstarr_b200.so")   # override: tuning builds only

# st_dtype codes (include/starr_b200.h) <-> NumPy dtypes: the members of the reference's fused
# ARRAY_TYPE (starr/src/_sumtree_array.pyx:4-11) that exist on the GPU.
DTYPE_CODES = {
    np.dtype(np.int16): 0,
    np.dtype(np.int32): 1,
    np.dtype(np.int64): 2,
    np.dtype(np.uint8): 3,
    np.dtype(np.uint16): 4,
    np.dtype(np.uint32): 5,
    np.dtype(np.uint64): 6,
    np.dtype(np.float32): 7,
    np.dtype(np.float64): 8,
}
# NumPy's default accumulator of ndarray.sum, which st_axis_fold_* mirrors
FOLD_DTYPES = {0: np.int64, 1: np.int64, 2: np.int64, 3: np.uint64, 4: np.uint64, 5: np.uint64,
               6: np.uint64, 7: np.float32, 8: np.float64}

ST_OK = 0
ST_ERR_NEGATIVE_VALS = 1
ST_ERR_NEGATIVE_ARRAY = 2
ST_ERR_INDEX = 3
ST_ERR_EMPTY_VALS = 4
ST_ERR_SIZE = 5
ST_ERR_DTYPE = 6
ST_ERR_ARG = 7
ST_ERR_CUDA = 8
ST_ERR_TOO_LARGE = 9

ST_UPDATE_SYNC_CHECK = 1
ST_UPDATE_ASSIGN_LEAF = 2
ST_OPS = {"add": 0, "subtract": 1, "multiply": 2, "true_divide": 3, "divide": 3, "fill": 4}
# dtype of st_sample_weights_device's probabilities: the tree dtype for floats, float64 for integers
PROB_DTYPES = {c: (np.float32 if c == 7 else np.float64) for c in range(9)}

# status -> (exception type, message of the reference where it has one)
_EXC = {
    ST_ERR_NEGATIVE_VALS: (ValueError, "vals must be non-negative"),
    ST_ERR_NEGATIVE_ARRAY: (ValueError, "array elements must be non-negative"),
    ST_ERR_INDEX: (IndexError, "Out of bounds on buffer access (axis 0)"),
    ST_ERR_EMPTY_VALS: (ZeroDivisionError, "integer division or modulo by zero"),
    ST_ERR_SIZE: (TypeError, None),
    ST_ERR_DTYPE: (TypeError, "No matching signature found"),
    ST_ERR_ARG: (ValueError, None),
    ST_ERR_CUDA: (RuntimeError, None),
    ST_ERR_TOO_LARGE: (ValueError, None),
}

_SIGNATURES = {
    "st_abi_version": (c_int, []),
    "st_status_string": (c_char_p, [c_int]),
    "st_last_error": (c_char_p, []),
    "st_dtype_size": (c_int, [c_int]),
    "st_kernel_launches": (ctypes.c_ulonglong, []),
    "st_create": (c_int, [c_int64, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "st_destroy": (c_int, [c_void_p]),
    "st_size": (c_int64, [c_void_p]),
    "st_dtype_of": (c_int, [c_void_p]),
    "st_device_of": (c_int, [c_void_p]),
    "st_heap_ptr": (c_void_p, [c_void_p]),
    "st_reserve": (c_int, [c_void_p, c_int64]),
    "st_load_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "st_store_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "st_gather_leaves_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "st_gather_leaves_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_build": (c_int, [c_void_p, c_void_p]),
    "st_build_from_host": (c_int, [c_void_p, c_void_p]),
    "st_build_from_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "st_update_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int]),
    "st_update_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64]),
    "st_update_host_ex": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int]),
    "st_update_deltas_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_apply_deltas_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_top_fold_record_bytes": (c_int, [c_void_p]),
    "st_top_fold_maps_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "st_top_fold_apply_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_shard_plan_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "st_shard_plan_wait": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "st_shard_plan_counts_device": (c_void_p, [c_void_p, c_int]),
    "st_search_counted_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "st_shard_expand_values_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_shard_expand_indices_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                               c_void_p]),
    "st_route_uniform_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "st_update_gather_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "st_update_gather_root_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_sample_root_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_search_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_search_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "st_sample_uniform_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "st_sample_uniform_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "st_root_host": (c_int, [c_void_p, c_void_p]),
    "st_sum_over_host": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "st_strided_sum_device": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "st_strided_sum_host": (c_int, [c_void_p, c_int64, c_void_p, c_int64]),
    "st_axis_fold_device": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "st_axis_fold_host": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "st_poll_status": (c_int, [c_void_p, c_void_p, c_int]),
    "st_update_stats": (c_int, [c_void_p, c_void_p, c_void_p]),
    "st_set_profiling": (c_int, [c_void_p, c_int]),
    "st_set_l2_persistence": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "st_update_breakdown": (c_int, [c_void_p, c_void_p]),
    "st_update_begin_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "st_update_finish_device": (c_int, [c_void_p, c_void_p]),
    "st_shard_plan_pos_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "st_peer_last_error": (c_char_p, []),
    "st_peer_handle_bytes": (c_int, []),
    "st_peer_create": (c_int, [c_int, c_int, c_int, c_int64, POINTER(c_void_p), c_void_p]),
    "st_peer_connect": (c_int, [c_void_p, c_void_p]),
    "st_peer_adopt": (c_int, [c_void_p, c_int, c_void_p]),
    "st_peer_base": (c_void_p, [c_void_p]),
    "st_peer_disconnect": (c_int, [c_void_p]),
    "st_peer_destroy": (c_int, [c_void_p]),
    "st_peer_payload": (c_void_p, [c_void_p]),
    "st_peer_payload_bytes": (c_int64, [c_void_p]),
    "st_peer_broadcast": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "st_peer_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "st_peer_put": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "st_peer_scatter_pairs": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "st_search_pairs_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "st_shard_expand_cat_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "st_peer_signal": (c_int, [c_void_p, c_int, ctypes.c_uint, c_void_p]),
    "st_peer_wait": (c_int, [c_void_p, c_int, ctypes.c_uint, c_void_p]),
    "st_peer_status": (c_int, [c_void_p, c_void_p]),
    "st_route_compact_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "st_search_push_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                      c_void_p]),
    "st_shard_partition_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "st_slices_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, ctypes.c_uint, c_void_p]),
    "st_slices_matrix_fetch": (c_int, [c_void_p, c_int64, c_void_p]),
    "st_slices_matrix_wait": (c_int, [c_void_p, c_void_p]),
    "st_slices_send": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "st_slices_return": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "st_slices_unpermute": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_void_p, c_int64,
                                    c_void_p]),
    "st_top_fold_maps_slice_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "st_slices_table": (c_int, [c_void_p, c_int, c_int64, c_int64]),
    "st_top_fold_apply_slices_device": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "st_route_push_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "st_slices_counts": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, ctypes.c_uint, c_void_p]),
    "st_search_segments_push_device": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "st_search_leaf32_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "st_sample_weights_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "st_per_step_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "st_per_step_host_submit": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                        c_void_p, c_void_p]),
    "st_per_step_host_wait": (c_int, [c_void_p, c_int]),
    "st_host_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "st_host_free": (c_int, [c_void_p]),
    "st_inplace_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "st_min_create": (c_int, [c_int64, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "st_min_update_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "st_min_update_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64]),
    "st_min_build": (c_int, [c_void_p, c_void_p]),
    "st_update_counters": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "st_reserve_sharded": (c_int, [c_void_p, c_int64, c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises ``ImportError`` if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m starr_b200._build` "
                "(or __graft_entry__.build()).  starr_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, size_message: str | None = None) -> None:
    """Raise the reference's exception for a non-zero st_status."""
    if status == ST_OK:
        return
    exc, msg = _EXC.get(status, (RuntimeError, None))
    detail = lib().st_last_error().decode("utf-8", "replace")
    if status == ST_ERR_SIZE and size_message:
        msg = size_message
    raise exc(msg if msg is not None else detail)


def dtype_code(dtype) -> int:
    try:
        return DTYPE_CODES[np.dtype(dtype)]
    except (KeyError, TypeError):
        # same failure the reference's fused-type dispatch produces (pyx:4-11)
        raise TypeError("No matching signature found") from None


def host_ptr(a: np.ndarray) -> c_void_p:
    return c_void_p(a.ctypes.data)


class TreeHandle:
    """Owner of one ``st_tree*``.  Thin: argument marshalling and status -> exception only."""

    def __init__(self, n: int, dtype, device: int = 0, external_heap: int | None = None, kind: str = "sum"):
        self.dtype = np.dtype(dtype)
        self.code = dtype_code(self.dtype)
        self.n = int(n)
        self.device = int(device)
        self.kind = kind
        out = c_void_p()
        if kind == "min":
            check(lib().st_min_create(self.n, self.code, self.device,
                                      c_void_p(external_heap) if external_heap else None, ctypes.byref(out)))
        else:
            check(lib().st_create(self.n, self.code, self.device,
                                  c_void_p(external_heap) if external_heap else None,
                                  ctypes.byref(out)))
        self._h = out

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.st_destroy(h)

    def __del__(self):  # noqa: D401
        try:
            self.close()
        except Exception:
            pass

    @property
    def ptr(self) -> c_void_p:
        if not self._h:
            raise RuntimeError("tree handle is closed")
        return self._h

    @property
    def heap_ptr(self) -> int:
        return int(lib().st_heap_ptr(self.ptr))

    def reserve(self, max_batch: int) -> None:
        check(lib().st_reserve(self.ptr, int(max_batch)))

    # -- host entry points (NumPy buffers) ---------------------------------------------------
    def load_host(self, leaves: np.ndarray | None, tree: np.ndarray | None) -> None:
        check(lib().st_load_host(self.ptr, host_ptr(leaves) if leaves is not None else None,
                                 host_ptr(tree) if tree is not None else None))

    def store_host(self, leaves: np.ndarray | None, tree: np.ndarray | None) -> None:
        check(lib().st_store_host(self.ptr, host_ptr(leaves) if leaves is not None else None,
                                  host_ptr(tree) if tree is not None else None))

    def gather_leaves_host(self, idx: np.ndarray) -> np.ndarray:
        out = np.empty(idx.size, dtype=self.dtype)
        check(lib().st_gather_leaves_host(self.ptr, host_ptr(idx), idx.size, host_ptr(out)))
        return out

    def build(self, stream: int = 0) -> None:
        check(lib().st_build(self.ptr, c_void_p(stream)))

    def build_from_host(self, leaves: np.ndarray) -> None:
        check(lib().st_build_from_host(self.ptr, host_ptr(leaves)))

    def update_host(self, idx: np.ndarray, vals: np.ndarray, assign_leaf: bool = False) -> None:
        check(lib().st_update_host_ex(self.ptr, host_ptr(idx), idx.size, host_ptr(vals), vals.size,
                                      ST_UPDATE_ASSIGN_LEAF if assign_leaf else 0))

    def update_gather_host(self, idx: np.ndarray, vals: np.ndarray) -> np.ndarray:
        out = np.empty(idx.size, dtype=self.dtype)
        check(lib().st_update_gather_host(self.ptr, host_ptr(idx), idx.size, host_ptr(vals), vals.size,
                                          host_ptr(out)))
        return out

    def update_gather_root_host(self, idx: np.ndarray, vals: np.ndarray):
        """``update_gather_host`` that also returns the root after the update (same staging round, one sync)."""
        out = np.empty(idx.size, dtype=self.dtype)
        root = np.zeros(1, dtype=self.dtype)
        check(lib().st_update_gather_root_host(self.ptr, host_ptr(idx), idx.size, host_ptr(vals), vals.size,
                                               host_ptr(out), host_ptr(root)))
        return out, root[0]

    def sample_root_host(self, u: np.ndarray, out: np.ndarray):
        root = np.zeros(1, dtype=self.dtype)
        check(lib().st_sample_root_host(self.ptr, host_ptr(u), u.size, host_ptr(out), host_ptr(root)))
        return root[0]

    def search_host(self, vals: np.ndarray, out: np.ndarray) -> None:
        check(lib().st_search_host(self.ptr, host_ptr(vals), vals.size, host_ptr(out)))

    def sample_uniform_host(self, u: np.ndarray, out: np.ndarray) -> None:
        check(lib().st_sample_uniform_host(self.ptr, host_ptr(u), u.size, host_ptr(out)))

    def root(self):
        out = np.zeros(1, dtype=self.dtype)
        check(lib().st_root_host(self.ptr, host_ptr(out)))
        return out[0]

    def sum_over(self, a: int, b: int):
        out = np.zeros(1, dtype=self.dtype)
        check(lib().st_sum_over_host(self.ptr, int(a), int(b), host_ptr(out)))
        return out[0]

    def strided_sum_host(self, stride: int) -> np.ndarray:
        stride = int(stride)
        nout = self.n // stride + (1 if self.n % stride else 0)
        out = np.empty(nout, dtype=self.dtype)
        check(lib().st_strided_sum_host(self.ptr, stride, host_ptr(out), nout), "invalid output size")
        return out

    def axis_fold_host(self, A: int, R: int, C: int) -> np.ndarray:
        out = np.empty((A, C), dtype=FOLD_DTYPES[self.code])
        check(lib().st_axis_fold_host(self.ptr, int(A), int(R), int(C), host_ptr(out)))
        return out

    # -- device entry points (raw device pointers + stream) ------------------------------------
    def build_from_device(self, leaves_ptr: int, stream: int = 0) -> None:
        check(lib().st_build_from_device(self.ptr, c_void_p(leaves_ptr), c_void_p(stream)))

    def update_device(self, idx_ptr: int, ni: int, vals_ptr: int, nv: int, stream: int = 0,
                      sync_check: bool = False, assign_leaf: bool = False) -> None:
        flags = (ST_UPDATE_SYNC_CHECK if sync_check else 0) | (ST_UPDATE_ASSIGN_LEAF if assign_leaf else 0)
        check(lib().st_update_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(vals_ptr), int(nv),
                                     c_void_p(stream), flags))

    def sample_weights_device(self, u_ptr: int, m: int, out_ptr: int, prio_ptr: int = 0, prob_ptr: int = 0,
                              stream: int = 0) -> None:
        check(lib().st_sample_weights_device(self.ptr, c_void_p(u_ptr), int(m), c_void_p(out_ptr),
                                             c_void_p(prio_ptr) if prio_ptr else None,
                                             c_void_p(prob_ptr) if prob_ptr else None, c_void_p(stream)))

    def per_step_device(self, idx_ptr: int, ni: int, vals_ptr: int, nv: int, u_ptr: int, m: int, out_ptr: int,
                        prio_ptr: int = 0, prob_ptr: int = 0, stream: int = 0) -> None:
        check(lib().st_per_step_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(vals_ptr), int(nv), c_void_p(u_ptr),
                                       int(m), c_void_p(out_ptr), c_void_p(prio_ptr) if prio_ptr else None,
                                       c_void_p(prob_ptr) if prob_ptr else None, c_void_p(stream)))

    def per_step_host_submit(self, slot: int, idx: np.ndarray, vals: np.ndarray, u: np.ndarray, out_idx: np.ndarray,
                             out_prio: np.ndarray | None = None, out_prob: np.ndarray | None = None) -> None:
        """Pipelined host step (slots 0 .. 3); the arrays must stay alive and untouched until per_step_host_wait."""
        check(lib().st_per_step_host_submit(self.ptr, int(slot), host_ptr(idx), idx.size, host_ptr(vals), vals.size,
                                            host_ptr(u), u.size, host_ptr(out_idx),
                                            host_ptr(out_prio) if out_prio is not None else None,
                                            host_ptr(out_prob) if out_prob is not None else None))

    def per_step_host_wait(self, slot: int) -> None:
        check(lib().st_per_step_host_wait(self.ptr, int(slot)))

    def inplace_device(self, op: str, scalar: np.ndarray | None, operand_ptr: int = 0, operand_len: int = 0,
                       stream: int = 0) -> None:
        """leaves = leaves (op) scalar | operand; no rebuild (call build afterwards)."""
        check(lib().st_inplace_device(self.ptr, ST_OPS[op], host_ptr(scalar) if scalar is not None else None,
                                      c_void_p(operand_ptr) if operand_ptr else None, int(operand_len),
                                      c_void_p(stream)))

    def min_update_device(self, idx_ptr: int, ni: int, vals_ptr: int, nv: int, stream: int = 0) -> None:
        check(lib().st_min_update_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(vals_ptr), int(nv),
                                         c_void_p(stream)))

    def min_update_host(self, idx: np.ndarray, vals: np.ndarray) -> None:
        check(lib().st_min_update_host(self.ptr, host_ptr(idx), idx.size, host_ptr(vals), vals.size))

    def min_build(self, stream: int = 0) -> None:
        check(lib().st_min_build(self.ptr, c_void_p(stream)))

    def update_counters(self, stream: int = 0, reset: bool = False) -> dict:
        out = (c_int64 * 2)()
        check(lib().st_update_counters(self.ptr, c_void_p(stream), out, 1 if reset else 0))
        return {"updates": int(out[0]), "zero_delta": int(out[1])}

    def reserve_sharded(self, max_batch: int, shards: int) -> None:
        check(lib().st_reserve_sharded(self.ptr, int(max_batch), int(shards)))

    def update_deltas_device(self, idx_ptr: int, ni: int, vals_ptr: int, nv: int, delta_out_ptr: int,
                             stream: int = 0) -> None:
        check(lib().st_update_deltas_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(vals_ptr), int(nv),
                                            c_void_p(delta_out_ptr), c_void_p(stream)))

    def apply_deltas_device(self, idx_ptr: int, ni: int, deltas_ptr: int, stream: int = 0) -> None:
        check(lib().st_apply_deltas_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(deltas_ptr),
                                           c_void_p(stream)))

    def top_fold_record_bytes(self) -> int:
        return int(lib().st_top_fold_record_bytes(self.ptr))

    def top_fold_maps_device(self, leaf_ptr: int, leaf_itemsize: int, deltas_ptr: int, m: int, chunk_lo: int,
                             chunk_hi: int, recs_ptr: int, stream: int = 0) -> None:
        check(lib().st_top_fold_maps_device(self.ptr, c_void_p(leaf_ptr), int(leaf_itemsize), c_void_p(deltas_ptr),
                                            int(m), int(chunk_lo), int(chunk_hi), c_void_p(recs_ptr),
                                            c_void_p(stream)))

    def top_fold_apply_device(self, leaf_ptr: int, leaf_itemsize: int, deltas_ptr: int, m: int, recs_ptr: int,
                              stream: int = 0) -> None:
        check(lib().st_top_fold_apply_device(self.ptr, c_void_p(leaf_ptr), int(leaf_itemsize), c_void_p(deltas_ptr),
                                             int(m), c_void_p(recs_ptr), c_void_p(stream)))

    def shard_plan_device(self, ids_ptr: int, m: int, shift: int, shards: int, rank: int, payload_ptr: int,
                          npayload: int, check_negative: bool, shard_out_ptr: int, slot_out_ptr: int,
                          own_ids_out_ptr: int, own_payload_out_ptr: int, result_slot: int = 0,
                          stream: int = 0, wait: bool = True) -> list[int] | None:
        """wait=False only enqueues; collect the counts with shard_plan_wait(result_slot, shards)."""
        counts = (c_int64 * int(shards))() if wait else None
        check(lib().st_shard_plan_device(self.ptr, c_void_p(ids_ptr), int(m), int(shift), int(shards), int(rank),
                                         c_void_p(payload_ptr), int(npayload), int(bool(check_negative)),
                                         c_void_p(shard_out_ptr), c_void_p(slot_out_ptr), c_void_p(own_ids_out_ptr),
                                         c_void_p(own_payload_out_ptr), int(result_slot), counts, c_void_p(stream)))
        return list(counts) if wait else None

    def shard_plan_pos_device(self, ids_ptr: int, m: int, shift: int, shards: int, rank: int, payload_ptr: int,
                              npayload: int, check_negative: bool, shard_out_ptr: int, slot_out_ptr: int,
                              own_ids_out_ptr: int, own_payload_out_ptr: int, own_pos_out_ptr: int, result_slot: int = 0,
                              stream: int = 0, wait: bool = True) -> list[int] | None:
        counts = (c_int64 * int(shards))() if wait else None
        check(lib().st_shard_plan_pos_device(self.ptr, c_void_p(ids_ptr), int(m), int(shift), int(shards), int(rank),
                                             c_void_p(payload_ptr), int(npayload), int(bool(check_negative)),
                                             c_void_p(shard_out_ptr), c_void_p(slot_out_ptr), c_void_p(own_ids_out_ptr),
                                             c_void_p(own_payload_out_ptr), c_void_p(own_pos_out_ptr), int(result_slot),
                                             counts, c_void_p(stream)))
        return list(counts) if wait else None

    def update_begin_device(self, idx_ptr: int, ni: int, vals_ptr: int, nv: int, delta_out_ptr: int, stream: int = 0) -> None:
        check(lib().st_update_begin_device(self.ptr, c_void_p(idx_ptr), int(ni), c_void_p(vals_ptr), int(nv),
                                           c_void_p(delta_out_ptr), c_void_p(stream)))

    def update_finish_device(self, stream: int = 0) -> None:
        check(lib().st_update_finish_device(self.ptr, c_void_p(stream)))

    def route_compact_device(self, u_ptr: int, m: int, rank: int, resid_ptr: int, pos_ptr: int, count_ptr: int,
                             stream: int = 0) -> None:
        check(lib().st_route_compact_device(self.ptr, c_void_p(u_ptr), int(m), int(rank), c_void_p(resid_ptr),
                                            c_void_p(pos_ptr), c_void_p(count_ptr), c_void_p(stream)))

    def search_push_device(self, vals_ptr: int, m_max: int, m_dev_ptr: int, pos_ptr: int, index_offset: int, peer: "PeerArea",
                           payload_offset: int, span_elems: int, stream: int = 0) -> None:
        check(lib().st_search_push_device(self.ptr, c_void_p(vals_ptr), int(m_max), c_void_p(m_dev_ptr), c_void_p(pos_ptr),
                                          int(index_offset), peer.ptr, int(payload_offset), int(span_elems), c_void_p(stream)))

    def search_pairs_device(self, vals_ptr: int, m_max: int, m_dev_ptr: int, pos_ptr: int, pairs_ptr: int, stream: int = 0) -> None:
        status = lib().st_search_pairs_device(self.ptr, c_void_p(vals_ptr), int(m_max), c_void_p(m_dev_ptr), c_void_p(pos_ptr),
                                              c_void_p(pairs_ptr), c_void_p(stream))
        if status != ST_OK:
            raise RuntimeError(lib().st_peer_last_error().decode("utf-8", "replace"))

    def shard_expand_cat_device(self, shard_ptr: int, slot_ptr: int, m: int, cat_ptr: int, starts: list, out_ptr: int,
                                stream: int = 0) -> None:
        arr = (c_int64 * len(starts))(*[int(x) for x in starts])
        check(lib().st_shard_expand_cat_device(self.ptr, c_void_p(shard_ptr), c_void_p(slot_ptr), int(m), c_void_p(cat_ptr), arr,
                                               len(starts), c_void_p(out_ptr), c_void_p(stream)))

    # ---- slices layout (batch distributed over the ranks) -------------------------------------------------------
    def shard_partition_device(self, ids_ptr: int, m: int, shift: int, shards: int, payload_ptr: int, check_negative: bool,
                               shard_out_ptr: int, slot_out_ptr: int, part_ids_out_ptr: int, part_payload_out_ptr: int,
                               row_out_ptr: int, stream: int = 0) -> None:
        check(lib().st_shard_partition_device(self.ptr, c_void_p(ids_ptr), int(m), int(shift), int(shards),
                                              c_void_p(payload_ptr) if payload_ptr else None, int(bool(check_negative)),
                                              c_void_p(shard_out_ptr), c_void_p(slot_out_ptr),
                                              c_void_p(part_ids_out_ptr) if part_ids_out_ptr else None,
                                              c_void_p(part_payload_out_ptr) if part_payload_out_ptr else None,
                                              c_void_p(row_out_ptr), c_void_p(stream)))

    @staticmethod
    def _pcheck(status: int) -> None:
        if status != ST_OK:
            raise RuntimeError(lib().st_peer_last_error().decode("utf-8", "replace"))

    def top_fold_maps_slice_device(self, leaf_ptr: int, deltas_ptr: int, m_slice: int, record_first: int, nchunks: int,
                                   recs_ptr: int, stream: int = 0) -> None:
        self._pcheck(lib().st_top_fold_maps_slice_device(self.ptr, c_void_p(leaf_ptr), c_void_p(deltas_ptr), int(m_slice),
                                                         int(record_first), int(nchunks), c_void_p(recs_ptr), c_void_p(stream)))

    def top_fold_apply_slices_device(self, peer: "PeerArea", parity: int, span: int, m_slice: int, recs_ptr: int,
                                     stream: int = 0) -> None:
        self._pcheck(lib().st_top_fold_apply_slices_device(self.ptr, peer.ptr, int(parity), int(span), int(m_slice),
                                                           c_void_p(recs_ptr), c_void_p(stream)))

    def route_push_device(self, u_ptr: int, m: int, peer: "PeerArea", resid_offset: int, pos_offset: int, cap: int,
                          cnt_ptr: int, stream: int = 0) -> None:
        self._pcheck(lib().st_route_push_device(self.ptr, c_void_p(u_ptr), int(m), peer.ptr, int(resid_offset), int(pos_offset),
                                                int(cap), c_void_p(cnt_ptr), c_void_p(stream)))

    def search_segments_push_device(self, peer: "PeerArea", resid_offset: int, pos_offset: int, cnt_offset: int, cap: int,
                                    back_offset: int, stream: int = 0) -> None:
        self._pcheck(lib().st_search_segments_push_device(self.ptr, peer.ptr, int(resid_offset), int(pos_offset), int(cnt_offset),
                                                          int(cap), int(back_offset), c_void_p(stream)))

    def search_leaf32_device(self, vals_ptr: int, m_max: int, m_dev_ptr: int, out_ptr: int, stream: int = 0) -> None:
        self._pcheck(lib().st_search_leaf32_device(self.ptr, c_void_p(vals_ptr), int(m_max), c_void_p(m_dev_ptr),
                                                   c_void_p(out_ptr), c_void_p(stream)))

    def shard_plan_wait(self, result_slot: int, shards: int) -> list[int]:
        counts = (c_int64 * int(shards))()
        check(lib().st_shard_plan_wait(self.ptr, int(result_slot), int(shards), counts))
        return list(counts)

    def shard_plan_counts_ptr(self, result_slot: int) -> int:
        return int(lib().st_shard_plan_counts_device(self.ptr, int(result_slot)) or 0)

    def search_counted_device(self, vals_ptr: int, m_max: int, m_dev_ptr: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().st_search_counted_device(self.ptr, c_void_p(vals_ptr), int(m_max), c_void_p(m_dev_ptr),
                                             c_void_p(out_ptr), c_void_p(stream)))

    def shard_expand_values_device(self, shard_ptr: int, slot_ptr: int, m: int, slabs_ptr: int, slab_stride: int,
                                   out_ptr: int, stream: int = 0) -> None:
        check(lib().st_shard_expand_values_device(self.ptr, c_void_p(shard_ptr), c_void_p(slot_ptr), int(m),
                                                  c_void_p(slabs_ptr), int(slab_stride), c_void_p(out_ptr),
                                                  c_void_p(stream)))

    def shard_expand_indices_device(self, shard_ptr: int, slot_ptr: int, m: int, slabs_ptr: int, slab_stride: int,
                                    leaves_per_shard: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().st_shard_expand_indices_device(self.ptr, c_void_p(shard_ptr), c_void_p(slot_ptr), int(m),
                                                   c_void_p(slabs_ptr), int(slab_stride), int(leaves_per_shard),
                                                   c_void_p(out_ptr), c_void_p(stream)))

    def route_uniform_device(self, u_ptr: int, m: int, leaf_out_ptr: int, resid_out_ptr: int,
                             stream: int = 0) -> None:
        check(lib().st_route_uniform_device(self.ptr, c_void_p(u_ptr), int(m), c_void_p(leaf_out_ptr),
                                            c_void_p(resid_out_ptr), c_void_p(stream)))

    def search_device(self, vals_ptr: int, m: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().st_search_device(self.ptr, c_void_p(vals_ptr), int(m), c_void_p(out_ptr), c_void_p(stream)))

    def sample_uniform_device(self, u_ptr: int, m: int, out_ptr: int, prio_ptr: int = 0,
                              stream: int = 0) -> None:
        check(lib().st_sample_uniform_device(self.ptr, c_void_p(u_ptr), int(m), c_void_p(out_ptr),
                                             c_void_p(prio_ptr) if prio_ptr else None, c_void_p(stream)))

    def gather_leaves_device(self, idx_ptr: int, m: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().st_gather_leaves_device(self.ptr, c_void_p(idx_ptr), int(m), c_void_p(out_ptr),
                                            c_void_p(stream)))

    def strided_sum_device(self, stride: int, out_ptr: int, nout: int, stream: int = 0) -> None:
        check(lib().st_strided_sum_device(self.ptr, int(stride), c_void_p(out_ptr), int(nout),
                                          c_void_p(stream)), "invalid output size")

    def axis_fold_device(self, A: int, R: int, C: int, out_ptr: int, stream: int = 0) -> None:
        check(lib().st_axis_fold_device(self.ptr, int(A), int(R), int(C), c_void_p(out_ptr), c_void_p(stream)))

    def poll_status(self, stream: int = 0, clear: bool = True) -> None:
        check(lib().st_poll_status(self.ptr, c_void_p(stream), 1 if clear else 0))

    def set_l2_persistence(self, nbytes: int, stream: int = 0) -> int:
        granted = c_int64(0)
        check(lib().st_set_l2_persistence(self.ptr, c_void_p(stream), int(nbytes), ctypes.byref(granted)))
        return int(granted.value)

    def set_profiling(self, enable: bool) -> None:
        check(lib().st_set_profiling(self.ptr, 1 if enable else 0))

    def update_breakdown(self) -> dict:
        out = (ctypes.c_double * 8)()
        check(lib().st_update_breakdown(self.ptr, out))
        # huge_stitch_ms: first-pass stitch of the huge segments + summaries for the far side of a binade
        # boundary (usually nothing to do); fold_scan_ms: second-pass stitch + long + medium folds (one launch)
        names = ("sort_ms", "leaf_pass_ms", "partition_ms", "retire_ms", "huge_maps_ms", "huge_stitch_ms", "fold_scan_ms")
        calls = max(out[7], 1.0)
        d = {n: out[i] / calls for i, n in enumerate(names)}
        d["calls"] = int(out[7])
        return d

    def update_stats(self, stream: int = 0) -> dict:
        out = (c_int64 * 4)()
        check(lib().st_update_stats(self.ptr, c_void_p(stream), out))
        # huge segments (more than 16384 updates under one node), cumulative since the tree was created
        return {"huge_segments_maps_only": out[0], "huge_segments_with_exact_chunks": out[1],
                "chunks_folded_exactly": out[2], "most_exact_chunks_in_one_segment": out[3]}


class PinnedBuffer:
    """Page-locked host memory from ``st_host_alloc`` with NumPy views over it (inputs / outputs of the
    pipelined host step, ``TreeHandle.per_step_host_submit``)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = c_void_p()
        check(lib().st_host_alloc(self.nbytes, ctypes.byref(p)))
        self._p = p
        self._raw = (ctypes.c_char * self.nbytes).from_address(p.value)

    def array(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        return np.frombuffer(self._raw, dtype=dtype, count=count, offset=offset)

    def close(self) -> None:
        p, self._p = getattr(self, "_p", None), None
        if p and _lib is not None:
            self._raw = None
            _lib.st_host_free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerArea:
    """One rank's side of the peer-memory exchange area (``st_peer_*``): a device buffer whose CUDA IPC
    handle is exchanged once, after which kernels store into the buffers of all ranks."""

    def __init__(self, device: int, rank: int, world: int, payload_bytes: int):
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        self._hb = int(lib().st_peer_handle_bytes())
        h = (ctypes.c_ubyte * self._hb)()
        out = c_void_p()
        self._check(lib().st_peer_create(self.device, self.rank, self.world, int(payload_bytes), ctypes.byref(out), h))
        self._p = out
        self.handle = bytes(h)

    @staticmethod
    def _check(status: int) -> None:
        if status != ST_OK:
            raise RuntimeError(lib().st_peer_last_error().decode("utf-8", "replace"))

    @property
    def ptr(self) -> c_void_p:
        if not self._p:
            raise RuntimeError("peer area is closed")
        return self._p

    def address(self) -> tuple:
        """What the other ranks need to reach this area: ``(pid, base pointer, CUDA IPC handle)``."""
        return (os.getpid(), int(lib().st_peer_base(self.ptr)), self.handle)

    def connect(self, addresses: list) -> None:
        """``addresses[q]`` = ``address()`` of rank q.  Ranks of this process (threads) are adopted by
        pointer, the others are opened through their IPC handle."""
        for q, (pid, base, _) in enumerate(addresses):
            if q != self.rank and pid == os.getpid():
                self._check(lib().st_peer_adopt(self.ptr, q, c_void_p(base)))
        blob = b"".join(a[2] for a in addresses)
        assert len(blob) == self.world * self._hb
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(lib().st_peer_connect(self.ptr, buf))

    @property
    def payload_ptr(self) -> int:
        return int(lib().st_peer_payload(self.ptr))

    @property
    def payload_bytes(self) -> int:
        return int(lib().st_peer_payload_bytes(self.ptr))

    def broadcast(self, offset: int, nbytes: int, stream: int = 0) -> None:
        self._check(lib().st_peer_broadcast(self.ptr, int(offset), int(nbytes), c_void_p(stream)))

    def scatter(self, src_ptr: int, pos_ptr: int, count: int, esize: int, offset: int, span_elems: int,
                count_dev_ptr: int = 0, stream: int = 0) -> None:
        self._check(lib().st_peer_scatter(self.ptr, c_void_p(src_ptr), c_void_p(pos_ptr), int(count),
                                          c_void_p(count_dev_ptr) if count_dev_ptr else None, int(esize), int(offset),
                                          int(span_elems), c_void_p(stream)))

    def put(self, src_ptr: int, count: int, esize: int, dst_offset: int, count_dev_ptr: int = 0, count_offset: int = -1,
            stream: int = 0) -> None:
        self._check(lib().st_peer_put(self.ptr, c_void_p(src_ptr), int(count), c_void_p(count_dev_ptr) if count_dev_ptr else None,
                                      int(esize), int(dst_offset), int(count_offset), c_void_p(stream)))

    def scatter_pairs(self, pairs_offset: int, slab_bytes: int, counts_offset: int, leaves_per_shard: int, out_ptr: int, m: int,
                      stream: int = 0) -> None:
        self._check(lib().st_peer_scatter_pairs(self.ptr, int(pairs_offset), int(slab_bytes), int(counts_offset),
                                                int(leaves_per_shard), c_void_p(out_ptr), int(m), c_void_p(stream)))

    # ---- slices layout (batch distributed over the ranks) -------------------------------------------------------
    def slices_rows(self, row_ptr: int, matrix_offset: int, channel: int, value: int, stream: int = 0) -> None:
        self._check(lib().st_slices_rows(self.ptr, c_void_p(row_ptr), int(matrix_offset), int(channel), int(value) & 0xffffffff,
                                         c_void_p(stream)))

    def slices_matrix_fetch(self, matrix_offset: int, stream: int = 0) -> None:
        self._check(lib().st_slices_matrix_fetch(self.ptr, int(matrix_offset), c_void_p(stream)))

    def slices_matrix_wait(self) -> list:
        """The count matrix as ``world`` rows of ``world`` counts + 1 word of validation flags."""
        w = self.world
        out = (ctypes.c_uint32 * (w * (w + 1)))()
        self._check(lib().st_slices_matrix_wait(self.ptr, out))
        return [list(out[r * (w + 1):(r + 1) * (w + 1)]) for r in range(w)]

    def slices_send(self, matrix_offset: int, src0: int, esize0: int, dst0: int, m_max: int, total_out_ptr: int = 0,
                    src1: int = 0, esize1: int = 0, dst1: int = 0, stream: int = 0) -> None:
        self._check(lib().st_slices_send(self.ptr, int(matrix_offset), c_void_p(src0), int(esize0), int(dst0),
                                         c_void_p(src1) if src1 else None, int(esize1), int(dst1), int(m_max),
                                         c_void_p(total_out_ptr) if total_out_ptr else None, c_void_p(stream)))

    def slices_return(self, matrix_offset: int, src: int, esize: int, dst: int, total_max: int, stream: int = 0) -> None:
        self._check(lib().st_slices_return(self.ptr, int(matrix_offset), c_void_p(src), int(esize), int(dst), int(total_max),
                                           c_void_p(stream)))

    def slices_unpermute(self, matrix_offset: int, shard_ptr: int, slot_ptr: int, m: int, ret_offset: int, esize: int,
                         out_ptr: int, leaves_per_shard: int = 0, shard_copy_offset: int = -1, stream: int = 0) -> None:
        self._check(lib().st_slices_unpermute(self.ptr, int(matrix_offset), c_void_p(shard_ptr), c_void_p(slot_ptr), int(m),
                                              int(ret_offset), int(esize), int(leaves_per_shard), c_void_p(out_ptr),
                                              int(shard_copy_offset), c_void_p(stream)))

    def slices_counts(self, cnt_ptr: int, cnt_offset: int, sent_offset: int, channel: int, value: int, stream: int = 0) -> None:
        self._check(lib().st_slices_counts(self.ptr, c_void_p(cnt_ptr), int(cnt_offset), int(sent_offset), int(channel),
                                           int(value) & 0xffffffff, c_void_p(stream)))

    def slices_table(self, parity: int, d_offset: int, leaf_offset: int) -> None:
        self._check(lib().st_slices_table(self.ptr, int(parity), int(d_offset), int(leaf_offset)))

    def signal(self, channel: int, value: int, stream: int = 0) -> None:
        self._check(lib().st_peer_signal(self.ptr, int(channel), int(value) & 0xffffffff, c_void_p(stream)))

    def wait(self, channel: int, value: int, stream: int = 0) -> None:
        self._check(lib().st_peer_wait(self.ptr, int(channel), int(value) & 0xffffffff, c_void_p(stream)))

    def status(self, stream: int = 0) -> None:
        self._check(lib().st_peer_status(self.ptr, c_void_p(stream)))

    def disconnect(self) -> None:
        if getattr(self, "_p", None):
            self._check(lib().st_peer_disconnect(self._p))

    def close(self) -> None:
        p, self._p = getattr(self, "_p", None), None
        if p and _lib is not None:
            _lib.st_peer_destroy(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
