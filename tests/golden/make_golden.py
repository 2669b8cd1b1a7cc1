"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the reference package (Python layer from /root/reference, kernel module compiled into
oracle/_ref by oracle/build_ref.sh), runs seeded cases through its own public API and stores the
inputs and outputs.  The fixtures travel with the repo; the GPU box never sees /root/reference.
Bit patterns are stored as raw arrays (np.savez keeps dtype and bytes exactly).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

HERE = os.path.dirname(os.path.abspath(__file__))


from cases import (API_SPECS, KERNEL_BATCHES, KERNEL_SPECS, KERNEL_STRIDES, api_script,  # noqa: E402
                   priorities)


def kernel_cases():
    """update / build / search / sums through the reference kernel module."""
    ref = oracle.ref()
    out = {}
    for ci, (name, dt, n, kind) in enumerate(KERNEL_SPECS):
        rng = np.random.default_rng(1000 + ci)
        leaves = priorities(rng, n, dt, kind)
        tree = np.zeros(n, dt)
        ref.build_sumtree_from_array(leaves, tree)
        out[f"{name}/leaves0"] = leaves.copy()
        out[f"{name}/tree0"] = tree.copy()
        for b in range(KERNEL_BATCHES):
            m = int(rng.integers(1, 3 * n))
            nv = m if b % 2 == 0 else int(rng.integers(1, m + 1))
            if b == 2:  # heavy duplicates: PER-style resampled indices
                idx = rng.integers(0, max(2, n // 16), m).astype(np.intp)
            else:
                idx = rng.integers(0, n, m).astype(np.intp)
            vals = priorities(rng, nv, dt, "loguniform" if kind != "uniform" else "uniform")
            ref.update_sumtree(idx, vals, leaves, tree)
            out[f"{name}/b{b}/idx"] = idx
            out[f"{name}/b{b}/vals"] = vals
            out[f"{name}/b{b}/leaves"] = leaves.copy()
            out[f"{name}/b{b}/tree"] = tree.copy()
        u = rng.random(512)
        q = (tree[1] * u).astype(dt)
        # edge queries: 0, the total, beyond the total
        q[:3] = np.array([0, tree[1], tree[1]], dtype=dt)
        found = np.zeros(q.size, np.intp)
        ref.get_prefix_sum_idx(found, q, leaves, tree)
        out[f"{name}/u"] = u
        out[f"{name}/queries"] = q
        out[f"{name}/found"] = found
        for s in KERNEL_STRIDES + (n,):
            out[f"{name}/strided{s}"] = ref.strided_sum(leaves, tree, s)
        pairs = np.sort(rng.integers(0, n + 1, (64, 2)), axis=1)
        out[f"{name}/ranges"] = pairs
        out[f"{name}/range_sums"] = np.array([ref.sum_over(leaves, tree, int(a), int(b)) for a, b in pairs],
                                             dtype=dt)
    return out


def api_cases():
    """SumTreeArray through the reference's Python layer (script in cases.py)."""
    api = oracle.ref_api()
    out = {}
    for name, dt, shape in API_SPECS:
        for k, v in api_script(api.SumTreeArray, name, dt, shape).items():
            out[f"{name}/{k}"] = v
    return out


def next_cases(reference_root="/root/reference"):
    """the experimental C++ twin (oracle/_ref/_xcython*.so, unmodified) and the pure-Python
    MinTree / CircularMinTree / CircularSumTree of starr/experimental/slow.py (imported from the reference tree)."""
    import importlib.util
    X = oracle.ref_exp()
    spec = importlib.util.spec_from_file_location("ref_slow", os.path.join(reference_root, "starr/experimental/slow.py"))
    slow = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(slow)
    out = {}
    for n in (2, 3, 5, 37, 1000, 4096):
        rng = np.random.default_rng(4000 + n)
        a = np.exp(rng.uniform(-6, 6, n))
        t = np.zeros(n)
        oracle.ref().build_sumtree_from_array(a, t)
        out[f"exp_n{n}/array0"], out[f"exp_n{n}/tree0"] = a.copy(), t.copy()
        for b in range(3):
            m = int(rng.integers(1, 3 * n)); nv = m if b == 0 else int(rng.integers(1, m + 1))
            idx = rng.integers(0, n if b < 2 else max(2, n // 8), m).astype(np.int32)
            vals = np.exp(rng.uniform(-6, 6, nv))
            X.update_tree_multi(idx, vals, a, t)
            out[f"exp_n{n}/b{b}/idx"], out[f"exp_n{n}/b{b}/vals"] = idx, vals
            out[f"exp_n{n}/b{b}/array"], out[f"exp_n{n}/b{b}/tree"] = a.copy(), t.copy()
        q = t[1] * rng.random(200) * 1.05
        found = np.ones(200, np.int32)
        X.get_prefix_sum_idx_multi(found, q, a, t)
        out[f"exp_n{n}/queries"], out[f"exp_n{n}/found"] = q, found
    for n in (2, 3, 5, 6, 7, 13, 64, 100, 1000):
        rng = np.random.default_rng(5000 + n)
        mt = slow.MinTree(n)
        for b in range(3):
            m = int(rng.integers(1, 3 * n))
            idx = rng.integers(0, n, m).astype(np.int64)
            vals = np.round(np.exp(rng.uniform(-3, 3, m)), 3)
            for i, v in zip(idx, vals):
                mt[int(i)] = float(v)
            out[f"min_n{n}/b{b}/idx"], out[f"min_n{n}/b{b}/vals"] = idx, vals
            out[f"min_n{n}/b{b}/data"] = np.array(mt._data, dtype=np.float64)
        cm, cs = slow.CircularMinTree(n), slow.CircularSumTree(n)
        for b in range(3):
            vals = np.round(np.exp(rng.uniform(-3, 3, int(rng.integers(1, 2 * n + 2)))), 3)
            for v in vals:
                cm.append(float(v)); cs.append(float(v))
            out[f"ring_n{n}/b{b}/vals"] = vals
            out[f"ring_n{n}/b{b}/min_data"] = np.array(cm._data, dtype=np.float64)
            out[f"ring_n{n}/b{b}/sum_data"] = np.array(cs._data, dtype=np.float64)
    return out


if __name__ == "__main__":
    if oracle.ref() is None or oracle.ref_api() is None:
        sys.exit("reference not available: run oracle/build_ref.sh in the build container")
    np.savez_compressed(os.path.join(HERE, "kernel_cases.npz"), **kernel_cases())
    np.savez_compressed(os.path.join(HERE, "api_cases.npz"), **api_cases())
    np.savez_compressed(os.path.join(HERE, "next_cases.npz"), **next_cases())
    for f in ("kernel_cases.npz", "api_cases.npz", "next_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
