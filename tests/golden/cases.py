"""Seeded case definitions shared by make_golden.py (run against the reference) and the
parity tests (run against starr_b200): the same code drives both, so operands are identical."""
from __future__ import annotations

import numpy as np


def priorities(rng, n, dtype, kind):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        if kind == "uniform":
            return rng.random(n).astype(dtype)
        if kind == "loguniform":
            return np.exp(rng.uniform(-8, 8, n)).astype(dtype)
        return np.zeros(n, dtype)
    hi = {1: 4, 2: 60}.get(dtype.itemsize, 1000)
    return rng.integers(0, hi, n).astype(dtype)


KERNEL_SPECS = [
    ("f32_n5", np.float32, 5, "uniform"),
    ("f32_n37", np.float32, 37, "loguniform"),
    ("f32_n1000", np.float32, 1000, "uniform"),
    ("f32_n4096", np.float32, 4096, "loguniform"),
    ("f32_zero_n1024", np.float32, 1024, "zero"),
    ("f64_n1000", np.float64, 1000, "loguniform"),
    ("f64_n2048", np.float64, 2048, "uniform"),
    ("i32_n1000", np.int32, 1000, "uniform"),
    ("i32_n4096", np.int32, 4096, "uniform"),
    ("i64_n257", np.int64, 257, "uniform"),
    ("i16_n100", np.int16, 100, "uniform"),
    ("u8_n64", np.uint8, 64, "uniform"),
    ("u16_n129", np.uint16, 129, "uniform"),
    ("u32_n1000", np.uint32, 1000, "uniform"),
    ("u64_n513", np.uint64, 513, "uniform"),
]
KERNEL_STRIDES = (1, 2, 3, 7, 8, 64)
KERNEL_BATCHES = 4

API_SPECS = (("api_f32_2d", np.float32, (48, 64)), ("api_i32_2d", np.int32, (48, 64)),
             ("api_f64_3d", np.float64, (6, 10, 16)), ("api_f32_1d", np.float32, (1000,)))


def api_script(SumTreeArray, name, dt, shape, seed=77):
    """Run one API case on the given SumTreeArray class; returns {key: ndarray}."""
    rng = np.random.default_rng(seed + len(name) + int(np.prod(shape)))
    out = {}
    base = priorities(rng, int(np.prod(shape)), dt, "uniform").reshape(shape)
    x = SumTreeArray(base)
    out["base"] = base
    if len(shape) == 2:
        x[3:9, 5:40] = priorities(rng, 7, dt, "uniform")          # cycling values (not broadcast)
        x[np.array([0, 1, 1, 47]), np.array([1, 0, 0, 63])] = priorities(rng, 4, dt, "uniform")
        mask = rng.random(shape[0]) < 0.2
        x[mask] = priorities(rng, 1, dt, "uniform")[0]
    elif len(shape) == 3:
        x[1:3] = priorities(rng, 5, dt, "uniform")
    else:
        x[-10:] = 2
        x[rng.integers(0, shape[0], 300)] = priorities(rng, 300, dt, "uniform")
    out["after"] = np.array(x.view(np.ndarray))
    out["sumtree"] = x.sumtree()
    out["sum"] = np.array(x.sum())
    for ax in range(len(shape)):
        out[f"sum_ax{ax}"] = x.sum(axis=ax)
        out[f"sum_ax{ax}_keep"] = x.sum(axis=ax, keepdims=True)
    if len(shape) == 3:
        out["sum_ax12"] = x.sum(axis=(1, 2))
        out["sum_ax01"] = x.sum(axis=(0, 1))
    np.random.seed(1234)
    out["sample"] = np.array(x.sample(257))
    np.random.seed(1234)
    out["sample_flat"] = x.sample(257, flatten_indices=True)
    out["prefix_ids"] = np.array(x.get_prefix_sum_id(np.linspace(0, float(x.sum()) * 1.05, 40).reshape(5, 8)))
    x += 1
    out["sumtree_iadd"] = x.sumtree()
    x.fill(3)
    out["sumtree_fill"] = x.sumtree()
    return out
