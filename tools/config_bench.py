#!/usr/bin/env python
"""Time every BASELINE.json config on the GPU next to the reference CPU path (one host core).

    python tools/config_bench.py [--quick]  > profiles/<name>.md

Not the graded bench (that is bench.py); this fills the per-config table in profiles/.
GPU numbers: CUDA events on the current stream, after warm-up, median of repeats.
CPU numbers: the compiled reference module (oracle/_ref) if present, else the C oracle port.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from starr_b200 import DeviceSumTree, SumTreeArray  # noqa: E402

oracle.build()
REF = oracle.ref() or oracle.c
KIND = "reference (oracle/_ref)" if oracle.ref() is not None else "C oracle port"


def gpu_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def gpu_time_graph(fn, calls=8, reps=9):
    """Device time per call with the calls replayed as one CUDA graph: no host gaps inside the timed events (a 20 us
    kernel otherwise measures the host's launch latency)."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn() for _ in range(calls)]
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / calls)
    del keep
    return float(np.median(ts))


def host_time(fn, reps=5, warm=1):
    """Host wall clock of a blocking API call (median), microseconds."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))


def gpu_time_fresh(fn, batches, warm=3):
    """Like gpu_time, but every call gets a batch that has not been applied before: re-applying a batch
    makes most differences exactly 0 and the update skips them."""
    for b in batches[:warm]:
        fn(*b)
    torch.cuda.synchronize()
    ts = []
    for b in batches[warm:]:
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(*b); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    return float(np.median(ts))


def cpu_time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e6


def row(name, gpu_us, cpu_us, units=None):
    rate = f"{units / gpu_us * 1e6:.3g}/s" if units else ""
    crate = f"{units / cpu_us * 1e6:.3g}/s" if units else ""
    print(f"| {name} | {gpu_us:.1f} | {rate} | {cpu_us:.1f} | {crate} | {cpu_us / gpu_us:.1f}x |")


def main():
    quick = "--quick" in sys.argv
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    print(f"# Per-config timings (GPU: {torch.cuda.get_device_name(0)}; CPU: {KIND}, 1 core)\n")
    print("| config / operation | GPU us | GPU rate | CPU us | CPU rate | speed-up |")
    print("|---|---:|---:|---:|---:|---:|")

    # ---- config #1: README bench, n=1e6 float64 through the NumPy API -----------------------
    x = SumTreeArray(np.ones(1_000_000))
    for _ in range(20):                      # warm-up: first calls allocate staging and load kernels
        x[-10:] = 2; x.sample(100); x.sum()
    t0 = time.perf_counter()
    for _ in range(200):
        x[-10:] = 2; x.sample(100); x.sum()
    api_us = (time.perf_counter() - t0) / 200 * 1e6
    ref_api = oracle.ref_api()
    if ref_api is not None:
        y = ref_api.SumTreeArray(np.ones(1_000_000))
        t0 = time.perf_counter()
        for _ in range(50):
            y[-10:] = 2; y.sample(100); y.sum()
        row("#1 NumPy API x[-10:]=2; x.sample(100); x.sum(), n=1e6 f64 (host round trips)", api_us, (time.perf_counter() - t0) / 50 * 1e6)
    else:
        print(f"| #1 NumPy API x[-10:]=2; x.sample(100); x.sum(), n=1e6 f64 (host round trips) | {api_us:.1f} | | n/a (reference package not on this box) | | |")

    # ---- configs #2 / #3: device entry point ------------------------------------------------------
    for label, logn, m in (("#2 PER step", 20, 4096), ("#3 large batch", 20 if quick else 27, 1 << 20)):
        n = 1 << logn
        leaves = rng.random(n, dtype=np.float32)
        tree_np = np.zeros(n, np.float32)
        oracle.c.build_sumtree_from_array(leaves, tree_np)
        t = DeviceSumTree(torch.from_numpy(leaves).to(dev))
        t.reserve(m)
        idx = rng.integers(0, n, m).astype(np.int64); val = rng.random(m, dtype=np.float32); u = rng.random(m)
        di, dv, du = (torch.from_numpy(a).to(dev) for a in (idx, val, u))
        out = torch.empty(m, dtype=torch.int64, device=dev)
        g_build = gpu_time(lambda: t.build_(), reps=5)
        arr = leaves.copy()
        c_upd = cpu_time(lambda: REF.update_sumtree(idx.astype(np.intp), val, arr, tree_np))
        q = (tree_np[1] * u).astype(np.float32); o = np.zeros(m, np.intp)
        c_smp = cpu_time(lambda: REF.get_prefix_sum_idx(o, q, arr, tree_np))
        fresh = [(torch.from_numpy(rng.integers(0, n, m).astype(np.int64)).to(dev),
                  torch.from_numpy(rng.random(m, dtype=np.float32)).to(dev)) for _ in range(23)]
        g_upd = gpu_time_fresh(lambda i, v: t.update_(i, v), fresh)
        g_smp = gpu_time(lambda: t.sample_uniform(du, out=out))
        c_build = cpu_time(lambda: REF.build_sumtree_from_array(arr, tree_np), reps=1) if logn <= 22 else float("nan")
        if m <= 4096:
            # the whole PER step as one CUDA graph (static input buffers)
            st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                t.update_(di, dv); t.sample_uniform(du, out=out, return_priorities=True)
            torch.cuda.current_stream().wait_stream(st)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                t.update_(di, dv); t.sample_uniform(du, out=out, return_priorities=True)

            def replay(i, v):          # fresh inputs into the graph's static buffers (outside the timed events)
                di.copy_(i); dv.copy_(v); torch.cuda.synchronize()
                gr.replay()
            ts = []
            for i, v in fresh:
                di.copy_(i); dv.copy_(v); torch.cuda.synchronize()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); gr.replay(); e.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(e) * 1e3)
            g_graph = float(np.median(ts[3:]))
            row(f"{label}: update+sample {m} as one CUDA graph replay", g_graph, c_upd + c_smp, 2 * m)
        row(f"{label}: build 2^{logn} f32 (validate + pairwise)", g_build, c_build, n)
        row(f"{label}: update {m} @ 2^{logn} f32", g_upd, c_upd, m)
        row(f"{label}: sample {m} @ 2^{logn} f32", g_smp, c_smp, m)
        del t
        if m > 4096:
            # the same configuration through the drop-in NumPy API: HOST arrays in, host results out, blocking calls
            # (st_update_gather_root_host / st_sample_root_host: staging, kernels, read-back and one sync per call)
            x = SumTreeArray(leaves.copy())
            hb = [(rng.integers(0, n, m).astype(np.intp), rng.random(m, dtype=np.float32)) for _ in range(7)]
            k = [0]

            def api_put():
                i, v = hb[k[0] % len(hb)]; k[0] += 1
                x[i] = v
            a_put = host_time(api_put)
            a_smp = host_time(lambda: x.sample(m))
            row(f"{label}: NumPy API x[idx] = vals, {m} host indices/values (blocking)", a_put, c_upd, m)
            row(f"{label}: NumPy API x.sample({m}) -> host indices (blocking)", a_smp, c_smp, m)
            del x

    # ---- config #5: 4096 x 4096 -------------------------------------------------------------------------
    side = 1024 if quick else 4096
    for dt, tdt in ((np.float32, torch.float32), (np.int32, torch.int32)):
        a = rng.random((side, side)).astype(dt) if dt == np.float32 else rng.integers(0, 1000, (side, side)).astype(dt)
        t = DeviceSumTree(torch.from_numpy(a).to(dev))
        flat = a.ravel().copy(); tr = np.zeros(flat.size, dt)
        oracle.c.build_sumtree_from_array(flat, tr)
        m = 1 << 20
        u = rng.random(m); du = torch.from_numpy(u).to(dev); out = torch.empty(m, dtype=torch.int64, device=dev)
        nm = np.dtype(dt).name
        row(f"#5 {side}^2 {nm}: sum(axis=1) (aligned tree read)", gpu_time_graph(lambda: t.strided_sum(side)),
            cpu_time(lambda: REF.strided_sum(flat, tr, side)), side)
        row(f"#5 {side}^2 {nm}: sum(axis=0) (row-order fold, TMA strip kernel)", gpu_time_graph(lambda: t.axis_fold(1, side, side)),
            cpu_time(lambda: a.sum(axis=0)), side * side)
        q = (tr[1] * u).astype(dt); o = np.zeros(m, np.intp)
        row(f"#5 {side}^2 {nm}: sample(2^20) flat indices", gpu_time(lambda: t.sample_uniform(du, out=out)),
            cpu_time(lambda: REF.get_prefix_sum_idx(o, q, flat, tr)), m)
        del t
        x2 = SumTreeArray(a.copy())
        row(f"#5 {side}^2 {nm}: NumPy API x.sum(axis=0) (blocking, host result)", host_time(lambda: x2.sum(axis=0)),
            cpu_time(lambda: a.sum(axis=0)), side * side)
        row(f"#5 {side}^2 {nm}: NumPy API x.sum(axis=1) (blocking, host result)", host_time(lambda: x2.sum(axis=1)),
            cpu_time(lambda: REF.strided_sum(flat, tr, side)), side)
        row(f"#5 {side}^2 {nm}: NumPy API x.sample(2^20) -> (rows, cols) (blocking)", host_time(lambda: x2.sample(m)),
            cpu_time(lambda: REF.get_prefix_sum_idx(o, q, flat, tr)), m)
        del x2


if __name__ == "__main__":
    main()
