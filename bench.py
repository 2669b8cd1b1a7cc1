#!/usr/bin/env python
"""bench.py -- the SumTreeArray hot path on B200: leaf updates + prefix-search samples per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input:
2^20 priority updates followed by 2^20 prefix-search samples on a float32 tree with 2^27 leaves
per GPU (BASELINE.json config #3; at N > 1 the leaf range is sharded by subtree, 2^27 leaves and
2^20 + 2^20 operations per rank, i.e. config #4's 2^30 leaves at N = 8 -- weak scaling).

Prints ONE JSON line (rank 0).  `value` = (updates + samples) of all ranks per second with the
inputs already resident in HBM, timed with CUDA events around exactly K steps (barrier +
synchronize on both sides, max over ranks).  `e2e` = the same through the public API with pinned
HOST buffers (H2D of indices/values/uniforms and D2H of the sampled indices inside the timed
region).  `roofline` = algorithmic bytes of the dominant kernel / its measured duration against
the measured HBM copy bandwidth (MEASURED_PEAKS.json).  `cpu_baseline` = the reference's own
compiled Cython module (oracle/_ref) -- or the C oracle port if that is absent -- timed on one
host core on the same inputs (the reference is single-threaded by construction).

`--impl reference` times that CPU implementation as the whole arm (same metric / config; with --gpus N one tree over
all N * 2^27 leaves, each step a bounded 2^20 + 2^20 sample of the sharded step, rank 0 only).
The oracle is only ever the checker / the baseline here, never part of the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the sharded step keeps wait kernels spinning on up to three streams: one hardware queue per stream
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

LOG2_LEAVES = int(os.environ.get("ST_BENCH_LOG2_LEAVES", "27"))   # per GPU
LOG2_BATCH = int(os.environ.get("ST_BENCH_LOG2_BATCH", "20"))     # per GPU
METRIC = "leaf updates + prefix-search samples per second (2^27 float32 leaves per GPU, 2^20 + 2^20 per step)"
UNIT = "ops/s"


def depth_of(n):  # levels above the leaves
    return int(n - 1).bit_length()


def bytes_per_update(n, s=4):   # idx + val + RMW of leaf and D ancestors
    return 8 + s + 2 * s * (depth_of(n) + 1)


def bytes_per_sample(n, s=4):   # u + idx out + one left-child read per level
    return 8 + 8 + s * depth_of(n)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return {k: v["dram_bytes_per_launch"] for k, v in json.load(f)["kernels"].items()}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, windows=()):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for ts, r in self.rows if any(a - 0.03 <= ts <= b + 0.03 for a, b in windows)]
        scope = "timed regions" if inside else "whole run (timed regions shorter than the sampling period)"
        for r in (inside or [r for _, r in self.rows]):
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope, "period_ms": 20}


def make_batches(rng, n_leaves_global, m_global, count):
    """SURVEY 8(d) synthetic inputs: idx with replacement, uniform priorities, float64 uniforms.
    `count` DISTINCT batches: re-applying a batch makes most differences exactly 0, which the update
    skips -- every timed step therefore gets a batch that has not been applied before."""
    out = []
    for _ in range(count):
        out.append((rng.integers(0, n_leaves_global, m_global).astype(np.int64),
                    rng.random(m_global, dtype=np.float32),
                    rng.random(m_global)))
    return out


# --------------------------------------------------------------------------------------------
# CPU baseline (reference module from oracle/_ref, else the C oracle port)
# --------------------------------------------------------------------------------------------
def cpu_arm(leaves, batches, steps, warmup, keep=False, full_step=None):
    """Time the reference CPU path: per step update_sumtree(2^20) + get_prefix_sum_idx(2^20), every step on
    a batch it has not seen.  keep=True also returns the final (leaves, tree) and every step's sampled
    indices: the in-run parity gate replays the same batches on the GPU and compares bits."""
    import oracle
    oracle.build()
    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    K = ref if ref is not None else oracle.c
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    saved_affinity = None
    try:
        saved_affinity = os.sched_getaffinity(0)
        os.sched_setaffinity(0, {sorted(saved_affinity)[0]})
    except Exception:
        pass
    n = leaves.size
    arr = leaves.copy()
    tree = np.zeros(n, np.float32)
    oracle.c.build_sumtree_from_array(arr, tree)   # bit-identical to the reference build, 10x faster
    m = batches[0][0].size
    assert len(batches) >= warmup + steps
    t_upd = t_srch = 0.0
    sampled = []
    for it in range(warmup + steps):
        idx, val, u = batches[it]
        idx = idx.astype(np.intp)
        t0 = time.perf_counter()
        K.update_sumtree(idx, val, arr, tree)
        t1 = time.perf_counter()
        q = (tree[1] * u).astype(np.float32)
        out = np.zeros(m, np.intp)
        K.get_prefix_sum_idx(out, q, arr, tree)
        t2 = time.perf_counter()
        if keep:
            sampled.append(out)
        if it >= warmup:
            t_upd += t1 - t0
            t_srch += t2 - t1
    total = t_upd + t_srch
    res = {
        "value": 2 * m * steps / total, "unit": UNIT, "cores": 1, "kind": kind,
        "sample": (f"{steps} full steps (2^{LOG2_BATCH} updates + 2^{LOG2_BATCH} samples each, a fresh batch per step) "
                   f"on one pinned host core; tree built by the C oracle") if not full_step or full_step == m else
                  (f"{steps} bounded steps on the {n}-leaf tree: {m} updates + {m} samples of each step's {full_step} + "
                   f"{full_step} (a fresh batch per step) on one pinned host core; ms_per_step is scaled to the full step; "
                   f"tree built by the C oracle"),
        "updates_per_s": m * steps / t_upd, "samples_per_s": m * steps / t_srch,
        "host_cores_available": avail,
        "ms_per_step": 1e3 * total / steps * ((full_step / m) if full_step else 1.0),
    }
    if keep:
        res["_state"] = (arr, tree, sampled)
    try:
        if saved_affinity:
            os.sched_setaffinity(0, saved_affinity)
    except Exception:
        pass
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the arm's config is ours: at N > 1 ONE tree over all N * 2^27 leaves (what the N shards hold together), each step a
    # bounded sample (2^20 + 2^20 operations) of the N * 2^20 + N * 2^20 a sharded step processes
    world = max(1, int(args.gpus))
    n = (1 << LOG2_LEAVES) * world
    m = 1 << LOG2_BATCH
    rng = np.random.default_rng(0)
    leaves = rng.random(n, dtype=np.float32)
    steps = max(1, min(args.steps, 8))
    warmup = max(1, min(args.warmup, 2))
    batches = make_batches(rng, n, m, steps + warmup)
    base = cpu_arm(leaves, batches, steps, warmup, full_step=m * world)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, "single" if world == 1 else args.layout),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "updates_per_s",
                                              "samples_per_s", "host_cores_available")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, layout="single"):
    how = {"single": "single tree",
           "slices": f"subtree shards x{world}; the batch is distributed (every rank supplies and samples its own 1/{world} "
                     "of it); exchange over CUDA-IPC peer memory (NVLink), NCCL only for set-up",
           "replicated": f"subtree shards x{world}; the whole batch is replicated on every rank; exchange of deltas / roots / "
                         "results over peer memory"}[layout]
    return {
        "workload": f"SumTreeArray float32, 2^{LOG2_LEAVES} leaves per GPU ({world} GPU: 2^{LOG2_LEAVES + (world - 1).bit_length()} "
                    f"leaves{' sharded by subtree' if world > 1 else ''}), per step 2^{LOG2_BATCH} leaf updates + 2^{LOG2_BATCH} prefix-search samples per GPU "
                    f"(BASELINE config #3; config #4 at 8 GPUs)",
        "leaves_per_gpu": 1 << LOG2_LEAVES, "updates_per_step_per_gpu": 1 << LOG2_BATCH,
        "samples_per_step_per_gpu": 1 << LOG2_BATCH,
        "l2_policy": "inputs larger than L2: the tree is 1 GiB per GPU; every warm-up and timed step applies a batch "
                     "(indices, values, uniforms) that has never been applied before (zero_delta_frac reports the share "
                     "of updates whose difference was exactly 0)",
        "parallelism": how,
    }


# --------------------------------------------------------------------------------------------
# parity gates (SURVEY 8(d) "parity gates in the same run")
# --------------------------------------------------------------------------------------------
def parity_single(torch, dev, leaves_host, batches_host, cpu_state):
    """Replay the CPU arm's batches on a fresh device tree: sumtree() bits, leaf bits and every step's
    sampled indices must equal the reference's."""
    from starr_b200 import DeviceSumTree
    arr, tree, sampled = cpu_state
    t = DeviceSumTree(torch.from_numpy(leaves_host).to(dev))
    ok_idx = True
    for k, want in enumerate(sampled):
        i, v, u = batches_host[k]
        t.update_(torch.from_numpy(i).to(dev), torch.from_numpy(v).to(dev))
        got = t.sample_uniform(torch.from_numpy(u).to(dev))
        ok_idx = ok_idx and bool(np.array_equal(got.cpu().numpy(), want))
    t.poll()
    ok_tree = t.tree.cpu().numpy().tobytes() == tree.tobytes()
    ok_leaves = t.leaves.cpu().numpy().tobytes() == arr.tobytes()
    del t
    return {"checked": True, "batches": len(sampled), "sumtree_bits_equal": bool(ok_tree), "leaves_bits_equal": bool(ok_leaves),
            "sampled_indices_equal": bool(ok_idx), "ok": bool(ok_tree and ok_leaves and ok_idx)}


def parity_sharded(torch, dist, dev, world, rank, layout):
    """Reduced sharded case against the C oracle on EVERY rank, before the timed region: 2^22 leaves over the
    ranks, float32 and int32, 3 batches of 2^16 updates (with duplicates) and 50 000 samples, in the layout the
    timed region uses (slices: every rank supplies 1/N of the batch; replicated: every rank holds all of it)."""
    import oracle
    from starr_b200 import ShardedSumTree
    oracle.build()
    C = oracle.c
    n, m, q = 1 << 22, 1 << 16, 50000 // world * world
    n_local = n // world
    detail = {}
    ok = True
    for name, dt, tdt in (("float32", np.float32, torch.float32), ("int32", np.int32, torch.int32)):
        rng = np.random.default_rng(77)
        leaves = rng.random(n).astype(dt) if dt == np.float32 else rng.integers(0, 1000, n).astype(dt)
        tree = np.zeros(n, dt)
        C.build_sumtree_from_array(leaves, tree)
        sh = ShardedSumTree(n, tdt, local_leaves=torch.from_numpy(leaves[rank * n_local:(rank + 1) * n_local].copy()).to(dev))
        good = True
        for b in range(3):
            idx = rng.integers(0, n if b < 2 else n // 64, m).astype(np.int64)
            val = rng.random(m).astype(dt) if dt == np.float32 else rng.integers(0, 1000, m).astype(dt)
            u = rng.random(q)
            C.update_sumtree(idx.astype(np.intp), val, leaves, tree)
            want = np.zeros(q, np.intp)
            C.get_prefix_sum_idx(want, (tree[1] * u).astype(dt), leaves, tree)
            if layout == "slices":      # every rank supplies its 1/N of the same batch
                ms, qs = m // world, q // world
                sh.update_slices_(torch.from_numpy(idx[rank * ms:(rank + 1) * ms]).to(dev),
                                  torch.from_numpy(val[rank * ms:(rank + 1) * ms]).to(dev))
                got = sh.sample_slices(torch.from_numpy(u[rank * qs:(rank + 1) * qs]).to(dev))
                good = good and bool(np.array_equal(got.cpu().numpy(), want[rank * qs:(rank + 1) * qs]))
            else:
                sh.update_(torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev))
                got = sh.sample_uniform(torch.from_numpy(u).to(dev))
                good = good and bool(np.array_equal(got.cpu().numpy(), want))
        sh.poll()
        good = good and sh.gather_sumtree().cpu().numpy().tobytes() == tree.tobytes()
        good = good and sh.gather_leaves().cpu().numpy().tobytes() == leaves.tobytes()
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        detail[name] = bool(flag.item())
        ok = ok and detail[name]
        sh.close()
        del sh
    torch.cuda.empty_cache()
    return ok, {"leaves": n, "batches": 3, "updates_per_batch": m, "samples_per_batch": q, "checked_on": "every rank",
                "layout": layout, **detail}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
MAX_DISTINCT_BYTES = 12 << 30    # device memory spent on pre-generated batches before cycling (with fresh values)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from starr_b200 import DeviceSumTree, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the single JSON line: whatever NCCL_DEBUG level the box exports goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    layout = args.layout if world > 1 else "single"
    n_local = 1 << LOG2_LEAVES
    m_local = 1 << LOG2_BATCH
    n_global = n_local * world
    m_global = m_local * world
    steps, warmup = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- sharded parity gate, before anything is timed -------------------------------------------
    par_sharded = None
    if world > 1:
        ok, par_sharded = parity_sharded(torch, dist, dev, world, rank, layout)
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "parity_sharded": False, "parity_sharded_detail": par_sharded}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)

    # synthetic priorities: every rank draws its own shard
    leaves_host = np.random.default_rng(1000 + rank).random(n_local, dtype=np.float32) if world > 1 else \
        np.random.default_rng(0).random(n_local, dtype=np.float32)

    def build_tree(n_loc, leaves_dev, m_res):
        if world == 1:
            t = DeviceSumTree(leaves_dev)
            t.reserve(m_res)
        else:
            from starr_b200 import ShardedSumTree
            t = ShardedSumTree(n_loc * world, torch.float32, local_leaves=leaves_dev)
            if layout == "slices":
                t.reserve_slices(m_res // world, expected_local=m_res // world * 5 // 4)
            else:
                t.reserve(m_res)
        return t

    tree = build_tree(n_local, torch.from_numpy(leaves_host).to(dev), m_global)

    # Batches are generated on the device with the same seed on every rank (replicated input, as a PER
    # learner broadcasts it): DISTINCT batches for every warm-up and timed step.
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260 + (rank if layout == "slices" else 0))

    def gen_batch(n_tot, m_tot):
        if layout == "slices":          # this rank's 1/N of the global batch (indices over the whole leaf range)
            m_tot //= world
        return (torch.randint(0, n_tot, (m_tot,), generator=gen, device=dev, dtype=torch.int64),
                torch.rand(m_tot, generator=gen, device=dev, dtype=torch.float32),
                torch.rand(m_tot, generator=gen, device=dev, dtype=torch.float64))

    def counters(t):
        h = t.handle if world == 1 else t.local.handle
        return h.update_counters(torch.cuda.current_stream().cuda_stream, reset=True)

    def timed_leg(t, n_tot, m_tot, nsteps, nwarm):
        """W warm-up + exactly K timed steps on never-applied batches; returns per-rank times (ms)."""
        m_mine = m_tot // world if layout == "slices" else m_tot
        per = m_mine * 20
        distinct = max(2, min(nwarm + nsteps + 1, MAX_DISTINCT_BYTES // per))
        cycling = distinct < nwarm + nsteps + 1
        bat = [gen_batch(n_tot, m_tot) for _ in range(distinct)]
        outb = torch.empty(m_mine, dtype=torch.int64, device=dev)
        s_plan = torch.cuda.Stream() if world > 1 else None
        if layout == "slices":
            upd, smp, mkplan = t.update_slices_, t.sample_slices, t.plan_update_slices
        elif world > 1:
            upd, smp, mkplan = t.update_, t.sample_uniform, t.plan_update
        else:
            upd, smp, mkplan = t.update_, t.sample_uniform, None

        def use(k):
            b = bat[k % distinct]
            return b

        def refresh(k):
            if cycling:   # the slot will be reused: give it values it has never had (a few us, inside the region)
                bat[k % distinct][1].uniform_(0, 1, generator=gen)

        for w in range(nwarm):
            i, v, u = use(w)
            upd(i, v)
            smp(u, out=outb)
            refresh(w)
        t.poll()
        barrier()
        counters(t)
        launches0 = int(_lib.lib().st_kernel_launches())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * nsteps + 1)]
        # Sharded runs plan ahead: owner / slot of batch k+1 and (slices layout) the exchange of its slices -- none of
        # which reads the tree -- run on a side stream while batch k is applied.  One plan per step inside the timed region.
        plan = mkplan(*use(nwarm)[:2], stream=s_plan) if world > 1 else None
        barrier()
        t_w0 = time.time()
        ev[0].record()
        for k in range(nsteps):
            i, v, u = use(nwarm + k)
            if world > 1:
                upd(plan=plan)
                ev[2 * k + 1].record()
                if k + 1 < nsteps:
                    plan = mkplan(*use(nwarm + k + 1)[:2], stream=s_plan)
            else:
                upd(i, v)
                ev[2 * k + 1].record()
            smp(u, out=outb)
            ev[2 * k + 2].record()
            refresh(nwarm + k)
        barrier()
        window = (t_w0, time.time())
        launches = int(_lib.lib().st_kernel_launches()) - launches0
        total_ms = ev[0].elapsed_time(ev[-1])
        upd_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(nsteps))
        smp_ms = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(nsteps))
        t.poll()
        cnt = counters(t)
        del bat
        return total_ms, upd_ms, smp_ms, launches, window, cnt, cycling

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    windows = []
    total_ms, upd_ms, smp_ms, launches, win, cnt_dev, cycling = timed_leg(tree, n_global, m_global, steps, warmup)
    windows.append(win)

    # ---- end-to-end: HOST buffers in, sampled indices back on the host -----------------------------
    rng_e = np.random.default_rng(555 + rank)
    e2e_distinct = min(steps + 3, 64)
    if world == 1:
        # through the C ABI: st_per_step_host_submit / st_per_step_host_wait (pinned host buffers from
        # st_host_alloc; three slots in flight, so the H2D of step k+1 and the D2H of step k-1 overlap the kernels of step k
        # and the copy stream never waits for the host)
        h = tree.handle
        pins = []
        for _ in range(e2e_distinct):
            pb = _lib.PinnedBuffer(m_local * 20)
            i = pb.array(np.int64, m_local, 0); v = pb.array(np.float32, m_local, 8 * m_local)
            u = pb.array(np.float64, m_local, 12 * m_local)
            i[:] = rng_e.integers(0, n_local, m_local); v[:] = rng_e.random(m_local, dtype=np.float32); u[:] = rng_e.random(m_local)
            pins.append((pb, i, v, u))
        NS = 3          # slots in flight (the library has ST_HOST_SLOTS = 4): step k is submitted before step k - 2 is collected
        outs = []
        for _ in range(NS):
            pb = _lib.PinnedBuffer(m_local * 8)
            outs.append((pb, pb.array(np.int64, m_local, 0)))

        def e2e_run(first, nsteps):
            for k in range(nsteps):
                slot = k % NS
                if k >= NS:
                    h.per_step_host_wait(slot)       # its results are in outs[slot]; the slot is reused for step k
                _, i, v, u = pins[(first + k) % e2e_distinct]
                h.per_step_host_submit(slot, i, v, u, outs[slot][1])
            for k in range(max(0, nsteps - NS), nsteps):
                h.per_step_host_wait(k % NS)

        torch.cuda.synchronize()
        e2e_run(0, NS)
        counters(tree)
        t_w0 = time.time()
        t0 = time.perf_counter()
        e2e_run(NS, steps)
        e2e_ms = (time.perf_counter() - t0) * 1e3     # host clock: the call is synchronous at its ends (wait() blocks)
        windows.append((t_w0, time.time()))
        cnt_e2e = counters(tree)
        e2e_how = ("C ABI st_per_step_host_submit/_wait on pinned host buffers (st_host_alloc): every step's indices, "
                   "values and uniforms are copied H2D and its sampled indices D2H inside the timed region; three slots in "
                   "flight, so copies of step k+1 / k-1 overlap the kernels of step k; timed with the host clock around "
                   "the blocking calls (the last wait returns when the last step's indices are on the host)")
        del pins, outs
    elif layout == "slices":
        # Sharded, slices layout: every rank holds ITS 1/N of the batch in pinned host memory, uploads it (PCIe carries
        # every byte once), hands it to update_slices_ / sample_slices as it is, and reads back its own indices.
        pinned = [(torch.from_numpy(rng_e.integers(0, n_global, m_local).astype(np.int64)).pin_memory(),
                   torch.from_numpy(rng_e.random(m_local, dtype=np.float32)).pin_memory(),
                   torch.from_numpy(rng_e.random(m_local)).pin_memory()) for _ in range(e2e_distinct)]
        host_out = torch.empty(m_local, dtype=torch.int64).pin_memory()
        s_main = torch.cuda.current_stream()
        s_in, s_out, s_plan = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        stage = [tuple(torch.empty_like(t, device=dev) for t in pinned[0]) for _ in range(2)]
        obuf = [torch.empty(m_local, dtype=torch.int64, device=dev) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_read = [torch.cuda.Event() for _ in range(2)]

        def prefetch(first, k):
            slot = k % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[slot])
                for dst, src in zip(stage[slot], pinned[(first + k) % e2e_distinct]):
                    dst.copy_(src, non_blocking=True)
                ev_in[slot].record(s_in)

        def e2e_run(first, nsteps):
            for slot in range(2):
                ev_free[slot].record(s_main)
                ev_read[slot].record(s_main)
            prefetch(first, 0)
            s_plan.wait_event(ev_in[0])
            plan = tree.plan_update_slices(*stage[0][:2], stream=s_plan)
            for k in range(nsteps):
                slot = k % 2
                if k + 1 < nsteps:
                    prefetch(first, k + 1)
                s_main.wait_event(ev_in[slot])
                s_main.wait_event(ev_read[slot])
                tree.update_slices_(plan=plan)
                if k + 1 < nsteps:
                    s_plan.wait_event(ev_in[(k + 1) % 2])
                    plan = tree.plan_update_slices(*stage[(k + 1) % 2][:2], stream=s_plan)
                tree.sample_slices(stage[slot][2], out=obuf[slot])
                ev_free[slot].record(s_main)
                ev_done[slot].record(s_main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[slot])
                    host_out.copy_(obuf[slot], non_blocking=True)
                    ev_read[slot].record(s_out)
            s_main.wait_stream(s_out)

        e2e_run(0, 2)
        barrier()
        counters(tree)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        e0.record()
        e2e_run(2, steps)
        e1.record()
        barrier()
        windows.append((t_w0, time.time()))
        e2e_ms = e0.elapsed_time(e1)
        cnt_e2e = counters(tree)
        e2e_how = ("pinned host buffers; every rank uploads its 1/N of the batch, passes it to update_slices_ / sample_slices and "
                   "reads back its own sampled indices (byte counts are totals over ranks); H2D of step k+1 and D2H of step k-1 "
                   "overlap step k")
        del pinned, stage, obuf
    else:
        # Sharded: every rank holds ITS 1/N slice of the batch in pinned host memory, uploads only that (PCIe
        # carries every byte once), the slices are all-gathered over NVLink, and every rank reads back its slice.
        mine = slice(rank * m_local, (rank + 1) * m_local)
        pinned = [(torch.from_numpy(rng_e.integers(0, n_global, m_local).astype(np.int64)).pin_memory(),
                   torch.from_numpy(rng_e.random(m_local, dtype=np.float32)).pin_memory(),
                   torch.from_numpy(rng_e.random(m_local)).pin_memory()) for _ in range(e2e_distinct)]
        host_out = torch.empty(m_local, dtype=torch.int64).pin_memory()
        s_main = torch.cuda.current_stream()
        s_in, s_out, s_plan = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        dbuf = [tuple(torch.empty(m_global, dtype=t.dtype, device=dev) for t in pinned[0]) for _ in range(2)]
        stage = [tuple(torch.empty_like(t, device=dev) for t in pinned[0]) for _ in range(2)]
        obuf = [torch.empty(m_global, dtype=torch.int64, device=dev) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_read = [torch.cuda.Event() for _ in range(2)]

        def prefetch(first, k):
            slot = k % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[slot])
                for dst, src in zip(stage[slot], pinned[(first + k) % e2e_distinct]):
                    dst.copy_(src, non_blocking=True)
                for full, part in zip(dbuf[slot], stage[slot]):
                    dist.all_gather_into_tensor(full, part)
                ev_in[slot].record(s_in)

        def e2e_run(first, nsteps):
            for slot in range(2):
                ev_free[slot].record(s_main)
                ev_read[slot].record(s_main)
            prefetch(first, 0)
            s_plan.wait_event(ev_in[0])
            plan = tree.plan_update(*dbuf[0][:2], stream=s_plan)
            for k in range(nsteps):
                slot = k % 2
                if k + 1 < nsteps:
                    prefetch(first, k + 1)
                s_main.wait_event(ev_in[slot])
                s_main.wait_event(ev_read[slot])
                tree.update_(plan=plan)
                if k + 1 < nsteps:
                    s_plan.wait_event(ev_in[(k + 1) % 2])
                    plan = tree.plan_update(*dbuf[(k + 1) % 2][:2], stream=s_plan)
                tree.sample_uniform(dbuf[slot][2], out=obuf[slot])
                ev_free[slot].record(s_main)
                ev_done[slot].record(s_main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[slot])
                    host_out.copy_(obuf[slot][mine], non_blocking=True)
                    ev_read[slot].record(s_out)
            s_main.wait_stream(s_out)

        e2e_run(0, 2)
        barrier()
        counters(tree)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        e0.record()
        e2e_run(2, steps)
        e1.record()
        barrier()
        windows.append((t_w0, time.time()))
        e2e_ms = e0.elapsed_time(e1)
        cnt_e2e = counters(tree)
        e2e_how = ("pinned host buffers; every rank uploads / reads back its 1/N slice of the batch (byte counts are totals "
                   "over ranks), slices are all-gathered over NVLink; H2D of step k+1 and D2H of step k-1 overlap step k")
        del pinned, dbuf, stage, obuf
    tree.poll()
    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- per-kernel durations of the update pipeline (CUDA events on the launching stream) -------
    breakdown = None
    if world == 1:
        nb = min(steps, 10)
        bat = [gen_batch(n_global, m_global) for _ in range(nb)]
        tree.handle.set_profiling(True)
        for i, v, u in bat:
            tree.update_(i, v)
        torch.cuda.synchronize()
        breakdown = tree.handle.update_breakdown()
        tree.handle.set_profiling(False)
        del bat

    # ---- BASELINE config #4 as stated: a FIXED 2^30 leaves over the N GPUs, 2^20 + 2^20 per step -------------
    fixed = None
    if world > 1 and not args.no_fixed:
        tree.close()        # collective: the exchange areas are released on every rank together
        del tree
        torch.cuda.empty_cache()
        nl = (1 << 30) // world
        gen_l = torch.Generator(device=dev)
        gen_l.manual_seed(4242 + rank)              # every rank its own shard; `gen` stays in lock step for the batches
        lv = torch.rand(nl, device=dev, dtype=torch.float32, generator=gen_l)
        tree = build_tree(nl, lv, 1 << LOG2_BATCH)
        del lv
        fsteps = max(1, min(steps, 20))
        f_total, f_upd, f_smp, _, _, f_cnt, _ = timed_leg(tree, 1 << 30, 1 << LOG2_BATCH, fsteps, 3)
        ft = torch.tensor([f_total, f_upd, f_smp], dtype=torch.float64, device=dev)
        dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        f_total, f_upd, f_smp = ft.tolist()
        mm = 1 << LOG2_BATCH
        fixed = {"workload": f"BASELINE config #4: 2^30 float32 leaves sharded over {world} GPUs, 2^{LOG2_BATCH} updates + "
                             f"2^{LOG2_BATCH} samples per step in total (strong scaling)",
                 "steps": fsteps, "ms_per_step": f_total / fsteps, "value": 2 * mm * fsteps / (f_total * 1e-3), "unit": UNIT,
                 "updates_per_s": mm * fsteps / (f_upd * 1e-3), "samples_per_s": mm * fsteps / (f_smp * 1e-3),
                 "zero_delta_frac": f_cnt["zero_delta"] / max(1, f_cnt["updates"])}

    times = torch.tensor([total_ms, upd_ms, smp_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, upd_ms, smp_ms, e2e_ms = times.tolist()

    rc = 0
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        ops_per_step = 2 * m_global
        upd_rate = m_global * steps / (upd_ms * 1e-3)
        smp_rate = m_global * steps / (smp_ms * 1e-3)
        bu, bs = bytes_per_update(n_global), bytes_per_sample(n_global)

        def roof(name, per_unit, units_per_launch, ms_per_launch, traffic=None):
            ach = per_unit * units_per_launch / (ms_per_launch * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "algorithmic_bytes_per_unit": per_unit, "peak_source": peak_src,
                    "ms_per_launch": ms_per_launch}

        traffic = ncu_traffic()
        r_upd = roof("update pipeline, all kernels of one batch (leaf pass + ordered partition + retirement + folds); "
                     "236 B/update is the work of the whole pipeline, so it is charged to the whole pipeline",
                     bu, m_local, upd_ms / steps, traffic.get("update_pipeline"))
        r_smp = roof("search_kernel<float, staged, uniform>", bs, m_local, smp_ms / steps, traffic.get("search_kernel"))
        r_step = roof("whole step (update pipeline + search_kernel)", bu + bs, m_local, total_ms / steps)
        dominant = dict(r_upd if upd_ms >= smp_ms else r_smp)
        dominant["share_of_step"] = max(upd_ms, smp_ms) / total_ms
        line = {
            "metric": METRIC, "value": ops_per_step * steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, layout),
            "updates_per_s": upd_rate, "samples_per_s": smp_rate,
            "update_ms_per_step": upd_ms / steps, "sample_ms_per_step": smp_ms / steps,
            "zero_delta_frac": cnt_dev["zero_delta"] / max(1, cnt_dev["updates"]),
            "fresh_batches": {"distinct_per_leg": "all" if not cycling else "cycled with regenerated values",
                              "updates_counted": cnt_dev["updates"]},
            "roofline": dominant, "roofline_update": r_upd, "roofline_sample": r_smp, "roofline_step": r_step,
            "e2e": {"value": ops_per_step * steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": m_global * (8 + 4 + 8), "d2h_bytes_per_step": m_global * 8,
                    "ms_per_step": e2e_ms / steps, "zero_delta_frac": cnt_e2e["zero_delta"] / max(1, cnt_e2e["updates"]),
                    "distinct_batches": e2e_distinct, "note": e2e_how},
            "gpu_launches": launches, "clocks": clocks, "update_breakdown_ms": breakdown,
        }
        if par_sharded is not None:
            line["parity_sharded"] = True
            line["parity_sharded_detail"] = par_sharded
        if fixed is not None:
            line["fixed_2p30"] = fixed
        if world == 1 and not args.no_cpu_baseline:
            # CPU leg + parity gate: the reference applies 4 fresh batches; the same batches are replayed on a
            # fresh device tree and every bit is compared
            del tree
            torch.cuda.empty_cache()
            rng = np.random.default_rng(0)
            rng.random(n_local, dtype=np.float32)     # leaves_host came from this stream (same as --impl reference)
            bh = make_batches(rng, n_local, m_local, 4)
            cb = cpu_arm(leaves_host, bh, steps=3, warmup=1, keep=True)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "updates_per_s",
                                                       "samples_per_s", "host_cores_available")}
            par = parity_single(torch, dev, leaves_host, bh, cb["_state"])
            par["oracle"] = cb["kind"]
            line["parity"] = par
            if not par["ok"]:
                rc = 4
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg and the in-run parity gate (N=1)")
    ap.add_argument("--no-fixed", action="store_true", help="skip the fixed-2^30-leaf leg (N>1)")
    ap.add_argument("--layout", default=os.environ.get("ST_BENCH_LAYOUT", "slices"), choices=["slices", "replicated"],
                    help="N>1: 'slices' = every rank supplies 1/N of each batch (default), 'replicated' = every rank holds all of it")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
