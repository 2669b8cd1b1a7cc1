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

`--impl reference` times that CPU implementation as the whole arm (same metric / config).
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

LOG2_LEAVES = int(os.environ.get("ST_BENCH_LOG2_LEAVES", "27"))   # per GPU
LOG2_BATCH = int(os.environ.get("ST_BENCH_LOG2_BATCH", "20"))     # per GPU
N_BATCHES = 4                                                      # distinct synthetic batches, cycled
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


def make_batches(rng, n_leaves_global, m_global):
    """SURVEY 8(d) synthetic inputs: idx with replacement, uniform priorities, float64 uniforms."""
    out = []
    for _ in range(N_BATCHES):
        out.append((rng.integers(0, n_leaves_global, m_global).astype(np.int64),
                    rng.random(m_global, dtype=np.float32),
                    rng.random(m_global)))
    return out


# --------------------------------------------------------------------------------------------
# CPU baseline (reference module from oracle/_ref, else the C oracle port)
# --------------------------------------------------------------------------------------------
def cpu_arm(leaves, batches, steps, warmup):
    """Time the reference CPU path: per step update_sumtree(2^20) + get_prefix_sum_idx(2^20)."""
    import oracle
    oracle.build()
    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    K = ref if ref is not None else oracle.c
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    n = leaves.size
    arr = leaves.copy()
    tree = np.zeros(n, np.float32)
    oracle.c.build_sumtree_from_array(arr, tree)   # bit-identical to the reference build, 10x faster
    m = batches[0][0].size
    t_upd = t_srch = 0.0
    for it in range(warmup + steps):
        idx, val, u = batches[it % len(batches)]
        idx = idx.astype(np.intp)
        t0 = time.perf_counter()
        K.update_sumtree(idx, val, arr, tree)
        t1 = time.perf_counter()
        q = (tree[1] * u).astype(np.float32)
        out = np.zeros(m, np.intp)
        K.get_prefix_sum_idx(out, q, arr, tree)
        t2 = time.perf_counter()
        if it >= warmup:
            t_upd += t1 - t0
            t_srch += t2 - t1
    total = t_upd + t_srch
    return {
        "value": 2 * m * steps / total, "unit": UNIT, "cores": 1, "kind": kind,
        "sample": f"{steps} full steps (2^{LOG2_BATCH} updates + 2^{LOG2_BATCH} samples each) on one pinned host core; "
                  f"tree built by the C oracle",
        "updates_per_s": m * steps / t_upd, "samples_per_s": m * steps / t_srch,
        "host_cores_available": avail,
        "ms_per_step": 1e3 * total / steps,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1 << LOG2_LEAVES
    m = 1 << LOG2_BATCH
    rng = np.random.default_rng(0)
    leaves = rng.random(n, dtype=np.float32)
    batches = make_batches(rng, n, m)
    steps = max(1, min(args.steps, 8))
    warmup = max(1, min(args.warmup, 2))
    base = cpu_arm(leaves, batches, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "updates_per_s",
                                              "samples_per_s", "host_cores_available")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world):
    return {
        "workload": f"SumTreeArray float32, 2^{LOG2_LEAVES} leaves per GPU ({world} GPU: 2^{LOG2_LEAVES + (world - 1).bit_length()} "
                    f"leaves{' sharded by subtree' if world > 1 else ''}), per step 2^{LOG2_BATCH} leaf updates + 2^{LOG2_BATCH} prefix-search samples per GPU "
                    f"(BASELINE config #3; config #4 at 8 GPUs)",
        "leaves_per_gpu": 1 << LOG2_LEAVES, "updates_per_step_per_gpu": 1 << LOG2_BATCH,
        "samples_per_step_per_gpu": 1 << LOG2_BATCH,
        "l2_policy": "inputs larger than L2: the tree is 1 GiB per GPU and every step uses a different batch",
        "parallelism": "single tree" if world == 1 else f"subtree shards x{world}, NCCL exchange of deltas / roots / results",
    }


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
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
        # keep stdout to the single JSON line (the box may export NCCL_DEBUG=VERSION, which prints there)
        os.environ["NCCL_DEBUG"] = os.environ.get("ST_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    n_local = 1 << LOG2_LEAVES
    m_local = 1 << LOG2_BATCH
    n_global = n_local * world
    m_global = m_local * world
    steps, warmup = args.steps, max(3, args.warmup)

    # synthetic priorities: every rank draws its own shard; batches are replicated (same seed)
    leaves_host = np.random.default_rng(1000 + rank).random(n_local, dtype=np.float32) if world > 1 else \
        np.random.default_rng(0).random(n_local, dtype=np.float32)
    rng = np.random.default_rng(0) if world > 1 else np.random.default_rng(0)
    if world == 1:
        rng.random(n_local, dtype=np.float32)     # keep the stream identical to the reference arm
    batches_host = make_batches(rng, n_global, m_global)

    if world == 1:
        tree = DeviceSumTree(torch.from_numpy(leaves_host).to(dev))
        tree.reserve(m_local)
    else:
        from starr_b200 import ShardedSumTree
        tree = ShardedSumTree(n_global, torch.float32, local_leaves=torch.from_numpy(leaves_host).to(dev))
        tree.reserve(m_global)
    dbatches = [(torch.from_numpy(i).to(dev), torch.from_numpy(v).to(dev), torch.from_numpy(u).to(dev))
                for i, v, u in batches_host]
    out = torch.empty(m_global, dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, v, u):
        tree.update_(i, v)
        return tree.sample_uniform(u, out=out)

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for w in range(warmup):
        step(*dbatches[w % N_BATCHES])
    tree.poll()
    barrier()
    launches0 = int(_lib.lib().st_kernel_launches())
    windows = []
    t_w0 = time.time()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    # Sharded runs plan ahead: owner / slot / compaction of batch k+1 (it reads nothing of the tree) is
    # computed on a side stream while batch k is applied.  One plan per step inside the timed region.
    s_plan = torch.cuda.Stream() if world > 1 else None
    plan = tree.plan_update(*dbatches[warmup % N_BATCHES][:2], stream=s_plan) if world > 1 else None
    barrier()
    ev[0].record()
    for k in range(steps):
        i, v, u = dbatches[(warmup + k) % N_BATCHES]
        if world > 1:
            tree.update_(plan=plan)
            ev[2 * k + 1].record()
            plan = tree.plan_update(*dbatches[(warmup + k + 1) % N_BATCHES][:2], stream=s_plan)
        else:
            tree.update_(i, v)
            ev[2 * k + 1].record()
        tree.sample_uniform(u, out=out)
        ev[2 * k + 2].record()
    barrier()
    launches = int(_lib.lib().st_kernel_launches()) - launches0
    windows.append((t_w0, time.time()))
    total_ms = ev[0].elapsed_time(ev[-1])
    upd_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps))
    smp_ms = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(steps))
    tree.poll()

    # ---- end-to-end: pinned host buffers in, sampled indices back on the host -------------------
    # Sharded runs feed the batch the way a data-parallel learner would: every rank holds ITS slice of
    # the batch in pinned host memory, uploads only that (PCIe carries every byte once, not once per
    # rank), the slices are all-gathered over NVLink, and every rank reads back its slice of the result.
    mine = slice(rank * m_local, (rank + 1) * m_local) if world > 1 else slice(0, m_global)
    pinned = [(torch.from_numpy(i[mine]).pin_memory(), torch.from_numpy(v[mine]).pin_memory(),
               torch.from_numpy(u[mine]).pin_memory()) for i, v, u in batches_host]
    host_out = torch.empty(mine.stop - mine.start, dtype=torch.int64).pin_memory()

    # Software-pipelined like a real PER loop: the H2D copy of step k+1 and the D2H copy of step k-1
    # run on side streams while step k computes; every step's copies are inside the timed region.
    s_main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dbuf = [tuple(torch.empty(m_global, dtype=t.dtype, device=dev) for t in pinned[0]) for _ in range(2)]
    stage = [tuple(torch.empty_like(t, device=dev) for t in pinned[0]) for _ in range(2)] if world > 1 else dbuf
    obuf = [torch.empty(m_global, dtype=torch.int64, device=dev) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]     # inputs of the slot consumed
    ev_done = [torch.cuda.Event() for _ in range(2)]     # results of the slot ready
    ev_read = [torch.cuda.Event() for _ in range(2)]     # results of the slot copied out

    def prefetch(k):
        slot = k % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[slot])
            for dst, src in zip(stage[slot], pinned[k % N_BATCHES]):
                dst.copy_(src, non_blocking=True)
            if world > 1:
                for full, part in zip(dbuf[slot], stage[slot]):
                    dist.all_gather_into_tensor(full, part)
            ev_in[slot].record(s_in)

    def e2e_run(nsteps):
        for slot in range(2):
            ev_free[slot].record(s_main)
            ev_read[slot].record(s_main)
        prefetch(0)
        plan = None
        if world > 1:
            s_plan.wait_event(ev_in[0])
            plan = tree.plan_update(*dbuf[0][:2], stream=s_plan)
        for k in range(nsteps):
            slot = k % 2
            if k + 1 < nsteps:
                prefetch(k + 1)
            s_main.wait_event(ev_in[slot])
            s_main.wait_event(ev_read[slot])            # previous results of this slot are on the host
            di, dv, du = dbuf[slot]
            if world > 1:
                tree.update_(plan=plan)
                if k + 1 < nsteps:                      # plan of the next batch as soon as its inputs have landed
                    s_plan.wait_event(ev_in[(k + 1) % 2])
                    plan = tree.plan_update(*dbuf[(k + 1) % 2][:2], stream=s_plan)
            else:
                tree.update_(di, dv)
            tree.sample_uniform(du, out=obuf[slot])
            ev_free[slot].record(s_main)
            ev_done[slot].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[slot])
                host_out.copy_(obuf[slot][mine], non_blocking=True)
                ev_read[slot].record(s_out)
        s_main.wait_stream(s_out)

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    e0.record()
    e2e_run(steps)
    e1.record()
    barrier()
    windows.append((t_w0, time.time()))
    e2e_ms = e0.elapsed_time(e1)
    tree.poll()
    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- per-kernel durations of the update pipeline (CUDA events on the launching stream) -------
    breakdown = None
    if world == 1:
        tree.handle.set_profiling(True)
        for k in range(min(steps, 10)):
            i, v, u = dbatches[k % N_BATCHES]
            tree.update_(i, v)
        torch.cuda.synchronize()
        breakdown = tree.handle.update_breakdown()
        tree.handle.set_profiling(False)

    times = torch.tensor([total_ms, upd_ms, smp_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, upd_ms, smp_ms, e2e_ms = times.tolist()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        ops_per_step = 2 * m_global
        upd_rate = m_global * steps / (upd_ms * 1e-3)
        smp_rate = m_global * steps / (smp_ms * 1e-3)
        bu, bs = bytes_per_update(n_global), bytes_per_sample(n_global)

        def roof(name, per_unit, units_per_launch, ms_per_launch):
            ach = per_unit * units_per_launch / (ms_per_launch * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "algorithmic_bytes_per_unit": per_unit, "peak_source": peak_src,
                    "ms_per_launch": ms_per_launch}

        r_upd = roof("update pipeline, all kernels (sort + leaf pass + partition + ordered folds)", bu, m_local,
                     upd_ms / steps)
        r_smp = roof("search_kernel<float, staged, uniform>", bs, m_local, smp_ms / steps)
        traffic = ncu_traffic()
        r_smp["traffic"] = traffic.get("search_kernel")
        dominant = r_upd if upd_ms >= smp_ms else r_smp
        if breakdown and breakdown["calls"]:
            phases = {"partition_levels_kernel": breakdown["partition_ms"], "retire_kernel": breakdown["retire_ms"],
                      "fold_scan_kernel": breakdown["fold_scan_ms"],
                      "fold_huge_maps_kernel": breakdown["huge_maps_ms"],
                      "leaf_pass_kernel": breakdown["leaf_pass_ms"], "search_kernel": smp_ms / steps}
            top = max(phases, key=phases.get)
            if top != "search_kernel":
                # the update's algorithmic bytes (236 B/update) are charged to its dominant kernel
                dominant = roof(f"{top}<float> (dominant kernel of the update; all 236 B/update charged to it)", bu,
                                m_local, phases[top])
                dominant["traffic"] = traffic.get(top)
                dominant["share_of_step"] = phases[top] / (total_ms / steps)
        line = {
            "metric": METRIC, "value": ops_per_step * steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "updates_per_s": upd_rate, "samples_per_s": smp_rate,
            "update_ms_per_step": upd_ms / steps, "sample_ms_per_step": smp_ms / steps,
            "roofline": dominant, "roofline_update": r_upd, "roofline_sample": r_smp,
            "e2e": {"value": ops_per_step * steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": m_global * (8 + 4 + 8), "d2h_bytes_per_step": m_global * 8,
                    "ms_per_step": e2e_ms / steps,
                    "note": "pinned host buffers; H2D of step k+1 and D2H of step k-1 overlap step k on side streams"
                            + ("; every rank uploads / reads back its 1/N slice of the batch (byte counts are "
                               "totals over ranks), slices are all-gathered over NVLink" if world > 1 else "")},
            "gpu_launches": launches, "clocks": clocks, "update_breakdown_ms": breakdown,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_arm(leaves_host, batches_host, steps=3, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "updates_per_s",
                                                       "samples_per_s", "host_cores_available")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
