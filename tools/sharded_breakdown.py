"""Per-phase device timing of one sharded step plus the NCCL collectives it could use.

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
            --master-port 29511 tools/sharded_breakdown.py [log2_leaves_per_gpu] [log2_batch_per_gpu]
Rank 0 prints a markdown table (median microseconds over the repetitions, max over ranks).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    lg_n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    lg_m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from starr_b200 import ShardedSumTree

    n_local, m = 1 << lg_n, (1 << lg_m) * world
    n = n_local * world
    leaves = torch.from_numpy(np.random.default_rng(1000 + rank).random(n_local, dtype=np.float32)).to(dev)
    tree = ShardedSumTree(n, torch.float32, local_leaves=leaves)
    if os.environ.get("ST_LAYOUT", "slices") != "slices":
        tree.reserve(m)
    rng = np.random.default_rng(0)
    batches = [(torch.from_numpy(rng.integers(0, n, m)).to(dev),
                torch.from_numpy(rng.random(m, dtype=np.float32)).to(dev),
                torch.from_numpy(rng.random(m)).to(dev)) for _ in range(4)]
    rows = []

    def timed(name, fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        t = torch.tensor([float(np.median(ts))], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append((name, float(t)))

    k = [0]

    def nxt():
        k[0] += 1
        return batches[k[0] % len(batches)]

    if os.environ.get("ST_LAYOUT", "slices") == "slices":
        # slices layout: every rank supplies its own 1/N of the batch (ShardedSumTree.update_slices_ / sample_slices)
        ml = 1 << lg_m
        tree.reserve_slices(ml, expected_local=ml * 5 // 4)
        rs = np.random.default_rng(100 + rank)
        sl = [(torch.from_numpy(rs.integers(0, n, ml)).to(dev), torch.from_numpy(rs.random(ml, dtype=np.float32)).to(dev),
               torch.from_numpy(rs.random(ml)).to(dev)) for _ in range(6)]
        ks = [0]

        def nxs():
            ks[0] += 1
            return sl[ks[0] % len(sl)]
        timed("update_slices_ (whole)", lambda: tree.update_slices_(*nxs()[:2]))
        timed("sample_slices (whole)", lambda: tree.sample_slices(nxs()[2]))
        for name, us in tree.phase_times(nxs, reps=10, slices=True):
            t = torch.tensor([us], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows.append(("  phase: " + name, float(t)))
    else:
        timed("update_ (whole)", lambda: tree.update_(*nxt()[:2]))
        timed("sample_uniform (whole)", lambda: tree.sample_uniform(nxt()[2]))
        for name, us in tree.phase_times(nxt, reps=10):
            t = torch.tensor([us], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows.append(("  phase: " + name, float(t)))

    # collectives at the sizes the step could use (ST_NO_COLLECTIVES=1 skips them)
    if os.environ.get("ST_NO_COLLECTIVES") == "1":
        m = 0
    f32 = torch.zeros(m, dtype=torch.float32, device=dev)
    i64 = torch.zeros(m, dtype=torch.int64, device=dev)
    i32 = torch.zeros(m, dtype=torch.int32, device=dev)
    part_f = torch.zeros(m // world, dtype=torch.float32, device=dev)
    part_i = torch.zeros(m // world, dtype=torch.int32, device=dev)
    tiny = torch.zeros(1, dtype=torch.float32, device=dev)
    tiny_all = torch.zeros(world, dtype=torch.float32, device=dev)
    if m == 0:
        timed = lambda *a, **k: None                                            # noqa: E731
    timed(f"all_reduce f32[{m}]", lambda: dist.all_reduce(f32))
    timed(f"all_reduce i64[{m}]", lambda: dist.all_reduce(i64))
    timed(f"all_reduce i32[{m}]", lambda: dist.all_reduce(i32))
    timed(f"reduce_scatter f32[{m}] -> [{m // world}]", lambda: dist.reduce_scatter_tensor(part_f, f32))
    timed(f"all_gather f32[{m // world}] -> [{m}]", lambda: dist.all_gather_into_tensor(f32, part_f))
    timed(f"all_gather i32[{m // world}] -> [{m}]", lambda: dist.all_gather_into_tensor(i32, part_i))
    timed("all_gather f32[1] (shard roots)", lambda: dist.all_gather_into_tensor(tiny_all, tiny))
    timed("two back-to-back tiny all_gathers", lambda: (dist.all_gather_into_tensor(tiny_all, tiny),
                                                        dist.all_gather_into_tensor(tiny_all, tiny)))
    if rank == 0:
        print(f"| phase ({world} GPUs, 2^{lg_n} leaves and 2^{lg_m} updates/samples per GPU) | us |")
        print("|---|---|")
        for name, us in rows:
            print(f"| {name} | {us:.1f} |")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
