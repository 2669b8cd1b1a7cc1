"""GPU timeline of one sharded step on rank 0 (torch.profiler / CUPTI): kernel start offsets, durations
and the idle gaps between them.  Diagnostics only -- nothing measured under the profiler is a bench number.
Launch with torchrun like tools/sharded_breakdown.py."""
import json, os, sys, tempfile
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from starr_b200 import ShardedSumTree
n_local, m = 1 << 27, (1 << 20) * world
tree = ShardedSumTree(n_local * world, torch.float32,
                      local_leaves=torch.from_numpy(np.random.default_rng(1000 + rank).random(n_local, dtype=np.float32)).to(dev))
tree.reserve(m)
rng = np.random.default_rng(0)
B = [(torch.from_numpy(rng.integers(0, n_local * world, m)).to(dev), torch.from_numpy(rng.random(m, dtype=np.float32)).to(dev),
      torch.from_numpy(rng.random(m)).to(dev)) for _ in range(3)]
out = torch.empty(m, dtype=torch.int64, device=dev)
side = torch.cuda.Stream()
def steps(k):
    plan = tree.plan_update(*B[0][:2], stream=side)
    for s in range(k):
        tree.update_(plan=plan)
        plan = tree.plan_update(*B[(s + 1) % 3][:2], stream=side)
        tree.sample_uniform(B[s % 3][2], out=out)
    torch.cuda.synchronize()
steps(4)
dist.barrier(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    steps(4)
path = os.path.join(tempfile.gettempdir(), f"trace{rank}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
names = [e["name"] for e in ev]
first = [i for i, nm in enumerate(names) if "validate_kernel" in nm]
lo, hi = first[2], first[3]
t0 = ev[lo]["ts"]; prev_end = t0
os.makedirs("gpurun_out", exist_ok=True)
if rank == 0:
    with open("gpurun_out/sharded_trace.md", "w") as f:
        f.write(f"| start us | dur us | gap before us | stream | kernel ({world} GPUs, rank 0, one step) |\n|---|---|---|---|---|\n")
        for e in ev[lo:hi]:
            gap = e["ts"] - prev_end
            f.write(f"| {e['ts'] - t0:.1f} | {e['dur']:.1f} | {gap:.1f} | {e['args'].get('stream')} | {e['name'][:70]} |\n")
            prev_end = max(prev_end, e["ts"] + e["dur"])
        f.write(f"\nstep length {ev[hi]['ts'] - t0:.1f} us\n")
# per-rank summary: when each phase ends relative to the step's first kernel
def end_of(sub, nth=0):
    hits = [e for e in ev[lo:hi] if sub in e["name"]]
    return (hits[nth]["ts"] + hits[nth]["dur"] - t0, hits[nth]["dur"]) if len(hits) > nth else (float("nan"), float("nan"))
row = [rank, end_of("partition_levels")[1], end_of("fold_scan_kernel")[0], end_of("AllGather", 0)[1], end_of("AllGather", 0)[0],
       end_of("AllGather", 1)[1], end_of("fold_huge_apply_kernel<float, true")[0], end_of("AllGather", 2)[1], ev[hi]["ts"] - t0]
with open(f"gpurun_out/trace_rank{rank}.txt", "w") as f:
    f.write("rank %d: partition %.0f us | local update ends %.0f | AG(deltas) dur %.0f ends %.0f | AG(recs) dur %.0f | top fold ends %.0f | AG(idx) dur %.0f | step %.0f\n" % tuple(row))
dist.barrier()
dist.destroy_process_group()
