"""H2D / D2H bandwidth of pinned host memory allocated while the process is bound to each NUMA node (is the e2e leg
bound by a cross-socket hop?)."""
import os, sys, time, glob
import numpy as np, torch
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
try:
    aff = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [i for i in range(os.cpu_count()) if (aff[i // 64] >> (i % 64)) & 1]
    print("GPU0 cpu affinity:", cpus[:8], "...", len(cpus), "cpus")
except Exception as e:
    print("affinity query failed:", e); cpus = []
print("allowed cpus:", sorted(os.sched_getaffinity(0))[:8], "...", len(os.sched_getaffinity(0)))
nodes = {}
for p in glob.glob("/sys/devices/system/node/node*/cpulist"):
    nodes[p.split("/")[-2]] = open(p).read().strip()
print("numa nodes:", nodes)
os.system("nvidia-smi topo -m 2>/dev/null | head -12")
dev = torch.device("cuda")
def bw(nbytes=20 << 20, reps=20):
    hbuf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3): dbuf.copy_(hbuf, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): dbuf.copy_(hbuf, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    h2d = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    a.record()
    for _ in range(reps): hbuf.copy_(dbuf, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    d2h = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    return h2d, d2h
print("default placement: H2D %.1f GB/s, D2H %.1f GB/s" % bw())
allowed = sorted(os.sched_getaffinity(0))
def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part: out.append(int(part))
    return out
for name, cl in sorted(nodes.items()):
    c = [x for x in parse(cl) if x in allowed]
    if not c: print(name, "no allowed cpus"); continue
    os.sched_setaffinity(0, c)
    print(name, "H2D %.1f GB/s, D2H %.1f GB/s" % bw())
os.sched_setaffinity(0, allowed)
print("256 MB copies: H2D %.1f GB/s, D2H %.1f GB/s" % bw(256 << 20, 5))
