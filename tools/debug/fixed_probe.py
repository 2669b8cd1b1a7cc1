"""Why is the fixed-2^30 leg slow at 8 GPUs?  All ranks as threads on ONE GPU (tests/_thread_group.py), bench.py's
seeds for that leg; prints per-phase device times (contended -- only their ratios mean anything) and the fold statistics
of the local and the top trees."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from _thread_group import ThreadGroup
import starr_b200.sharded as sharded
from starr_b200 import DeviceSumTree

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lgn = int(sys.argv[2]) if len(sys.argv) > 2 else 27
lgm = int(sys.argv[3]) if len(sys.argv) > 3 else 20      # total
peer = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
group = ThreadGroup(world)
sharded.dist = group
dev = torch.device("cuda", 0)
nl, m = 1 << lgn, 1 << lgm
n = nl * world
gen = torch.Generator(device=dev); gen.manual_seed(20260)
bat = [(torch.randint(0, n, (m,), generator=gen, device=dev), torch.rand(m, generator=gen, device=dev),
        torch.rand(m, generator=gen, device=dev, dtype=torch.float64)) for _ in range(8)]
# warm every kernel up before the ranks start (see tests/test_sharded_gpu.py)
w = DeviceSumTree(torch.rand(1 << 20, device=dev)); w.reserve(m)
wi = torch.randint(0, 1 << 20, (m // world,), device=dev); wv = torch.rand(m // world, device=dev)
w.update_(wi, wv); w.update_begin_(wi, wv, torch.empty_like(wv)); w.update_finish_(); w.sample_uniform(bat[0][2]); w.poll()
torch.cuda.synchronize(); del w
res = {}

def body(r):
    torch.cuda.set_device(0)
    with torch.cuda.stream(torch.cuda.Stream()):
        g = torch.Generator(device=dev); g.manual_seed(4242 + r)
        lv = torch.rand(nl, device=dev, generator=g)
        st = sharded.ShardedSumTree(n, torch.float32, local_leaves=lv, peer=peer)
        del lv
        st.reserve(m)
        for i, v, u in bat[:3]:
            st.update_(i, v); st.sample_uniform(u)
        torch.cuda.current_stream().synchronize()
        k = [3]
        def nxt():
            b = bat[k[0] % len(bat)]; k[0] += 1; return b
        ph = st.phase_times(nxt, reps=5)
        s = torch.cuda.current_stream().cuda_stream
        res[r] = (ph, st.local.handle.update_stats(s), st.top.handle.update_stats(s), st.top.heap.cpu().numpy().copy(),
                  st.local.heap[:8].cpu().numpy().copy())
        torch.cuda.current_stream().synchronize()
        st.close()

group.run(body)
np.set_printoptions(precision=1, floatmode="fixed", linewidth=200)
for r in sorted(res):
    ph, ls, ts, top, loc = res[r]
    print(f"rank {r}: local stats {ls}\n        top stats {ts}\n        top heap {top}\n        local heap[:8] {loc}  log2 {np.log2(np.maximum(loc,1))}")
print("phases (us, rank 0):")
for kname, v in res[0][0]:
    print(f"  {kname:60s} {v:9.1f}   max over ranks {max(dict(res[r][0])[kname] for r in res):9.1f}")
