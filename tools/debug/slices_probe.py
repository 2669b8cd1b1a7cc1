"""Debugging aid: one tiny distributed batch through the slices layout with all ranks as threads on one GPU; prints what
every rank saw (count matrix, received elements, differences, roots) next to the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("STARR_B200_WAIT_TIMEOUT_S", "3")
os.environ.setdefault("STARR_B200_SLICES_SERIAL", "1")
import numpy as np, torch
import oracle
from _thread_group import ThreadGroup
import starr_b200.sharded as sharded
from starr_b200 import DeviceSumTree
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
group = ThreadGroup(world); sharded.dist = group
dev = torch.device("cuda", 0)
n = 1 << 16; nl = n // world
rng = np.random.default_rng(12)
leaves = np.exp(rng.uniform(-8, 8, n)).astype(np.float32)
ref_l, ref_t = leaves.copy(), np.zeros(n, np.float32)
oracle.c.build_sumtree_from_array(ref_l, ref_t)
idx = rng.integers(0, n, m * world).astype(np.int64); vals = np.exp(rng.uniform(-8, 8, m * world)).astype(np.float32)
oracle.c.update_sumtree(idx.astype(np.intp), vals, ref_l, ref_t)
w = DeviceSumTree(torch.from_numpy(leaves[:nl].copy()).to(dev)); w.reserve(1 << 16)
wi = torch.from_numpy(idx % nl).to(dev); wv = torch.from_numpy(vals).to(dev)
w.update_begin_(wi, wv, torch.empty_like(wv)); w.update_finish_(); w.sample_uniform(torch.rand(100, device=dev, dtype=torch.float64)); w.poll()
torch.cuda.synchronize(); del w
out = {}
import threading
lock = threading.Lock()
def body(r):
    torch.cuda.set_device(0)
    with torch.cuda.stream(torch.cuda.Stream()):
        st = sharded.ShardedSumTree(n, torch.float32, local_leaves=torch.from_numpy(leaves[r * nl:(r + 1) * nl].copy()).to(dev), peer=True)
        st.reserve_slices(max(m, 1024), expected_local=max(m, 1024) * world)
        st.update_slices_(torch.from_numpy(idx[r * m:(r + 1) * m]).to(dev), torch.from_numpy(vals[r * m:(r + 1) * m]).to(dev))
        torch.cuda.current_stream().synchronize()
        sx = st._slices
        par = 1
        mat = sx._view("mat_u", par, torch.int32, world * (world + 1)).cpu().numpy().reshape(world, world + 1)
        c = int(mat[:, r].sum())
        info = dict(mat=mat, c=c, ridx=sx._view("recv_idx", par, torch.int64, max(c, 1)).cpu().numpy()[:c],
                    rval=sx._view("recv_val", par, torch.float32, max(c, 1)).cpu().numpy()[:c],
                    shard=sx.shard[par][:m].cpu().numpy(), slot=sx.slot[par][:m].cpu().numpy(),
                    dslice=sx._view("dslice", par, torch.float32, m).cpu().numpy(), roots=sx._view("roots", par, torch.float32, world).cpu().numpy(),
                    top=st.top.heap.cpu().numpy())
        try:
            sx.area.status(torch.cuda.current_stream().cuda_stream); info["wait"] = "ok"
        except RuntimeError as e:
            info["wait"] = str(e)
        full_t = st.gather_sumtree().cpu().numpy(); full_l = st.gather_leaves().cpu().numpy()
        info["tree_ok"] = full_t.tobytes() == ref_t.tobytes(); info["leaves_ok"] = full_l.tobytes() == ref_l.tobytes()
        if not info["leaves_ok"]:
            bad = np.flatnonzero(full_l != ref_l); info["bad_leaves"] = (bad[:8], full_l[bad[:8]], ref_l[bad[:8]])
        if not info["tree_ok"]:
            bad = np.flatnonzero(full_t.view(np.uint32) != ref_t.view(np.uint32)); info["bad_nodes"] = (len(bad), bad[:12])
        with lock: out[r] = info
        st.close()
group.run(body)
np.set_printoptions(linewidth=220)
print("global idx", idx, "owners", idx // nl)
print("expected top", ref_t[:world], "roots", ref_t[world:2 * world] if world * 2 <= n else None)
for r in sorted(out):
    print(f"--- rank {r}")
    for k, v in out[r].items():
        print(f"  {k}: {v}")
