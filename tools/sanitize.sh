#!/bin/bash
# compute-sanitizer over smoke() and one large float batch (hybrid partition, deep pass, retirement, folds,
# search, axis fold); summaries go to gpurun_out/ (copy the tails into profiles/).
cd "$(dirname "$0")/.."
cat > /tmp/san_batch.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import oracle
from starr_b200 import DeviceSumTree
oracle.build()
C = oracle.c
rng = np.random.default_rng(5)
n, m = 1 << 22, 1 << 18
for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64), (np.int32, torch.int32)):
    leaves = rng.random(n).astype(dt) if dt != np.int32 else rng.integers(0, 1000, n).astype(dt)
    tree = np.zeros(n, dt); C.build_sumtree_from_array(leaves, tree)
    t = DeviceSumTree(torch.from_numpy(leaves).cuda())
    for skew in (False, True):
        idx = rng.integers(0, n if not skew else 64, m).astype(np.intp)
        vals = rng.random(m).astype(dt) if dt != np.int32 else rng.integers(0, 1000, m).astype(dt)
        C.update_sumtree(idx, vals, leaves, tree)
        t.update_(torch.from_numpy(idx.astype(np.int64)).cuda(), torch.from_numpy(vals).cuda(), check=True)
    assert t.tree.cpu().numpy().tobytes() == tree.tobytes(), dt
    u = rng.random(1 << 16)
    want = np.zeros(u.size, np.intp); C.get_prefix_sum_idx(want, (tree[1] * u).astype(dt), leaves, tree)
    assert np.array_equal(t.sample_uniform(torch.from_numpy(u).cuda()).cpu().numpy(), want)
    assert t.axis_fold(1, 2048, 2048).cpu().numpy().tobytes() == leaves.reshape(2048, 2048).sum(axis=0).tobytes()
print("sanitize batch ok")
PY
for tool in memcheck racecheck; do
  for what in smoke batch; do
    if [ $what = smoke ]; then cmd=(python -c "import __graft_entry__ as g; g.smoke()"); else cmd=(python /tmp/san_batch.py); fi
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 "${cmd[@]}" > gpurun_out/sanitize_${tool}_${what}.log 2>&1
    echo "== $tool $what: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${what}.log | tail -1)"
  done
done
