"""Executable specification (CPU, NumPy only) of the index arithmetic of the slices layout (starr_b200/_slices.py,
csrc/st_shard.cu shard_partition_kernel, csrc/st_peer.cu slices_send / slices_return / slices_unpermute kernels and the
matrix-free sample: route_push / segmented search / scatter_pairs).

Every offset the kernels derive on the device comes from the G x G count matrix M[r][q] (elements of rank r's slice owned
by shard q).  This file restates those derivations and checks the properties the bit-parity of the sharded tree rests on:

  * an owner receives its elements contiguously and in GLOBAL batch order (rank-major, then slice order) -- the local
    update applies them in the order ONE reference tree would (reference pyx:41-49);
  * a result written back at `sum_{q' < q} M[r][q'] + k` lands where `start[shard[j]] + slot[j]` looks for it;
  * the matrix-free sample addresses every result by (source rank, position in that source's per-owner run).
The GPU tests check the kernels themselves; this one pins the protocol."""
import numpy as np
import pytest


def partition(owner, G):
    """shard_partition_kernel: stable slot of every element among those of its shard, and the partitioned order."""
    slot = np.empty(owner.size, np.int64)
    counts = np.bincount(owner, minlength=G)
    for q in range(G):
        sel = np.flatnonzero(owner == q)
        slot[sel] = np.arange(sel.size)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    order = np.argsort(owner, kind="stable")          # element at partitioned position start[q] + k
    return slot, counts, start, order


@pytest.mark.parametrize("G,m,skew", [(2, 1, False), (4, 37, False), (8, 1000, False), (8, 1000, True), (4, 0, False)])
def test_update_exchange_keeps_global_batch_order(G, m, skew):
    rng = np.random.default_rng(G * 1000 + m)
    n_local = 64
    idx = [rng.integers(0, G * n_local, m) for _ in range(G)]            # rank r's slice: GLOBAL leaf ids
    if skew and m:
        idx[0][:] = rng.integers(0, n_local, m)                           # all of rank 0's slice belongs to shard 0
        idx[1][: m // 2] = 5                                              # one leaf, many times
    vals = [rng.random(m) for _ in range(G)]
    owner = [i // n_local for i in idx]
    parts = [partition(o, G) for o in owner]
    M = np.array([p[1] for p in parts])                                   # count matrix: row r = rank r's counts
    # ---- slices_send_kernel: element k of rank r's partitioned slice -> owner q at base[q] + (k - start[q]),
    #      base[q] = sum_{r' < r} M[r'][q]
    recv_idx = [np.full(M[:, q].sum(), -1, np.int64) for q in range(G)]
    recv_val = [np.full(M[:, q].sum(), np.nan) for q in range(G)]
    recv_src = [np.full((M[:, q].sum(), 2), -1, np.int64) for q in range(G)]   # (source rank, slice position): for the check
    for r in range(G):
        slot, counts, start, order = parts[r]
        for k, j in enumerate(order):
            q = owner[r][j]
            di = M[:r, q].sum() + (k - start[q])
            recv_idx[q][di] = idx[r][j] - q * n_local
            recv_val[q][di] = vals[r][j]
            recv_src[q][di] = (r, j)
    for q in range(G):
        assert (recv_idx[q] >= 0).all() and (recv_idx[q] < n_local).all()
        # rank-major, then slice order: exactly the subsequence of the concatenated batch that shard q owns
        want = [(r, j) for r in range(G) for j in range(m) if owner[r][j] == q]
        assert [tuple(x) for x in recv_src[q]] == want
    # ---- the owner's results (here: a tag of the element) in receive order -> slices_return_kernel:
    #      element k came from source r with rstart[r] <= k < rstart[r + 1]; it goes to r at
    #      sum_{q' < q} M[r][q'] + (k - rstart[r])
    ret = [np.full(m, -1.0) for _ in range(G)]
    for q in range(G):
        rstart = np.concatenate([[0], np.cumsum(M[:, q])])
        for k in range(recv_val[q].size):
            r = int(np.searchsorted(rstart, k, side="right") - 1)
            ret[r][M[r, :q].sum() + (k - rstart[r])] = recv_val[q][k] * 2 + 1          # "difference" of that element
    # ---- slices_unpermute_kernel: result of slice element j = ret[start[shard[j]] + slot[j]]
    for r in range(G):
        slot, counts, start, order = parts[r]
        got = ret[r][start[owner[r]] + slot] if m else np.zeros(0)
        assert np.array_equal(got, vals[r] * 2 + 1)


@pytest.mark.parametrize("G,m", [(2, 3), (8, 500), (4, 0)])
def test_sample_push_addresses_results_by_source_and_run_position(G, m):
    rng = np.random.default_rng(G + m)
    owner = [rng.integers(0, G, m) for _ in range(G)]                     # route: owner of rank r's query j
    resid = [rng.random(m) for _ in range(G)]
    cap = max(m, 1)
    # route_push_kernel: any order inside a run, position from a counter per owner; region of source r on owner q
    rq_resid = np.full((G, G, cap), np.nan)                                # [owner][source][k]
    rq_pos = np.full((G, G, cap), -1, np.int64)
    cnt = np.zeros((G, G), np.int64)                                       # [source][owner]
    for r in range(G):
        for j in rng.permutation(m):                                       # unordered append
            q = owner[r][j]
            k = cnt[r, q]; cnt[r, q] += 1
            rq_resid[q, r, k] = resid[r][j]; rq_pos[q, r, k] = j
    # segmented search on owner q: flat query i -> (source r, k); result pair (pos, leaf) -> back[r][q][k]
    back = np.full((G, G, cap, 2), -1.0)                                   # [source][owner][k] = (slice position, "leaf")
    for q in range(G):
        seg = np.concatenate([[0], np.cumsum(cnt[:, q])])
        for i in range(seg[-1]):
            r = int(np.searchsorted(seg, i, side="right") - 1)
            k = i - seg[r]
            back[r, q, k] = (rq_pos[q, r, k], rq_resid[q, r, k] * 3)       # the "leaf" found for that prefix
    # peer_scatter_pairs_kernel: out[pos] = q * leaves_per_shard + leaf, for the cnt[r][q] pairs owner q returned
    for r in range(G):
        out = np.full(m, -1.0)
        for q in range(G):
            for k in range(cnt[r, q]):
                pos, leaf = back[r, q, k]
                out[int(pos)] = q * 1000 + leaf
        assert np.array_equal(out, owner[r] * 1000 + resid[r] * 3)
