// st_partition.cuh -- one persistent cooperative kernel that walks the tree top-down and keeps
// the update batch ordered by (ancestor at the current level, batch position j).
//
// Why: for floating point the reference's result at every node is the ORDERED fold of the
// updates below it (starr/src/_sumtree_array.pyx:41-49).  Ordering the batch by
// (node at level l, j) makes every node's updates one contiguous, batch-ordered segment; going
// from level l to l+1 is a stable split of every segment by the next key bit (an MSD radix
// step).  The whole descent runs inside ONE launch: per level two grid-wide barriers
// (zero-count scan published -> scatter), all arrays (<= 4 MiB each per 2^20 updates) stay in L2.
//
// Per level l every segment head does the fold for its node:
//    1 update            -> one add
//    2 .. SHORT_SEG      -> sequential adds by the head thread (already in batch order)
//    longer              -> queued for the exact parallel ordered fold (st_ordered_fold.cuh),
//                           which reads the level's ordered differences from lev_d[l].
#pragma once
#include <cooperative_groups.h>

#include "st_common.cuh"
#include "st_update_defs.cuh"

namespace st {

namespace cg = cooperative_groups;

constexpr int PART_THREADS = 256;
constexpr int PART_E = 4;
constexpr int PART_TILE = PART_THREADS * PART_E;  // 1024
constexpr int PART_MAX_TILES = 1024;              // UPDATE_MAX_CHUNK / PART_TILE

template <typename T> struct part_params {
    T *H;
    uint32_t n;
    int depth;
    int m;
    int ntiles;
    const uint32_t *key0;  // keys in batch order (level 0)
    uint32_t *keyA, *keyB; // ping-pong for levels >= 1
    T *lev_d;              // lev_d[l * lev_stride + i]: differences in level-l order; level 0 filled by caller
    int64_t lev_stride;
    int *sA, *eA, *sB, *eB;  // segment [s, e) of every element, ping-pong (level 0: [0, m))
    int *Z;                  // exclusive count of zero-bit elements before i
    unsigned int *tilesum;   // [2][PART_MAX_TILES] zero-bit counts per tile
    seg_entry *qlong, *qmed, *qhuge;
    int *huge_off;   // first chunk record of every huge segment
    unsigned int *status;
};

// exclusive scan of `v` over the CTA; returns the prefix, *total gets the CTA sum (valid in all threads)
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, unsigned int *s_warp, unsigned int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    __syncthreads();  // s_warp free
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned int pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PART_THREADS / 32; ++w) {
        const unsigned int t = s_warp[w];
        if (w < warp) pre += t;
        tot += t;
    }
    *total = tot;
    return pre + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(PART_THREADS) partition_levels_kernel(part_params<T> p)
{
    cg::grid_group grid = cg::this_grid();
    if (p.status[SW_STATUS]) return;  // uniform over the grid: a failed batch is skipped as a whole
    __shared__ unsigned int s_tpre[PART_MAX_TILES + 1];
    __shared__ unsigned int s_warp[PART_THREADS / 32];
    const int tid = threadIdx.x;
    const int m = p.m, depth = p.depth;
    const uint64_t two_n = 2ull * p.n;

    // ---- P0: zero counts of the first split bit per tile; clear the other counter set -----------
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        unsigned int z = 0;
        if (depth >= 2 && m > SHORT_SEG) {
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = t * PART_TILE + tid * PART_E + k;
                if (i < m) z += ((p.key0[i] >> (depth - 1)) & 1u) ^ 1u;
            }
        }
        unsigned int tot;
        block_excl_scan(z, s_warp, &tot);
        if (tid == 0) {
            p.tilesum[t] = tot;
            p.tilesum[PART_MAX_TILES + t] = 0;
        }
    }
    grid.sync();

    for (int l = 0; l < depth; ++l) {
        const int cur = l & 1;
        const int sh = depth - l;            // node at this level = key >> sh
        const bool last = (l == depth - 1);
        const uint32_t *key = (l == 0) ? p.key0 : (cur ? p.keyA : p.keyB);
        uint32_t *keyN = cur ? p.keyB : p.keyA;   // level l+1 lives in A when l+1 is odd
        const int *sC = cur ? p.sA : p.sB, *eC = cur ? p.eA : p.eB;
        int *sN = cur ? p.sB : p.sA, *eN = cur ? p.eB : p.eA;
        const T *dC = p.lev_d + (int64_t)l * p.lev_stride;
        T *dN = p.lev_d + (int64_t)(l + 1) * p.lev_stride;
        unsigned int *tsC = p.tilesum + cur * PART_MAX_TILES;
        unsigned int *tsN = p.tilesum + (cur ^ 1) * PART_MAX_TILES;

        // tile prefix of the zero counts (every CTA scans the <= 1024 tile sums itself)
        unsigned int ztotal = 0;
        if (!last) {
            unsigned int v[PART_E], run = 0;
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int t = tid * PART_E + k;
                v[k] = (t < p.ntiles) ? tsC[t] : 0u;
                run += v[k];
            }
            unsigned int pre = block_excl_scan(run, s_warp, &ztotal);
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                s_tpre[tid * PART_E + k] = pre;
                pre += v[k];
            }
            __syncthreads();
        }

        // ---- P1: segment heads fold / retire / queue; publish Z ------------------------------------------
        if (blockIdx.x == 0 && tid == 0) p.status[SW_ANYLONG + ((l + 1) & 1)] = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            unsigned int zloc[PART_E], zsum = 0;
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = t * PART_TILE + tid * PART_E + k;
                zloc[k] = 0;
                if (i < m) {
                    const int eraw = (l == 0) ? m : eC[i];
                    if (!(eraw & SEG_DONE)) {
                        const int s = (l == 0) ? 0 : sC[i];
                        const int e = eraw;
                        const int len = e - s;
                        const uint32_t ky = key[i];
                        if (len <= SHORT_SEG) {
                            if (i == s)
                                retire_segment<T>(p.H, s, e, l, depth, two_n, [&](int x) { return key[x]; },
                                                  [&](int x) { return dC[x]; });
                        } else {
                            if (i == s && l < key_leaf_level(ky, two_n, depth)) {
                                const uint32_t node = ky >> sh;
                                if (len > HUGE_SEG) {
                                    const unsigned int slot = atomicAdd(p.status + SW_QCOUNT_HUGE, 1u);
                                    p.qhuge[slot] = seg_entry{l, s, len, node};
                                    p.huge_off[slot] = (int)atomicAdd(p.status + SW_HUGE_CHUNKS,
                                                                      (unsigned int)((len + 1023) / 1024));
                                } else if (len > MEDIUM_SEG) {
                                    p.qlong[atomicAdd(p.status + SW_QCOUNT, 1u)] = seg_entry{l, s, len, node};
                                } else {
                                    p.qmed[atomicAdd(p.status + SW_QCOUNT_MED, 1u)] = seg_entry{l, s, len, node};
                                }
                            }
                            if (i == s) p.status[SW_ANYLONG + (l & 1)] = 1;   // another split is needed
                            if (!last) zloc[k] = ((ky >> (sh - 1)) & 1u) ^ 1u;
                        }
                    }
                }
                zsum += zloc[k];
            }
            if (!last) {
                unsigned int tot;
                unsigned int pre = block_excl_scan(zsum, s_warp, &tot) + s_tpre[t];
#pragma unroll
                for (int k = 0; k < PART_E; ++k) {
                    const int i = t * PART_TILE + tid * PART_E + k;
                    if (i < m) p.Z[i] = (int)pre;
                    pre += zloc[k];
                }
                if (tid == 0) tsN[t] = 0;   // counters of the level after next start from zero
            }
        }
        if (last) break;
        grid.sync();
        if (((volatile unsigned int *)p.status)[SW_ANYLONG + (l & 1)] == 0) break;   // only retired segments are left (grid-uniform)

        // ---- P2: stable split of every long segment by the next key bit ------------------------------------
        const bool need_next = (l + 1 < depth - 1);   // level l+1 may split again by bit sh-2
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = t * PART_TILE + tid * PART_E + k;
                const bool live = i < m;
                int np = 0;
                unsigned int znext = 0;
                if (live) {
                    const int eraw = (l == 0) ? m : eC[i];
                    const int s = (l == 0) ? 0 : sC[i];
                    const int e = eraw & ~SEG_DONE;
                    if ((eraw & SEG_DONE) || e - s <= SHORT_SEG) {
                        // retired: stays where it is, nobody reads its key / difference again
                        np = i;
                        sN[i] = s;
                        eN[i] = e | SEG_DONE;
                    } else {
                        const uint32_t ky = key[i];
                        const int zs = p.Z[s];
                        const int ze = (e < m) ? p.Z[e] : (int)ztotal;
                        const int zb = p.Z[i] - zs;      // zero-bit elements of my segment before me
                        const int nz = ze - zs;          // zero-bit elements of my segment
                        int s2, e2;
                        if (((ky >> (sh - 1)) & 1u) == 0) {
                            np = s + zb;
                            s2 = s;
                            e2 = s + nz;
                        } else {
                            np = s + nz + (i - s - zb);
                            s2 = s + nz;
                            e2 = e;
                        }
                        keyN[np] = ky;
                        dN[np] = dC[i];
                        sN[np] = s2;
                        eN[np] = e2;
                        if (need_next && e2 - s2 > SHORT_SEG) znext = ((ky >> (sh - 2)) & 1u) ^ 1u;
                    }
                }
                if (need_next) {
                    // one atomic per (warp, destination tile)
                    const int dt = live ? (np / PART_TILE) : -1;
                    const unsigned int peers = __match_any_sync(0xffffffffu, dt);
                    const unsigned int zmask = __ballot_sync(0xffffffffu, znext != 0);
                    if (live && (threadIdx.x & 31) == (__ffs(peers) - 1)) {
                        const unsigned int c = __popc(peers & zmask);
                        if (c) atomicAdd(tsN + dt, c);
                    }
                }
            }
        }
        grid.sync();
    }
}

}  // namespace st
