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

constexpr int FOLD_CHUNK_ELEMS = 1024;   // == FOLD_CHUNK of st_ordered_fold.cuh (chunk records of huge segments)

constexpr int PART_THREADS = 256;
constexpr int PART_E = 8;
constexpr int PART_TILE = PART_THREADS * PART_E;                       // 2048
constexpr int PART_MAX_TILES = (int)(UPDATE_MAX_CHUNK / PART_TILE);    // 1024

template <typename T> struct part_params {
    T *H;
    uint32_t n;
    int depth;
    int m;
    int ntiles;
    const uint32_t *key0;  // keys in batch order (level 0)
    uint32_t *lev_key;     // lev_key[l * lev_stride + i]: keys in level-l order for l >= 1 (kept for the retire kernel)
    T *lev_d;              // lev_d[l * lev_stride + i]: differences in level-l order; level 0 filled by caller
    int64_t lev_stride;
    int *sA, *eA, *sB, *eB;  // segment [s, e) of every element, ping-pong (level 0: [0, m))
    int *Z;                  // exclusive count of zero-bit elements before i
    unsigned int *tilesum;   // [2][PART_MAX_TILES] zero-bit counts per tile
    seg_entry *qlong, *qmed, *qhuge;
    seg_entry *qret;       // short segments to retire (node unused): finished by retire_kernel after this kernel
    int *huge_off;   // first chunk record of every huge segment
    unsigned int *status;
    int l0;          // first level to process.  0: the whole descent.  > 0: fallback of the hybrid partition
                     // (st_partition_hybrid.cuh) -- lev_key[l0] / lev_d[l0] and the segment bounds of level l0 are
                     // given, level l0 is already classified, and the kernel only runs if SW_NEED_COOP is set
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

#ifdef ST_SMALL_TIMING
__device__ unsigned long long g_part_ticks[256];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PT_TICK(k) if (blockIdx.x == 0 && threadIdx.x == 0 && (k) < 256) g_part_ticks[(k)] = gtime()
#else
#define PT_TICK(k)
#endif

template <typename T>
__global__ void __launch_bounds__(PART_THREADS) partition_levels_kernel(part_params<T> p)
{
    cg::grid_group grid = cg::this_grid();
    PT_TICK(0);
    if (p.status[SW_STATUS]) return;  // uniform over the grid: a failed batch is skipped as a whole
    if (p.l0 > 0 && !p.status[SW_NEED_COOP]) return;   // the per-segment kernel has done everything (grid-uniform)
    __shared__ unsigned int s_tpre[PART_MAX_TILES + 1];
    __shared__ unsigned int s_warp[PART_THREADS / 32];
    __shared__ unsigned int s_wk[PART_E * (PART_THREADS / 32)];
    __shared__ unsigned int s_hist[PART_MAX_TILES];
    for (int j = threadIdx.x; j < PART_MAX_TILES; j += PART_THREADS) s_hist[j] = 0;
    const int tid = threadIdx.x;
    const int m = p.m, depth = p.depth;
    const uint64_t two_n = 2ull * p.n;

    // ---- P0: zero counts of the first split bit per tile; clear the other counter set -----------
    const int l0 = p.l0;
    {
        const int cur0 = l0 & 1;
        const uint32_t *key = (l0 == 0) ? p.key0 : p.lev_key + (int64_t)l0 * p.lev_stride;
        const int *sC = cur0 ? p.sA : p.sB, *eC = cur0 ? p.eA : p.eB;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            unsigned int z = 0;
            if (depth - l0 >= 2 && (l0 > 0 || m > RETIRE_SEG)) {
#pragma unroll
                for (int k = 0; k < PART_E; ++k) {
                    const int i = t * PART_TILE + k * PART_THREADS + tid;
                    if (i < m) {
                        bool live = true;
                        if (l0 > 0) {
                            const int e = eC[i];
                            live = !(e & SEG_DONE) && (e - sC[i] > RETIRE_SEG);
                        }
                        if (live) z += ((key[i] >> (depth - l0 - 1)) & 1u) ^ 1u;
                    }
                }
            }
            unsigned int tot;
            block_excl_scan(z, s_warp, &tot);
            if (tid == 0) {
                p.tilesum[cur0 * PART_MAX_TILES + t] = tot;
                p.tilesum[(cur0 ^ 1) * PART_MAX_TILES + t] = 0;
            }
        }
    }
    grid.sync();
    PT_TICK(1);

    for (int l = l0; l < depth; ++l) {
        const bool classify = !(l0 > 0 && l == l0);   // level l0 of the hybrid fallback was classified by hy_scan_kernel
        const int cur = l & 1;
        const int sh = depth - l;            // node at this level = key >> sh
        const bool last = (l == depth - 1);
        const uint32_t *key = (l == 0) ? p.key0 : p.lev_key + (int64_t)l * p.lev_stride;
        uint32_t *keyN = p.lev_key + (int64_t)(l + 1) * p.lev_stride;
        const int *sC = cur ? p.sA : p.sB, *eC = cur ? p.eA : p.eB;
        int *sN = cur ? p.sB : p.sA, *eN = cur ? p.eB : p.eA;
        const T *dC = p.lev_d + (int64_t)l * p.lev_stride;
        T *dN = p.lev_d + (int64_t)(l + 1) * p.lev_stride;
        unsigned int *tsC = p.tilesum + cur * PART_MAX_TILES;
        unsigned int *tsN = p.tilesum + (cur ^ 1) * PART_MAX_TILES;

        // tile prefix of the zero counts (every CTA scans the <= 1024 tile sums itself)
        unsigned int ztotal = 0;
        if (!last) {
            constexpr int TPT = (PART_MAX_TILES + PART_THREADS - 1) / PART_THREADS;   // tile sums per thread
            unsigned int v[TPT], run = 0;
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                const int t = tid * TPT + k;
                v[k] = (t < p.ntiles) ? tsC[t] : 0u;
                run += v[k];
            }
            unsigned int pre = block_excl_scan(run, s_warp, &ztotal);
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                if (tid * TPT + k <= PART_MAX_TILES) s_tpre[tid * TPT + k] = pre;
                pre += v[k];
            }
            __syncthreads();
        }

        // ---- P1: segment heads fold / retire / queue; publish Z ------------------------------------------
        if (blockIdx.x == 0 && tid == 0) p.status[SW_ANYLONG + ((l + 1) & 1)] = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            // striped ownership: element (k, tid) = tile base + k * 256 + tid, so that every load / store of
            // a warp covers 32 consecutive elements (coalesced) and a split sends them to two dense runs
            unsigned int zloc[PART_E];
            const int i0 = t * PART_TILE + tid;
            uint32_t kyv[PART_E];
            int sv[PART_E], ev[PART_E];
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = i0 + k * PART_THREADS;
                kyv[k] = 0; sv[k] = 0; ev[k] = SEG_DONE;
                if (i < m) {
                    ev[k] = (l == 0) ? m : eC[i];
                    sv[k] = (l == 0) ? 0 : sC[i];
                    kyv[k] = key[i];
                }
            }
            // what each element asks for: 1 = retire queue, 2 = medium queue, 3 = long queue (heads only)
            int cls[PART_E];
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = i0 + k * PART_THREADS;
                zloc[k] = 0;
                cls[k] = 0;
                if (!(ev[k] & SEG_DONE)) {
                    const int s = sv[k], e = ev[k];
                    const int len = e - s;
                    const uint32_t ky = kyv[k];
                    if (len <= RETIRE_SEG) {
                        if (i == s && classify) cls[k] = 1;   // retire_kernel finishes it on all SMs
                    } else {
                        if (i == s && classify && l < key_leaf_level(ky, two_n, depth)) {
                            if (len > HUGE_SEG) {
                                const unsigned int slot = atomicAdd(p.status + SW_QCOUNT_HUGE, 1u);
                                p.qhuge[slot] = seg_entry{l, s, len, ky >> sh};
                                p.huge_off[slot] = (int)atomicAdd(p.status + SW_HUGE_CHUNKS,
                                                                  (unsigned int)((len + FOLD_CHUNK_ELEMS - 1) / FOLD_CHUNK_ELEMS));
                            } else {
                                cls[k] = len > MEDIUM_SEG ? 3 : 2;
                            }
                        }
                        if (i == s) p.status[SW_ANYLONG + (l & 1)] = 1;   // another split is needed
                        if (!last) zloc[k] = ((ky >> (sh - 1)) & 1u) ^ 1u;
                    }
                }
            }
            // queue slots: ONE atomic per warp and queue for the warp's 256 elements (the counters are
            // single addresses; per-head atomics would serialise hundreds of thousands of them in L2)
            {
                const int ln = tid & 31;
                const unsigned int lt = (1u << ln) - 1u;
                unsigned int b1[PART_E], b2[PART_E], b3[PART_E];
                unsigned int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
                for (int k = 0; k < PART_E; ++k) {
                    b1[k] = __ballot_sync(0xffffffffu, cls[k] == 1);
                    b2[k] = __ballot_sync(0xffffffffu, cls[k] == 2);
                    b3[k] = __ballot_sync(0xffffffffu, cls[k] == 3);
                    c1 += __popc(b1[k]); c2 += __popc(b2[k]); c3 += __popc(b3[k]);
                }
                unsigned int q1 = 0, q2 = 0, q3 = 0;
                if (ln == 0) {
                    if (c1) q1 = atomicAdd(p.status + SW_QCOUNT_RET, c1);
                    if (c2) q2 = atomicAdd(p.status + SW_QCOUNT_MED, c2);
                    if (c3) q3 = atomicAdd(p.status + SW_QCOUNT, c3);
                }
                q1 = __shfl_sync(0xffffffffu, q1, 0);
                q2 = __shfl_sync(0xffffffffu, q2, 0);
                q3 = __shfl_sync(0xffffffffu, q3, 0);
#pragma unroll
                for (int k = 0; k < PART_E; ++k) {
                    const int len = ev[k] - sv[k];
                    if (cls[k] == 1) p.qret[q1 + __popc(b1[k] & lt)] = seg_entry{l, sv[k], len, 0u};
                    if (cls[k] == 2) p.qmed[q2 + __popc(b2[k] & lt)] = seg_entry{l, sv[k], len, kyv[k] >> sh};
                    if (cls[k] == 3) p.qlong[q3 + __popc(b3[k] & lt)] = seg_entry{l, sv[k], len, kyv[k] >> sh};
                    q1 += __popc(b1[k]); q2 += __popc(b2[k]); q3 += __popc(b3[k]);
                }
            }
            if (!last) {
                // exclusive zero count in element order (k-major, then thread): ballots give the in-warp
                // rank, one shared table of 8 x 8 warp totals gives the rest (one barrier pair per tile)
                const int lane = tid & 31, warp = tid >> 5;
                unsigned int bal[PART_E];
#pragma unroll
                for (int k = 0; k < PART_E; ++k) bal[k] = __ballot_sync(0xffffffffu, zloc[k] != 0);
                __syncthreads();   // s_wk free (previous tile)
                if (lane < PART_E) s_wk[lane * (PART_THREADS / 32) + warp] = __popc(bal[lane]);
                __syncthreads();
                // every warp scans the 64 totals itself: lane holds entries 2*lane, 2*lane+1
                const unsigned int e0 = s_wk[2 * lane], e1 = s_wk[2 * lane + 1];
                unsigned int inc = e0 + e1;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc += o;
                }
                const unsigned int excl_pair = inc - (e0 + e1);   // exclusive prefix of entry 2*lane
                const unsigned int tbase = s_tpre[t];
#pragma unroll
                for (int k = 0; k < PART_E; ++k) {
                    const int ent = k * (PART_THREADS / 32) + warp;
                    unsigned int pre = __shfl_sync(0xffffffffu, excl_pair, ent >> 1);
                    const unsigned int first = __shfl_sync(0xffffffffu, e0, ent >> 1);
                    if (ent & 1) pre += first;
                    const int i = i0 + k * PART_THREADS;
                    if (i < m) p.Z[i] = (int)(tbase + pre + __popc(bal[k] & ((1u << lane) - 1u)));
                }
                if (tid == 0) tsN[t] = 0;   // counters of the level after next start from zero
            }
        }
        if (last) break;
        PT_TICK(2 + 4 * l);
        grid.sync();
        PT_TICK(3 + 4 * l);
        if (((volatile unsigned int *)p.status)[SW_ANYLONG + (l & 1)] == 0) break;   // only retired segments are left (grid-uniform)

        // ---- P2: stable split of every long segment by the next key bit ------------------------------------
        // Loads are issued in batches (all keys / bounds, then all zero counts) so that one thread has
        // 8 independent L2 requests in flight per round trip instead of 8 dependent chains.
        const bool need_next = (l + 1 < depth - 1);   // level l+1 may split again by bit sh-2
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int i0 = t * PART_TILE + tid;
            uint32_t kyv[PART_E];
            int sv[PART_E], ev[PART_E];
            bool live[PART_E], moving[PART_E];
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = i0 + k * PART_THREADS;
                live[k] = i < m;
                kyv[k] = 0; sv[k] = 0; ev[k] = 0;
                if (live[k]) {
                    const int eraw = (l == 0) ? m : eC[i];
                    sv[k] = (l == 0) ? 0 : sC[i];
                    ev[k] = eraw;
                    kyv[k] = key[i];
                }
            }
            int zsv[PART_E], zev[PART_E], ziv[PART_E];
            T dv[PART_E];
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = i0 + k * PART_THREADS;
                const int e = ev[k] & ~SEG_DONE;
                moving[k] = live[k] && !(ev[k] & SEG_DONE) && (e - sv[k] > RETIRE_SEG);
                zsv[k] = zev[k] = ziv[k] = 0;
                dv[k] = T(0);
                if (moving[k]) {
                    zsv[k] = p.Z[sv[k]];
                    zev[k] = (e < m) ? p.Z[e] : (int)ztotal;
                    ziv[k] = p.Z[i];
                    dv[k] = dC[i];
                }
            }
            int npv[PART_E];
            unsigned int znext[PART_E];
#pragma unroll
            for (int k = 0; k < PART_E; ++k) {
                const int i = i0 + k * PART_THREADS;
                npv[k] = i;
                znext[k] = 0;
                if (live[k]) {
                    const int s = sv[k], e = ev[k] & ~SEG_DONE;
                    if (!moving[k]) {
                        // retired: stays where it is, nobody reads its key / difference again
                        sN[i] = s;
                        eN[i] = e | SEG_DONE;
                    } else {
                        const uint32_t ky = kyv[k];
                        const int zb = ziv[k] - zsv[k];      // zero-bit elements of my segment before me
                        const int nz = zev[k] - zsv[k];      // zero-bit elements of my segment
                        int np, s2, e2;
                        if (((ky >> (sh - 1)) & 1u) == 0) {
                            np = s + zb; s2 = s; e2 = s + nz;
                        } else {
                            np = s + nz + (i - s - zb); s2 = s + nz; e2 = e;
                        }
                        npv[k] = np;
                        keyN[np] = ky;
                        dN[np] = dv[k];
                        sN[np] = s2;
                        eN[np] = e2;
                        if (need_next && e2 - s2 > RETIRE_SEG) znext[k] = ((ky >> (sh - 2)) & 1u) ^ 1u;
                    }
                }
            }
            if (need_next) {
                // zero counts of the NEXT split per destination tile: shared-memory histogram first
#pragma unroll
                for (int k = 0; k < PART_E; ++k)
                    if (znext[k]) atomicAdd(&s_hist[npv[k] / PART_TILE], 1u);
            }
        }
        if (need_next) {
            __syncthreads();
            for (int j = tid; j < p.ntiles; j += PART_THREADS) {
                const unsigned int v = s_hist[j];
                if (v) {
                    atomicAdd(tsN + j, v);
                    s_hist[j] = 0;
                }
            }
        }
        PT_TICK(4 + 4 * l);
        grid.sync();
        PT_TICK(5 + 4 * l);
    }
}

}  // namespace st
