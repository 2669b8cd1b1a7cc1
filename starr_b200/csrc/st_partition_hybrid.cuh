// st_partition_hybrid.cuh -- the ordered partition of st_partition.cuh without grid barriers.
//
// Same contract as partition_levels_kernel (reference semantics: starr/src/_sumtree_array.pyx:41-49, every
// node is the ORDERED fold of the updates below it): for every level l the batch ordered by
// (ancestor at level l, batch position j) in lev_key[l] / lev_d[l], segments longer than RETIRE_SEG queued
// for the exact parallel folds, shorter ones queued once for retire_kernel.  What changes is how the order
// is produced:
//
//   top levels 1 .. LT (segments of ~m / 2^l updates, far larger than a CTA):  the order of level l is a
//     stable counting sort of the batch by the top l key bits, and ALL LT levels come out of ONE histogram:
//       hy_hist_kernel     per tile of 2048 updates, counts per level-LT node
//       hy_scan_kernel     per node, prefix over tiles; the last CTA derives every level's segment table
//                          (start, length, state) and queues the top segments
//       hy_scatter_kernel  per tile, LT stable splits in shared memory; after each one the tile's runs
//                          are stored to their global places in lev_key[l] / lev_d[l] (coalesced runs)
//   levels below LT (segments <= HY_CAP updates):  hy_segment_kernel, one CTA per level-LT segment, splits
//     in shared memory (__syncthreads only) until nothing longer than RETIRE_SEG is left, writing
//     lev_key / lev_d of the levels it passes and queueing the segments it creates.
//
// Three ordinary launches + one CTA-per-segment launch replace 12 levels x 2 grid-wide barriers.  Skewed
// batches (a level-LT segment longer than HY_CAP) set SW_NEED_COOP: hy_segment_kernel then leaves everything
// to partition_levels_kernel, which starts from level LT instead of level 0.
#pragma once
#include "st_partition.cuh"

namespace st {

constexpr int HY_TILE = 2048;
constexpr int HY_THREADS = 256;
constexpr int HY_E = HY_TILE / HY_THREADS;   // 8
constexpr int HY_MAX_LT = 10;
constexpr int HY_MAX_BINS = 1 << HY_MAX_LT;  // 1024
constexpr int HY_MAX_TILES = (int)(UPDATE_MAX_CHUNK / HY_TILE);   // 1024
#ifndef ST_HY_SEG
#define ST_HY_SEG 4096                       // level-LT segments of a uniform batch hold ST_HY_SEG/2+1 .. ST_HY_SEG updates
                                             // (2048 measured the same: partition 113 vs 117 us)
#endif
constexpr int HY_SEG = ST_HY_SEG;
constexpr int HY_CAP = HY_SEG + HY_SEG / 2;  // longest level-LT segment a CTA takes (uniform: <= HY_SEG + noise)
constexpr int HY_B_E = 6;
constexpr int HY_B_THREADS = HY_CAP / HY_B_E;   // 512 (HY_SEG 2048) or 1024 (HY_SEG 4096)
constexpr int HY_SEG_LIVE = 0, HY_SEG_DONE = 1;

template <typename T> struct hybrid_params {
    uint32_t n;
    int depth, m, ntiles, LT;
    const uint32_t *key0;   // keys in batch order
    uint32_t *lev_key;      // [level][lev_stride]
    T *lev_d;               // [level][lev_stride]; level 0 (batch order) filled by the caller
    int64_t lev_stride;
    unsigned int *hist;     // [bins][HY_MAX_TILES]: counts per (level-LT node, tile), then exclusive prefix over tiles
    unsigned int *bintotal; // [bins]
    int *segstart, *seglen; // [2 * bins], indexed by node = (1 << l) + c, l <= LT
    unsigned char *segstate;
    int *sC, *eC;           // segment bounds of level LT for the cooperative fallback
    seg_entry *qlong, *qmed, *qhuge, *qret;
    int *huge_off;
    unsigned int *status;
};

// level-LT node (0 .. bins-1) of a key
__device__ __forceinline__ uint32_t hy_bin(uint32_t key, int depth, int LT) { return (key >> (depth - LT)) & ((1u << LT) - 1u); }

__global__ void __launch_bounds__(HY_THREADS) hy_hist_kernel(const uint32_t *__restrict__ key0, int m, int depth, int LT,
                                                             unsigned int *__restrict__ hist, unsigned int *status)
{
    if (status[SW_STATUS]) return;
    __shared__ unsigned int h[HY_MAX_BINS];
    const int bins = 1 << LT, tid = threadIdx.x, t = blockIdx.x;
    for (int b = tid; b < bins; b += HY_THREADS) h[b] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < HY_E; ++k) {
        const int i = t * HY_TILE + k * HY_THREADS + tid;
        if (i < m) atomicAdd(&h[hy_bin(key0[i], depth, LT)], 1u);
    }
    __syncthreads();
    for (int b = tid; b < bins; b += HY_THREADS) hist[(size_t)b * HY_MAX_TILES + t] = h[b];
    if (t == 0 && tid == 0) { status[SW_NEED_COOP] = 0; status[SW_HY_TICKET] = 0; status[SW_HY_KEYLEVELS] = 0; }
}

// CTA b: exclusive prefix over the tiles of node b.  The last CTA to finish turns the node totals into the
// segment table of every level <= LT and classifies those segments (what P1 of partition_levels_kernel does
// for the top levels).
template <typename T>
__global__ void __launch_bounds__(HY_THREADS) hy_scan_kernel(hybrid_params<T> p)
{
    if (p.status[SW_STATUS]) return;
    __shared__ unsigned int s_warp[HY_THREADS / 32];
    __shared__ int s_start[HY_MAX_BINS + 1];
    __shared__ int s_len[2 * HY_MAX_BINS];
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x, b = blockIdx.x, bins = 1 << p.LT;
    constexpr int TPT = HY_MAX_TILES / HY_THREADS;   // 4 tiles per thread
    unsigned int v[TPT], run = 0;
    unsigned int *row = p.hist + (size_t)b * HY_MAX_TILES;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int t = tid * TPT + k;
        v[k] = t < p.ntiles ? row[t] : 0u;
        run += v[k];
    }
    unsigned int total;
    unsigned int pre = block_excl_scan(run, s_warp, &total);
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int t = tid * TPT + k;
        if (t < p.ntiles) row[t] = pre;
        pre += v[k];
    }
    if (tid == 0) {
        p.bintotal[b] = total;
        __threadfence();
        s_ticket = atomicAdd(p.status + SW_HY_TICKET, 1u);
    }
    __syncthreads();
    if (s_ticket != (unsigned int)bins - 1u) return;
    __threadfence();
    // ---- last CTA: segment table ----------------------------------------------------------------------
    unsigned int tv[HY_MAX_BINS / HY_THREADS], trun = 0;   // 2 consecutive nodes per thread
#pragma unroll
    for (int k = 0; k < HY_MAX_BINS / HY_THREADS; ++k) {
        const int c = tid * (HY_MAX_BINS / HY_THREADS) + k;
        tv[k] = c < bins ? __ldcg(p.bintotal + c) : 0u;
        trun += tv[k];
    }
    unsigned int tpre = block_excl_scan(trun, s_warp, &total);
#pragma unroll
    for (int k = 0; k < HY_MAX_BINS / HY_THREADS; ++k) {
        const int c = tid * (HY_MAX_BINS / HY_THREADS) + k;
        if (c < bins) s_start[c] = (int)tpre;
        tpre += tv[k];
    }
    if (tid == 0) s_start[bins] = p.m;
    __syncthreads();
    for (int node = 1 + tid; node < 2 * bins; node += HY_THREADS) {
        const int l = heap_level((uint32_t)node), c = node - (1 << l), w = p.LT - l;
        const int s = s_start[c << w], e = s_start[(c + 1) << w];
        s_len[node] = e - s;
        p.segstart[node] = s;
        p.seglen[node] = e - s;
    }
    __syncthreads();
    for (int node = 1 + tid; node < 2 * bins; node += HY_THREADS) {
        const int l = heap_level((uint32_t)node);
        const int len = s_len[node], s = p.segstart[node];
        bool above = false;                          // an ancestor segment was short enough to be retired as a whole
        for (int a = node >> 1; a >= 1; a >>= 1) above |= s_len[a] <= RETIRE_SEG;
        int state = HY_SEG_DONE;
        if (!above && len > 0) {
            if (len <= RETIRE_SEG) {
                p.qret[atomicAdd(p.status + SW_QCOUNT_RET, 1u)] = seg_entry{l, s, len, 0u};
                atomicOr(p.status + SW_HY_KEYLEVELS, 1u << l);   // retire_kernel reads lev_key[l] for this segment
            } else {
                state = HY_SEG_LIVE;                 // l <= LT < leaf level: `node` is a proper ancestor
                if (len > HUGE_SEG) {
                    const unsigned int slot = atomicAdd(p.status + SW_QCOUNT_HUGE, 1u);
                    p.qhuge[slot] = seg_entry{l, s, len, (uint32_t)node};
                    p.huge_off[slot] = (int)atomicAdd(p.status + SW_HUGE_CHUNKS,
                                                      (unsigned int)((len + FOLD_CHUNK_ELEMS - 1) / FOLD_CHUNK_ELEMS));
                } else if (len > MEDIUM_SEG) {
                    p.qlong[atomicAdd(p.status + SW_QCOUNT, 1u)] = seg_entry{l, s, len, (uint32_t)node};
                } else {
                    p.qmed[atomicAdd(p.status + SW_QCOUNT_MED, 1u)] = seg_entry{l, s, len, (uint32_t)node};
                }
                if (l == p.LT && len > HY_CAP) p.status[SW_NEED_COOP] = 1u;
            }
        }
        p.segstate[node] = (unsigned char)state;
    }
}

// exclusive count of flagged elements in element order (i = k * THREADS + tid, k-major) over one CTA.
// bal[k] = warp ballot of element k's flag.  Returns through tab (shared, E * NW entries) the exclusive
// prefix of every (k, warp) cell; *total = number of flagged elements.
template <int E, int THREADS>
__device__ __forceinline__ void cta_flag_scan(const unsigned int (&bal)[E], unsigned int *tab, unsigned int *total_out)
{
    constexpr int NW = THREADS / 32, N = E * NW, PER = (N + 31) / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // tab free
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < E; ++k) tab[k * NW + warp] = __popc(bal[k]);
    }
    __syncthreads();
    if (warp == 0) {
        unsigned int loc[PER], sum = 0;
#pragma unroll
        for (int x = 0; x < PER; ++x) {
            const int idx = lane * PER + x;
            const unsigned int v = idx < N ? tab[idx] : 0u;
            loc[x] = sum;
            sum += v;
        }
        unsigned int inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        const unsigned int excl = inc - sum;
#pragma unroll
        for (int x = 0; x < PER; ++x) {
            const int idx = lane * PER + x;
            if (idx < N) tab[idx] = excl + loc[x];
        }
        if (lane == 31) *total_out = inc;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(HY_THREADS) hy_scatter_kernel(hybrid_params<T> p)
{
    if (p.status[SW_STATUS]) return;
    __shared__ uint32_t s_key[HY_TILE];
    __shared__ T s_d[HY_TILE];
    __shared__ unsigned short s_Z[HY_TILE];
    __shared__ unsigned int s_cnt[2 * HY_MAX_BINS];    // tile-local count per node, then tile-local run start
    __shared__ unsigned int s_base[2 * HY_MAX_BINS];   // global position of the tile's first element of the node
    __shared__ unsigned int s_tab[HY_E * (HY_THREADS / 32)];
    __shared__ unsigned int s_total;
    const int tid = threadIdx.x, t = blockIdx.x, LT = p.LT, bins = 1 << LT, depth = p.depth;
    const int lane = tid & 31, warp = tid >> 5;
    const int mt = p.m - t * HY_TILE < HY_TILE ? p.m - t * HY_TILE : HY_TILE;
    const bool need_coop = p.status[SW_NEED_COOP] != 0;
    // keys of a top level are only read where a segment is retired (skewed batches) and at level LT
    const unsigned int key_levels = p.status[SW_HY_KEYLEVELS] | (1u << LT);

    for (int b = tid; b < 2 * bins; b += HY_THREADS) s_cnt[b] = 0;
    __syncthreads();
    uint32_t ky[HY_E];
    T dv[HY_E];
#pragma unroll
    for (int k = 0; k < HY_E; ++k) {
        const int i = k * HY_THREADS + tid;
        ky[k] = 0; dv[k] = T(0);
        if (i < mt) {
            ky[k] = p.key0[t * HY_TILE + i];
            dv[k] = p.lev_d[t * HY_TILE + i];
            atomicAdd(&s_cnt[bins + hy_bin(ky[k], depth, LT)], 1u);
        }
    }
    for (int b = tid; b < bins; b += HY_THREADS) s_base[bins + b] = p.hist[(size_t)b * HY_MAX_TILES + t];
    __syncthreads();
    for (int l = LT - 1; l >= 0; --l) {      // counts and cross-tile prefixes of the coarser nodes
        for (int c = tid; c < (1 << l); c += HY_THREADS) {
            const int node = (1 << l) + c;
            s_cnt[node] = s_cnt[2 * node] + s_cnt[2 * node + 1];
            s_base[node] = s_base[2 * node] + s_base[2 * node + 1];
        }
        __syncthreads();
    }
    for (int node = 1 + tid; node < 2 * bins; node += HY_THREADS) s_base[node] += (unsigned int)p.segstart[node];
    // tile-local run starts, top-down: children split the parent's run
    unsigned int carry0 = 0, carry1 = 0;
    (void)carry0; (void)carry1;
    __syncthreads();
    if (tid == 0) s_total = s_cnt[1];   // == mt
    for (int l = 0; l < LT; ++l) {
        // s_cnt[node] of level l currently holds: level 0 -> count (start is 0); deeper -> already a start
        for (int c = tid; c < (1 << l); c += HY_THREADS) {
            const int node = (1 << l) + c;
            const unsigned int start = (l == 0) ? 0u : s_cnt[node];
            const unsigned int c0 = s_cnt[2 * node];
            if (l == 0) s_cnt[node] = 0;
            s_cnt[2 * node] = start;
            s_cnt[2 * node + 1] = start + c0;
        }
        __syncthreads();
    }
    // s_cnt[node] = tile-local start of node's run for every level now

    for (int l = 1; l <= LT; ++l) {
        const int bit = depth - l;                 // node at level l = key >> (depth - l)
        unsigned int bal[HY_E];
        bool zero[HY_E];
#pragma unroll
        for (int k = 0; k < HY_E; ++k) {
            const int i = k * HY_THREADS + tid;
            zero[k] = i < mt && ((ky[k] >> bit) & 1u) == 0u;
            bal[k] = __ballot_sync(0xffffffffu, zero[k]);
        }
        unsigned int ztotal;
        cta_flag_scan<HY_E, HY_THREADS>(bal, s_tab, &ztotal);
        unsigned int zi[HY_E];
#pragma unroll
        for (int k = 0; k < HY_E; ++k) {
            const int i = k * HY_THREADS + tid;
            zi[k] = s_tab[k * (HY_THREADS / 32) + warp] + __popc(bal[k] & ((1u << lane) - 1u));
            if (i < mt) s_Z[i] = (unsigned short)zi[k];
        }
        __syncthreads();
        int np[HY_E];
#pragma unroll
        for (int k = 0; k < HY_E; ++k) {
            const int i = k * HY_THREADS + tid;
            np[k] = i;
            if (i < mt) {
                const uint32_t parent = ky[k] >> (bit + 1);          // node at level l-1 (heap numbering)
                const int rs = (int)s_cnt[parent];                    // start of the parent's run in the tile
                const int zb = (int)zi[k] - (int)s_Z[rs];
                np[k] = zero[k] ? (int)s_cnt[2 * parent] + zb : (int)s_cnt[2 * parent + 1] + (i - rs - zb);
            }
        }
        __syncthreads();   // everybody has read s_Z
#pragma unroll
        for (int k = 0; k < HY_E; ++k) {
            const int i = k * HY_THREADS + tid;
            if (i < mt) { s_key[np[k]] = ky[k]; s_d[np[k]] = dv[k]; }
        }
        __syncthreads();
        uint32_t *gk = p.lev_key + (int64_t)l * p.lev_stride;
        T *gd = p.lev_d + (int64_t)l * p.lev_stride;
#pragma unroll
        for (int k = 0; k < HY_E; ++k) {
            const int i = k * HY_THREADS + tid;
            if (i < mt) {
                ky[k] = s_key[i];
                dv[k] = s_d[i];
                const uint32_t node = ky[k] >> bit;
                const unsigned int g = s_base[node] + ((unsigned int)i - s_cnt[node]);
                if ((key_levels >> l) & 1u) gk[g] = ky[k];
                gd[g] = dv[k];
                if (l == LT && need_coop) {
                    const int s = p.segstart[node];
                    p.sC[g] = s;
                    p.eC[g] = (s + p.seglen[node]) | (p.segstate[node] == HY_SEG_LIVE ? 0 : SEG_DONE);
                }
            }
        }
        // the next level reads s_Z / s_key again only after its own barriers (cta_flag_scan starts with one)
    }
}

// One CTA per live level-LT segment: the rest of the partition in shared memory.
// (48 registers: a CTA of 1024 threads must leave room on its SM for the side stream's kernels -- with the whole
// register file it would wait for deep_red_kernel to drain, which serialised the two streams)
template <typename T>
__global__ void __maxnreg__(48) hy_segment_kernel(hybrid_params<T> p)   // launched with HY_B_THREADS threads
{
    if (p.status[SW_STATUS] || p.status[SW_NEED_COOP]) return;
    const int node0 = (1 << p.LT) + blockIdx.x;
    if (p.segstate[node0] != HY_SEG_LIVE) return;
    const int g0 = p.segstart[node0], len = p.seglen[node0];     // RETIRE_SEG < len <= HY_CAP
    extern __shared__ __align__(16) unsigned char hy_smem[];
    T *s_d = reinterpret_cast<T *>(hy_smem);
    uint32_t *s_key = reinterpret_cast<uint32_t *>(hy_smem + (size_t)HY_CAP * sizeof(T));
    unsigned short *s_s = reinterpret_cast<unsigned short *>(s_key + HY_CAP);
    unsigned short *s_e = s_s + HY_CAP;
    unsigned short *s_Z = s_e + HY_CAP;
    __shared__ unsigned int s_tab[HY_B_E * (HY_B_THREADS / 32)];
    __shared__ unsigned int s_total;
    constexpr unsigned short DONE16 = 0x8000u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, depth = p.depth;
    const uint64_t two_n = 2ull * p.n;
    {
        const uint32_t *gk = p.lev_key + (int64_t)p.LT * p.lev_stride + g0;
        const T *gd = p.lev_d + (int64_t)p.LT * p.lev_stride + g0;
        for (int i = tid; i < len; i += HY_B_THREADS) {
            s_key[i] = gk[i];
            s_d[i] = gd[i];
            s_s[i] = 0;
            s_e[i] = (unsigned short)len;
        }
    }
    __syncthreads();
    for (int l = p.LT; l < depth; ++l) {
        const bool last = (l == depth - 1);
        const int sh = depth - l;
        uint32_t ky[HY_B_E];
        int sv[HY_B_E], ev[HY_B_E];
        bool live[HY_B_E];
        int cls[HY_B_E];
        unsigned int bal[HY_B_E];
        bool any = false;
#pragma unroll
        for (int k = 0; k < HY_B_E; ++k) {
            const int i = k * HY_B_THREADS + tid;
            ky[k] = 0; sv[k] = 0; ev[k] = 0; live[k] = false; cls[k] = 0;
            if (i < len) {
                ky[k] = s_key[i];
                sv[k] = s_s[i];
                const unsigned short e = s_e[i];
                ev[k] = e & ~DONE16;
                if (!(e & DONE16)) {
                    const int sl = ev[k] - sv[k];
                    live[k] = sl > RETIRE_SEG;
                    if (l > p.LT && i == sv[k]) {       // segment head: classify (level LT was classified by hy_scan_kernel)
                        if (!live[k]) cls[k] = 1;
                        else if (l < key_leaf_level(ky[k], two_n, depth)) cls[k] = sl > MEDIUM_SEG ? 3 : 2;
                    }
                }
            }
            any |= live[k];
            bal[k] = __ballot_sync(0xffffffffu, live[k] && !last && (((ky[k] >> (sh - 1)) & 1u) == 0u));
        }
        // queue slots: one atomic per warp and queue
        {
            const unsigned int lt = (1u << lane) - 1u;
            unsigned int b1[HY_B_E], b2[HY_B_E], b3[HY_B_E], c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
            for (int k = 0; k < HY_B_E; ++k) {
                b1[k] = __ballot_sync(0xffffffffu, cls[k] == 1);
                b2[k] = __ballot_sync(0xffffffffu, cls[k] == 2);
                b3[k] = __ballot_sync(0xffffffffu, cls[k] == 3);
                c1 += __popc(b1[k]); c2 += __popc(b2[k]); c3 += __popc(b3[k]);
            }
            unsigned int q1 = 0, q2 = 0, q3 = 0;
            if (lane == 0) {
                if (c1) q1 = atomicAdd(p.status + SW_QCOUNT_RET, c1);
                if (c2) q2 = atomicAdd(p.status + SW_QCOUNT_MED, c2);
                if (c3) q3 = atomicAdd(p.status + SW_QCOUNT, c3);
            }
            q1 = __shfl_sync(0xffffffffu, q1, 0);
            q2 = __shfl_sync(0xffffffffu, q2, 0);
            q3 = __shfl_sync(0xffffffffu, q3, 0);
#pragma unroll
            for (int k = 0; k < HY_B_E; ++k) {
                const int sl = ev[k] - sv[k];
                if (cls[k] == 1) p.qret[q1 + __popc(b1[k] & lt)] = seg_entry{l, g0 + sv[k], sl, 0u};
                if (cls[k] == 2) p.qmed[q2 + __popc(b2[k] & lt)] = seg_entry{l, g0 + sv[k], sl, ky[k] >> sh};
                if (cls[k] == 3) p.qlong[q3 + __popc(b3[k] & lt)] = seg_entry{l, g0 + sv[k], sl, ky[k] >> sh};
                q1 += __popc(b1[k]); q2 += __popc(b2[k]); q3 += __popc(b3[k]);
            }
        }
        if (!__syncthreads_or(any ? 1 : 0) || last) break;     // only retired segments are left
        // ---- stable split of every live segment by the next key bit ----
        unsigned int ztotal;
        cta_flag_scan<HY_B_E, HY_B_THREADS>(bal, s_tab, &s_total);
        ztotal = s_total;
        unsigned int zi[HY_B_E];
#pragma unroll
        for (int k = 0; k < HY_B_E; ++k) {
            const int i = k * HY_B_THREADS + tid;
            zi[k] = s_tab[k * (HY_B_THREADS / 32) + warp] + __popc(bal[k] & ((1u << lane) - 1u));
            if (i < len) s_Z[i] = (unsigned short)zi[k];
        }
        __syncthreads();
        int np[HY_B_E], s2[HY_B_E], e2[HY_B_E];
        T dv[HY_B_E];
#pragma unroll
        for (int k = 0; k < HY_B_E; ++k) {
            const int i = k * HY_B_THREADS + tid;
            np[k] = i; s2[k] = 0; e2[k] = 0; dv[k] = T(0);
            if (live[k]) {
                const int s = sv[k], e = ev[k];
                const int zs = s_Z[s], ze = e < len ? (int)s_Z[e] : (int)ztotal;
                const int zb = (int)zi[k] - zs, nz = ze - zs;
                if (((ky[k] >> (sh - 1)) & 1u) == 0u) { np[k] = s + zb; s2[k] = s; e2[k] = s + nz; }
                else { np[k] = s + nz + (i - s - zb); s2[k] = s + nz; e2[k] = e; }
                dv[k] = s_d[i];
            }
        }
        __syncthreads();   // all reads of the old order are done
        uint32_t *gk = p.lev_key + (int64_t)(l + 1) * p.lev_stride + g0;
        T *gd = p.lev_d + (int64_t)(l + 1) * p.lev_stride + g0;
#pragma unroll
        for (int k = 0; k < HY_B_E; ++k) {
            const int i = k * HY_B_THREADS + tid;
            if (live[k]) {
                s_key[np[k]] = ky[k];
                s_d[np[k]] = dv[k];
                s_s[np[k]] = (unsigned short)s2[k];
                s_e[np[k]] = (unsigned short)e2[k];
            } else if (i < len) {
                s_e[i] = (unsigned short)(ev[k] | DONE16);   // retired: stays where it is, nobody reads it again
            }
        }
        __syncthreads();
        // level l+1 order of the elements that moved -> global (coalesced), for retire_kernel and the folds
#pragma unroll
        for (int k = 0; k < HY_B_E; ++k) {
            const int i = k * HY_B_THREADS + tid;
            if (i < len && !(s_e[i] & DONE16)) {
                gk[i] = s_key[i];
                gd[i] = s_d[i];
            }
        }
    }
}

template <typename T> static inline size_t hy_segment_smem() { return (size_t)HY_CAP * (sizeof(T) + 4 + 2 + 2 + 2); }

// number of top levels for a batch of m updates: level-LT segments of a uniform batch hold HY_SEG/2+1 .. HY_SEG updates
static inline int hy_top_levels(int m)
{
    int LT = 1;
    while (LT < HY_MAX_LT && (m >> LT) > HY_SEG) ++LT;
    return LT;
}

}  // namespace st
