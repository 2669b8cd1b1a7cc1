// st_update_small.cuh -- the float update for small batches (<= 4096 updates, the PER step of
// BASELINE config #2) in ONE CTA: validation, key build, block radix sort, leaf pass and the whole
// top-down ordered partition run out of shared memory with __syncthreads instead of grid barriers
// and extra launches.  Same arithmetic, same order, same results as the multi-CTA pipeline
// (st_update.cu / st_partition.cuh): segments up to SMALL_RETIRE_SEG are queued for retirement, longer ones are
// queued with their level's ordered differences for the exact parallel ordered fold.
#pragma once
#include <cub/block/block_radix_sort.cuh>

#include "st_common.cuh"
#include "st_update_defs.cuh"

namespace st {

constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_M = SMALL_THREADS * 4;   // 4096: the largest batch of the single-CTA path (4 per thread)
#ifndef ST_SMALL_RETIRE_SEG
#define ST_SMALL_RETIRE_SEG 64
#endif
constexpr int SMALL_RETIRE_SEG = ST_SMALL_RETIRE_SEG;             // small batches: up to this length retired with ordered atomics,
                                                  // longer segments go to the CTA fold

template <typename T> struct small_params {
    T *H;
    uint32_t n;
    int depth;
    int m;
    const int64_t *idx;
    const T *vals;
    uint64_t nv;
    uint64_t j0;
    const T *deltas_in;  // nullable: apply-deltas mode
    T *delta_out;        // nullable
    T *lev_d;
    int64_t lev_stride;
    seg_entry *qlong, *qmed;
    unsigned int *status;
    bool assign;           // leaf = val instead of fl(leaf + d)
    seg_entry *qret;       // retirements for the follow-up kernel
    uint32_t *key0_out;    // level-0 keys (batch order) for the follow-up kernel
    uint32_t *lev_key;     // [level][lev_stride] keys of levels with retirements
};

template <typename T, int SMALL_E> struct small_smem {
    static constexpr int CAP = SMALL_THREADS * SMALL_E;
    uint32_t key[2][CAP];
    T d[2][CAP];
    uint32_t se[2][CAP];   // s | e << 16  (e == SMALL_M is stored as 0xffff + 1 -> use 13-bit fields)
    uint32_t z[CAP + 1];
    uint32_t wtot[SMALL_THREADS / 32];
    int flag;
    int flag_retire;
    unsigned int nret;
};

template <int SMALL_E> using SmallSort = cub::BlockRadixSort<uint32_t, SMALL_THREADS, SMALL_E, int>;

template <typename T, int SMALL_E> constexpr size_t small_smem_bytes()
{
    return sizeof(small_smem<T, SMALL_E>) > sizeof(typename SmallSort<SMALL_E>::TempStorage)
               ? sizeof(small_smem<T, SMALL_E>)
               : sizeof(typename SmallSort<SMALL_E>::TempStorage);
}

__device__ __forceinline__ uint32_t pack_se(int s, int e) { return (uint32_t)s | ((uint32_t)e << 16); }

// exclusive scan over the 1024-thread CTA; *total valid in every thread
__device__ __forceinline__ uint32_t small_block_scan(uint32_t v, uint32_t *s_wtot, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    __syncthreads();
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    // every warp scans the 32 warp totals itself
    const uint32_t wt = s_wtot[lane];
    uint32_t winc = wt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, winc, off);
        if (lane >= off) winc += o;
    }
    *total = __shfl_sync(0xffffffffu, winc, 31);
    const uint32_t wpre = __shfl_sync(0xffffffffu, winc - wt, warp);
    return wpre + inc - v;
}

// SMALL_E = updates per thread: 1 (batches <= 1024, a quarter of the instructions) or 4 (<= 4096)
template <typename T, int SMALL_E>
__global__ void __launch_bounds__(SMALL_THREADS, 1) update_small_kernel(small_params<T> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    small_smem<T, SMALL_E> &S = *reinterpret_cast<small_smem<T, SMALL_E> *>(smem_raw);
    using Sort = SmallSort<SMALL_E>;
    typename Sort::TempStorage &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(smem_raw);
    const int tid = threadIdx.x;
    const int m = p.m, depth = p.depth;
    const uint64_t two_n = 2ull * p.n;
    {
        // a new batch: an earlier batch's failure moves to the sticky word (st_poll_status still reports it)
        // and no longer blocks this one.  One CTA, every thread reads before the barrier: no race.
        const unsigned int prev = p.status[SW_STATUS];
        __syncthreads();
        if (tid == 0 && prev) {
            p.status[ST_SW_STICKY] |= prev;
            p.status[SW_STATUS] = 0;
        }
    }
#ifdef ST_SMALL_TIMING
    long long t_start = clock64();
    #define ST_TICK(k) if (tid == 0) p.status[32 + (k)] = (unsigned int)(clock64() - t_start)
#else
    #define ST_TICK(k)
#endif

    // ---- validation (pyx:37-39 negative scan over ALL vals; index range) ------------------------------
    unsigned int bad = 0;
    if (!p.deltas_in) {
        if constexpr (is_signed_t<T>::value)
            for (uint64_t i = tid; i < p.nv; i += SMALL_THREADS)
                if (p.vals[i] < T(0)) bad |= STF_NEG_VALS;
    }
    uint32_t keys[SMALL_E];
    int js[SMALL_E];
#pragma unroll
    for (int k = 0; k < SMALL_E; ++k) {
        const int j = tid * SMALL_E + k;
        keys[k] = 0xffffffffu;
        js[k] = j;
        if (j < m) {
            const int64_t l = p.idx[j];
            if (l < 0 || l >= (int64_t)p.n) {
                bad |= STF_INDEX;
            } else {
                const uint32_t h = (uint32_t)l + p.n;
                keys[k] = h << (depth - heap_level(h));
            }
        }
    }
    if (__syncthreads_or(bad != 0)) {
        if (bad) atomicOr(p.status + SW_STATUS, bad);
        return;
    }
    if (tid == 0) {
        p.status[SW_QCOUNT] = 0; p.status[SW_QTAKE] = 0;
        p.status[SW_QCOUNT_MED] = 0; p.status[SW_QTAKE_MED] = 0;
        p.status[SW_QCOUNT_HUGE] = 0; p.status[SW_HUGE_CHUNKS] = 0; p.status[SW_QTAKE_HUGE] = 0;
        p.status[SW_RETIRE_MASK] = 0;
        p.status[SW_QCOUNT_RET] = 0;
    }

    ST_TICK(0);
    // ---- stable sort by key (duplicates of one leaf adjacent, in batch order) -------------------------------
    uint32_t key0[SMALL_E];
#pragma unroll
    for (int k = 0; k < SMALL_E; ++k) key0[k] = keys[k];
    Sort(sort_tmp).Sort(keys, js, 0, depth + 1);
    __syncthreads();   // sort_tmp dead; the arrays below alias it
#pragma unroll
    for (int k = 0; k < SMALL_E; ++k) {
        const int i = tid * SMALL_E + k;
        S.key[1][i] = keys[k];          // sorted keys
        S.se[1][i] = (uint32_t)js[k];   // sorted batch positions
        S.key[0][i] = key0[k];          // level-0 (batch) order
        S.se[0][i] = pack_se(0, m);
    }
    __syncthreads();

    ST_TICK(1);
    // ---- leaf pass (pyx:43-45).  All global loads are issued up front (one round trip): every
    // thread fetches the new values of its 4 sorted positions into shared memory, run heads fetch
    // their leaf; then the heads walk their runs out of shared memory. ------------------------------------
    {
        T vreg[SMALL_E];
        T hreg[SMALL_E];
        bool head[SMALL_E];
#pragma unroll
        for (int k = 0; k < SMALL_E; ++k) {
            const int i = tid * SMALL_E + k;
            head[k] = false;
            vreg[k] = T(0);
            hreg[k] = T(0);
            if (i < m) {
                const uint32_t ky = S.key[1][i];
                const int j = (int)S.se[1][i];
                if (p.deltas_in) {
                    vreg[k] = p.deltas_in[j];
                } else {
                    const uint64_t g = p.j0 + (uint64_t)j;
                    vreg[k] = p.vals[g < p.nv ? g : g % p.nv];
                    head[k] = (i == 0 || S.key[1][i - 1] != ky);
                    if (head[k]) hreg[k] = p.H[key_to_heap(ky, two_n)];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < SMALL_E; ++k) {
            const int i = tid * SMALL_E + k;
            if (i < m) {
                if (p.deltas_in) S.d[0][(int)S.se[1][i]] = vreg[k];   // batch order
                else S.d[1][i] = vreg[k];                               // new value per sorted position
            }
        }
        __syncthreads();
        if (!p.deltas_in) {
#pragma unroll
            for (int k = 0; k < SMALL_E; ++k) {
                const int i = tid * SMALL_E + k;
                if (head[k]) {
                    const uint32_t ky = S.key[1][i];
                    T cur = hreg[k];
                    int e = i;
                    unsigned int nzero = 0;
                    do {
                        const T d = t_sub<T>(S.d[1][e], cur);
                        cur = p.assign ? S.d[1][e] : t_add<T>(cur, d);
                        S.d[0][(int)S.se[1][e]] = d;
                        nzero += (d == T(0));
                        ++e;
                    } while (e < m && S.key[1][e] == ky);
                    p.H[key_to_heap(ky, two_n)] = cur;
                    if (nzero) atomicAdd(p.status + SW_STAT_ZERO_DELTA, nzero);
                }
            }
        }
    }
    __syncthreads();
    if (p.delta_out)
        for (int j = tid; j < m; j += SMALL_THREADS) p.delta_out[j] = S.d[0][j];

    ST_TICK(2);
    // ---- top-down ordered partition; a segment of <= SMALL_RETIRE_SEG updates is queued for retirement
    // (finished for all remaining levels by the follow-up kernel) and the loop ends as soon as no longer
    // segment is left ------------------------------------------------------------------------------------
    constexpr uint32_t DONE = 0x80000000u;
    if (tid == 0) S.nret = 0;
    for (int l = 0; l < depth; ++l) {
        const int c = l & 1;
        const int sh = depth - l;
        const bool last = (l == depth - 1);
        if (l < 24) { ST_TICK(3 + l); }
        if (tid == 0) { S.flag = 0; S.flag_retire = 0; }
        __syncthreads();
        uint32_t zloc[SMALL_E], zsum = 0;
        uint32_t kreg[SMALL_E], sereg[SMALL_E];
        if constexpr (SMALL_E == 4) {
            // one 128-bit shared load per array instead of four conflicting 32-bit ones
            const uint4 k4 = *reinterpret_cast<const uint4 *>(&S.key[c][tid * SMALL_E]);
            const uint4 s4 = *reinterpret_cast<const uint4 *>(&S.se[c][tid * SMALL_E]);
            kreg[0] = k4.x; kreg[1] = k4.y; kreg[2] = k4.z; kreg[3] = k4.w;
            sereg[0] = s4.x; sereg[1] = s4.y; sereg[2] = s4.z; sereg[3] = s4.w;
        } else {
#pragma unroll
            for (int k = 0; k < SMALL_E; ++k) {
                kreg[k] = S.key[c][tid * SMALL_E + k];
                sereg[k] = S.se[c][tid * SMALL_E + k];
            }
        }
        int anylong = 0;
#pragma unroll
        for (int k = 0; k < SMALL_E; ++k) {
            const int i = tid * SMALL_E + k;
            zloc[k] = 0;
            if (i < m && !(sereg[k] & DONE)) {
                const uint32_t ky = kreg[k];
                const int s = (int)(sereg[k] & 0xffffu), e = (int)((sereg[k] >> 16) & 0x7fffu);
                const int len = e - s;
                if (len <= SMALL_RETIRE_SEG) {
                    // retired here; the adds themselves are issued by the follow-up kernel so that the
                    // random node updates spread over all SMs instead of queueing up on this one
                    if (i == s) {
                        S.flag_retire = 1;
                        p.qret[atomicAdd(&S.nret, 1u)] = seg_entry{l, s, len, 0u};
                    }
                } else {
                    anylong = 1;
                    if (i == s && l < key_leaf_level(ky, two_n, depth)) {
                        const uint32_t node = ky >> sh;
                        S.flag = 1;   // this level's ordered differences are needed by the fold kernel
                        p.qlong[atomicAdd(p.status + SW_QCOUNT, 1u)] = seg_entry{l, s, len, node};
                    }
                    if (!last) zloc[k] = ((ky >> (sh - 1)) & 1u) ^ 1u;
                }
            }
            zsum += zloc[k];
        }
        const int more = __syncthreads_or(anylong);   // also settles S.flag / S.flag_retire
        if (S.flag || S.flag_retire) {
            T *dst = p.lev_d + (int64_t)l * p.lev_stride;
            for (int i = tid; i < m; i += SMALL_THREADS) dst[i] = S.d[c][i];
        }
        if (S.flag_retire) {
            uint32_t *kd = (l == 0) ? p.key0_out : p.lev_key + (int64_t)l * p.lev_stride;
            for (int i = tid; i < m; i += SMALL_THREADS) kd[i] = S.key[c][i];
        }
        if (!more) break;    // only retired segments are left
        uint32_t ztotal = 0;
        uint32_t pre = small_block_scan(zsum, S.wtot, &ztotal);
        if (last) break;
        uint32_t zown[SMALL_E];
#pragma unroll
        for (int k = 0; k < SMALL_E; ++k) {
            zown[k] = pre;
            pre += zloc[k];
        }
        if constexpr (SMALL_E == 4) {
            *reinterpret_cast<uint4 *>(&S.z[tid * SMALL_E]) = make_uint4(zown[0], zown[1], zown[2], zown[3]);
        } else {
#pragma unroll
            for (int k = 0; k < SMALL_E; ++k) S.z[tid * SMALL_E + k] = zown[k];
        }
        T dreg[SMALL_E];
        if constexpr (sizeof(T) == 4 && SMALL_E == 4) {
            const float4 d4 = *reinterpret_cast<const float4 *>(&S.d[c][tid * SMALL_E]);
            dreg[0] = d4.x; dreg[1] = d4.y; dreg[2] = d4.z; dreg[3] = d4.w;
        } else {
#pragma unroll
            for (int k = 0; k < SMALL_E; ++k) dreg[k] = S.d[c][tid * SMALL_E + k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SMALL_E; ++k) {
            const int i = tid * SMALL_E + k;
            if (i < m) {
                const int s = (int)(sereg[k] & 0xffffu), e = (int)((sereg[k] >> 16) & 0x7fffu);
                if ((sereg[k] & DONE) || e - s <= SMALL_RETIRE_SEG) {
                    S.se[c ^ 1][i] = pack_se(s, e) | DONE;   // retired: stays put, never read again
                } else {
                    const uint32_t ky = kreg[k];
                    const int zs = (int)S.z[s];
                    const int ze = (e < m) ? (int)S.z[e] : (int)ztotal;
                    const int zb = (int)zown[k] - zs;
                    const int nz = ze - zs;
                    int np, s2, e2;
                    if (((ky >> (sh - 1)) & 1u) == 0) {
                        np = s + zb; s2 = s; e2 = s + nz;
                    } else {
                        np = s + nz + (i - s - zb); s2 = s + nz; e2 = e;
                    }
                    S.key[c ^ 1][np] = ky;
                    S.d[c ^ 1][np] = dreg[k];
                    S.se[c ^ 1][np] = pack_se(s2, e2);
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) p.status[SW_QCOUNT_RET] = S.nret;
}

}  // namespace st
