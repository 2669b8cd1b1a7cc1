// st_tree_kernels.cu -- build, prefix-sum search / sampling, range sums, axis fold.
//
// Hand-written sm_100a kernels that replace (reference file:line)
//   build_sumtree_from_array   starr/src/_sumtree_array.pyx:51-64
//   get_prefix_sum_idx         starr/src/_sumtree_array.pyx:66-122
//   sum_over / strided_sum     starr/src/_sumtree_array.pyx:131-194
//   the ndarray.sum fallback   starr/sumtree_array.py:514
// on the single device buffer heap[2n] described in include/starr_b200.h.
// All of them are HBM/L2 latency- or bandwidth-bound gathers and streams; there is no
// contraction anywhere, so no tensor cores.
#include <cstring>

#include "st_common.cuh"
#include "st_sidx.cuh"

namespace st {

// =========================================================================================
// first kernel of a batch: the previous batch's validation flags move to the sticky word
// =========================================================================================
__global__ void batch_begin_kernel(unsigned int *status)
{
    if (threadIdx.x == 0) {
        const unsigned int f = status[0];
        if (f) {
            status[ST_SW_STICKY] |= f;
            status[0] = 0;
        }
    }
}

cudaError_t launch_batch_begin(st_tree *t, cudaStream_t s)
{
    ++g_kernel_launches;
    batch_begin_kernel<<<1, 32, 0, s>>>(t->d_status);
    return cudaGetLastError();
}

// =========================================================================================
// leaf validation (pyx:58-59: min(array) < 0 -> ValueError before any write)
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256) check_leaves_kernel(const T *__restrict__ leaves, int64_t n,
                                                           unsigned int *status)
{
    if constexpr (!is_signed_t<T>::value) {
        return;
    } else {
        bool bad = false;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        // 128-bit streaming loads over the 16-byte aligned middle, scalars for the ragged ends
        constexpr int VEC = 16 / (int)sizeof(T);
        const uintptr_t addr = reinterpret_cast<uintptr_t>(leaves);
        int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
        if (head > n) head = n;
        const int64_t nvec = (n - head) / VEC;
        const int4 *v4 = reinterpret_cast<const int4 *>(leaves + head);
        int64_t i = tid;
        for (; i + 3 * stride < nvec; i += 4 * stride) {
            int4 q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = __ldg(v4 + i + k * stride);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T *e = reinterpret_cast<const T *>(&q[k]);
#pragma unroll
                for (int j = 0; j < VEC; ++j) bad |= (e[j] < T(0));
            }
        }
        for (; i < nvec; i += stride) {
            const int4 q = __ldg(v4 + i);
            const T *e = reinterpret_cast<const T *>(&q);
#pragma unroll
            for (int j = 0; j < VEC; ++j) bad |= (e[j] < T(0));
        }
        for (int64_t j = tid; j < head; j += stride) bad |= (leaves[j] < T(0));
        for (int64_t j = head + nvec * VEC + tid; j < n; j += stride) bad |= (leaves[j] < T(0));
        if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, STF_NEG_ARRAY);
    }
}

cudaError_t launch_check_leaves(st_tree *t, cudaStream_t s)
{
    ST_DISPATCH(t->dtype, {
        const T *leaves = (const T *)t->heap + t->n;
        unsigned g = grid_for(t->n, 256 * 16, (int64_t)t->sm_count * 8);
        ++g_kernel_launches;
        check_leaves_kernel<T><<<g, 256, 0, s>>>(leaves, t->n, t->d_status);
    });
    return cudaGetLastError();
}

// =========================================================================================
// build: node = left + right, bottom-up, K levels per pass inside shared memory
// =========================================================================================
// One CTA owns the subtree rooted at node r of level `root_level` and computes its levels
// root_level+k-1 .. root_level from the already-final level root_level+k.  A node h >= n is
// a leaf (loaded, never computed); h >= 2n does not exist.  Pairwise left+right is exactly
// the reference's loop order-independent result (every node is one add of its two children).
template <typename T> struct alignas(2 * sizeof(T)) pair_t { T l, r; };
template <typename T> struct alignas(4 * sizeof(T)) quad_t { T a, b, c, d; };

constexpr int BUILD_K = 11;        // levels per pass
constexpr int BUILD_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(BUILD_THREADS)
build_pass_kernel(T *__restrict__ H, uint32_t n, int root_level, int k, const unsigned int *status)
{
    if (*status) return;  // failed validation: leave the tree untouched (pyx:58-59)
    __shared__ T buf[2][1 << (BUILD_K - 1)];
    const uint32_t r = (1u << root_level) + blockIdx.x;
    const uint64_t two_n = 2ull * n;

    // level root_level + k - 1, children read from global as aligned pairs
    int lev = k - 1;
    {
        const uint32_t cnt = 1u << lev;
        const uint64_t base = (uint64_t)r << lev;
        if (lev >= 1 && base + cnt <= n) {
            // interior tile: two parents per thread from one aligned 4-child vector (16 B for 4-byte T)
            for (uint32_t i = 2 * threadIdx.x; i < cnt; i += 2 * BUILD_THREADS) {
                const uint64_t h = base + i;
                quad_t<T> c = *reinterpret_cast<const quad_t<T> *>(H + 2 * h);
                pair_t<T> o;
                o.l = t_add<T>(c.a, c.b);
                o.r = t_add<T>(c.c, c.d);
                *reinterpret_cast<pair_t<T> *>(H + h) = o;
                buf[0][i] = o.l;
                buf[0][i + 1] = o.r;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < cnt; i += BUILD_THREADS) {
                const uint64_t h = base + i;
                T v = T(0);
                if (h < n) {
                    pair_t<T> c = *reinterpret_cast<const pair_t<T> *>(H + 2 * h);
                    v = t_add<T>(c.l, c.r);
                    H[h] = v;
                } else if (h < two_n) {
                    v = H[h];
                }
                buf[0][i] = v;
            }
        }
    }
    int cur = 0;
    while (lev > 0) {
        __syncthreads();
        --lev;
        const uint32_t cnt = 1u << lev;
        const uint64_t base = (uint64_t)r << lev;
        for (uint32_t i = threadIdx.x; i < cnt; i += BUILD_THREADS) {
            const uint64_t h = base + i;
            T v = T(0);
            if (h < n) {
                v = t_add<T>(buf[cur][2 * i], buf[cur][2 * i + 1]);
                H[h] = v;
            } else if (h < two_n) {
                v = H[h];
            }
            buf[cur ^ 1][i] = v;
        }
        cur ^= 1;
    }
}

// ---- search index (st_sidx.cuh) ------------------------------------------------------------------------
size_t sidx_bytes_for(int64_t n, int depth, int esize, int kind, int *nb_out)
{
    // Opt-in (ST_SEARCH_INDEX=1).  Measured at 2^27 float32 leaves, 2^20 updates + 2^20 samples per step
    // (profiles/r2_experiments.md): the search drops from 150 to 110 us (14 -> 6 dependent loads per query,
    // 9.5e9 samples/s), but every update of a left child below level 13 becomes two atomics and the update grows
    // from 0.61 to 0.92 ms -- the L2 atomic units are what bounds the update.  Worth it only for sample-heavy loads.
    static const bool on = getenv("ST_SEARCH_INDEX") && atoi(getenv("ST_SEARCH_INDEX")) != 0;
    int nb = 0;
    if (on && esize == 4 && kind == 0 && 2 * n >= 16384) nb = (depth - 2 - SIDX_L0) / 3;
    if (nb < 0) nb = 0;
    if (nb > SIDX_MAX_BLOCK_LEVELS) nb = SIDX_MAX_BLOCK_LEVELS;
    if (nb_out) *nb_out = nb;
    size_t blocks = 0;
    for (int i = 0; i < nb; ++i) blocks += (size_t)1 << (SIDX_L0 + 3 * i);
    return blocks * 32;
}

sidx_ref sidx_of(const st_tree *t)
{
    sidx_ref sx;
    if (!t->sidx) return sx;
    sx.base = t->sidx;
    sx.nb = t->sidx_nb;
    uint32_t off = 0;
    for (int i = 0; i < sx.nb; ++i) {
        sx.off[i] = off;
        off += 8u << (SIDX_L0 + 3 * i);
    }
    return sx;
}

// one thread per record: block root r of level L0 + 3i -> [0, H[2r], H[4r], H[4r+2], H[8r], H[8r+2], H[8r+4], H[8r+6]]
__global__ void __launch_bounds__(256) sidx_build_kernel(const uint32_t *__restrict__ H, sidx_ref sx, const unsigned int *status)
{
    if (*status) return;   // failed validation: nothing was built (pyx:58-59)
    uint32_t total = 0;
    for (int i = 0; i < sx.nb; ++i) total += 1u << (SIDX_L0 + 3 * i);
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        int i = 0;
        uint32_t b = g;
        while (b >= (1u << (SIDX_L0 + 3 * i))) { b -= 1u << (SIDX_L0 + 3 * i); ++i; }
        const uint32_t r = (1u << (SIDX_L0 + 3 * i)) + b;
        uint4 lo, hi;
        lo.x = 0u;        lo.y = H[2 * r];     lo.z = H[4 * r];     lo.w = H[4 * r + 2];
        hi.x = H[8 * r];  hi.y = H[8 * r + 2]; hi.z = H[8 * r + 4]; hi.w = H[8 * r + 6];
        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(sx.base) + sx.off[i] + b * 8u);
        dst[0] = lo;
        dst[1] = hi;
    }
}

cudaError_t launch_sidx_build(st_tree *t, cudaStream_t s)
{
    if (!t->sidx) return cudaSuccess;
    ++g_kernel_launches;
    sidx_build_kernel<<<t->sm_count * 8, 256, 0, s>>>((const uint32_t *)t->heap, sidx_of(t), t->d_status);
    return cudaGetLastError();
}

cudaError_t launch_build(st_tree *t, cudaStream_t s)
{
    const uint32_t n = (uint32_t)t->n;
    int top = t->depth;  // deepest level, already final (leaves)
    ST_DISPATCH(t->dtype, {
        T *H = (T *)t->heap;
        while (top > 0) {
            int root_level = top - BUILD_K;
            if (root_level < 0) root_level = 0;
            int k = top - root_level;
            uint64_t first = 1ull << root_level;
            uint64_t last = first * 2 < n ? first * 2 : n;  // roots r >= n have nothing to compute
            if (last > first) {
                ++g_kernel_launches;
                build_pass_kernel<T><<<(unsigned)(last - first), BUILD_THREADS, 0, s>>>(
                    H, n, root_level, k, t->d_status);
            }
            top = root_level;
        }
    });
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_sidx_build(t, s);
}

// =========================================================================================
// prefix-sum search (pyx:84-120) and sampling (sumtree_array.py:417-421)
// =========================================================================================
// Each thread walks Q independent queries in lock step so that Q dependent-load chains are
// in flight per thread.  The top of the tree (heap[0 .. stage_nodes)) is staged once per
// CTA into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier); deeper levels are
// read-only gathers through L1/L2 (ld.global.nc).
// large batches: 512 threads x 4 queries per thread; small batches: 128 threads x 1 query so that
// a 4096-query PER step still spreads over 32 SMs
constexpr int SEARCH_THREADS_BIG = 512, SEARCH_Q_BIG = 4;
constexpr int SEARCH_THREADS_SMALL = 128, SEARCH_Q_SMALL = 1;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Stage `bytes` (multiple of 16) from global to shared with the TMA engine; all threads
// return once the data has landed.
__device__ __forceinline__ void tma_stage(void *smem_dst, const void *gsrc, uint32_t bytes,
                                          uint64_t *bar)
{
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes)
                     : "memory");
        const uint32_t CH = 32768;
        for (uint32_t off = 0; off < bytes; off += CH) {
            uint32_t sz = bytes - off < CH ? bytes - off : CH;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(smem_u32((const char *)smem_dst + off)),
                "l"((const char *)gsrc + off), "r"(sz), "r"(bar_a)
                : "memory");
        }
    }
    // every thread waits on phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar_a)
            : "memory");
    }
}

template <typename T> __device__ __forceinline__ T scale_uniform(T root, double u)
{
    // (root * u).astype(dtype): float64 multiply, then one conversion (sumtree_array.py:420)
    if constexpr (sizeof(T) == 4 && is_float_t<T>::value) {
        return __double2float_rn(__dmul_rn((double)root, u));
    } else if constexpr (is_float_t<T>::value) {
        return __dmul_rn(root, u);
    } else {
        return (T)__dmul_rn((double)root, u);  // C truncation toward zero
    }
}

// NumPy's `x[idx] / x.sum()`: float32 / float32 -> one float32 division, float64 likewise, integer
// dtypes -> true division in float64 (README.md:7 PER importance weights are built from this ratio)
template <typename T> struct prob_type { using type = double; };
template <> struct prob_type<float> { using type = float; };
template <typename T> __device__ __forceinline__ typename prob_type<T>::type prob_div(T p, T total)
{
    if constexpr (sizeof(T) == 4 && is_float_t<T>::value) return __fdiv_rn(p, total);
    else return __ddiv_rn((double)p, (double)total);
}

template <typename T, bool STAGE, bool UNIFORM, int SEARCH_THREADS, int SEARCH_Q>
__global__ void __launch_bounds__(SEARCH_THREADS)
search_kernel(const T *__restrict__ H, uint32_t n, int depth, const void *__restrict__ in,
              int64_t m, int64_t *__restrict__ out, T *__restrict__ prio, T *__restrict__ resid,
              uint32_t stage_nodes, const unsigned int *__restrict__ m_dev,
              typename prob_type<T>::type *__restrict__ prob, const search_push push, const sidx_ref sx)
{
    if (m_dev) {   // the host did not know the count when it launched (sharded sample: st_shard_plan_device)
        const int64_t md = (int64_t)*m_dev;
        if (md < m) m = md;
    }
    // segmented input (search_push::seg_cnt): query i is element i - s_seg[r] of the segment r with s_seg[r] <= i < s_seg[r + 1]
    __shared__ unsigned int s_seg[65];
    if (push.seg_cnt) {
        if (threadIdx.x == 0) {
            unsigned int run = 0;
            for (int r = 0; r < push.n; ++r) { s_seg[r] = run; run += push.seg_cnt[r]; }
            s_seg[push.n] = run;
        }
        __syncthreads();
        m = (int64_t)s_seg[push.n];
    }
    auto seg_of = [&](int64_t i, int &r) -> int64_t {   // element index inside the strided segment layout
        r = 0;
        while (i >= (int64_t)s_seg[r + 1]) ++r;
        return (int64_t)r * push.seg_stride + (i - (int64_t)s_seg[r]);
    };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *top = reinterpret_cast<T *>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    if constexpr (STAGE) {
        tma_stage(top, H, stage_nodes * (uint32_t)sizeof(T), &bar);
    } else {
        stage_nodes = 0;
    }
    T root = T(0);
    if constexpr (UNIFORM) root = STAGE ? top[1] : __ldg(H + 1);

    const int64_t per_block = (int64_t)SEARCH_THREADS * SEARCH_Q;
    for (int64_t base = (int64_t)blockIdx.x * per_block; base < m; base += (int64_t)gridDim.x * per_block) {
        uint32_t h[SEARCH_Q];
        T p[SEARCH_Q];
#pragma unroll
        for (int q = 0; q < SEARCH_Q; ++q) {
            const int64_t i = base + threadIdx.x + (int64_t)q * SEARCH_THREADS;
            h[q] = 1;
            p[q] = T(0);
            if (i < m) {
                if constexpr (UNIFORM) {
                    p[q] = scale_uniform<T>(root, __ldg((const double *)in + i));
                } else if (push.seg_cnt) {
                    int r;
                    p[q] = __ldcg((const T *)in + seg_of(i, r));   // written by a peer a moment ago: not through the read-only path
                } else {
                    p[q] = __ldg((const T *)in + i);
                }
            } else {
                h[q] = n;  // inactive: already "at a leaf"
            }
        }
        int lev = 0;
        if constexpr (STAGE && sizeof(T) == 4) {
            if (sx.base) {
                // (a) the staged levels: every left child is in shared memory (L0 = 13 steps)
                for (; lev < SIDX_L0; ++lev) {
#pragma unroll
                    for (int q = 0; q < SEARCH_Q; ++q) {
                        if (h[q] < n) {
                            const uint32_t c = 2u * h[q];
                            const T lv = top[c];
                            if (p[q] >= lv) { p[q] = t_sub<T>(p[q], lv); h[q] = c + 1; } else { h[q] = c; }
                        }
                    }
                }
                // (b) three levels per 32-byte index record (st_sidx.cuh); all nodes here are internal
                for (int i = 0; i < sx.nb; ++i, lev += 3) {
                    uint4 lo[SEARCH_Q], hi[SEARCH_Q];
                    const uint4 *rec = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(sx.base) + sx.off[i]);
#pragma unroll
                    for (int q = 0; q < SEARCH_Q; ++q) {
                        lo[q] = make_uint4(0u, 0u, 0u, 0u);
                        hi[q] = lo[q];
                        if (h[q] < n) {
                            const uint32_t b = h[q] - (1u << lev);
                            lo[q] = __ldcg(rec + 2 * b);
                            hi[q] = __ldcg(rec + 2 * b + 1);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < SEARCH_Q; ++q) {
                        if (h[q] < n) {
                            auto as_t = [](uint32_t w) { T v; memcpy(&v, &w, 4); return v; };
                            T lv = as_t(lo[q].y);
                            uint32_t path = 0;
                            if (p[q] >= lv) { p[q] = t_sub<T>(p[q], lv); path = 1; }
                            lv = as_t(path ? lo[q].w : lo[q].z);
                            path <<= 1;
                            if (p[q] >= lv) { p[q] = t_sub<T>(p[q], lv); path |= 1; }
                            const uint32_t w3 = (path & 2u) ? ((path & 1u) ? hi[q].w : hi[q].z) : ((path & 1u) ? hi[q].y : hi[q].x);
                            lv = as_t(w3);
                            path <<= 1;
                            if (p[q] >= lv) { p[q] = t_sub<T>(p[q], lv); path |= 1; }
                            h[q] = (h[q] << 3) + path;
                        }
                    }
                }
            }
        }
        for (; lev < depth; ++lev) {
            T lv[SEARCH_Q];
#pragma unroll
            for (int q = 0; q < SEARCH_Q; ++q) {
                const uint32_t c = 2u * h[q];
                lv[q] = T(0);
                if (h[q] < n) lv[q] = (c < stage_nodes) ? top[c] : __ldcg(H + c);
            }
#pragma unroll
            for (int q = 0; q < SEARCH_Q; ++q) {
                if (h[q] < n) {
                    const uint32_t c = 2u * h[q];
                    if (p[q] >= lv[q]) {  // pyx:107: ties go right
                        p[q] = t_sub<T>(p[q], lv[q]);
                        h[q] = c + 1;
                    } else {
                        h[q] = c;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SEARCH_Q; ++q) {
            const int64_t i = base + threadIdx.x + (int64_t)q * SEARCH_THREADS;
            if (i < m) {
                if (push.seg_cnt) {
                    int r;
                    const int64_t e = seg_of(i, r);
                    reinterpret_cast<int2 *>(push.base[r])[e - (int64_t)r * push.seg_stride] =
                        make_int2(__ldcg(push.pos + e), (int)(h[q] - n));
                } else if (push.leaf32) {
                    push.leaf32[i] = (int32_t)(h[q] - n);
                } else if (push.pairs) {
                    // sharded sample, coalesced exchange: (batch position, local leaf) in this rank's own order
                    reinterpret_cast<int2 *>(push.pairs)[i] = make_int2(push.pos[i], (int)(h[q] - n));
                } else if (push.n) {
                    // sharded sample: the global leaf index goes straight into every rank's result buffer
                    const int64_t g = (int64_t)h[q] - (int64_t)n + push.index_offset;
                    const int32_t at = push.pos[i];
                    for (int r = 0; r < push.n; ++r) reinterpret_cast<int64_t *>(push.base[r])[at] = g;
                } else {
                    out[i] = (int64_t)h[q] - (int64_t)n;
                }
                if (prio || prob) {
                    const T pv = __ldcg(H + h[q]);
                    if (prio) prio[i] = pv;
                    // importance-sampling probability p_i / total: what `x[idx] / x.sum()` gives on the host
                    if constexpr (UNIFORM) { if (prob) prob[i] = prob_div<T>(pv, root); }
                }
                if (resid) resid[i] = p[q];   // prefix left over inside the found leaf's subtree
            }
        }
    }
}

template <typename T, bool UNIFORM, int THREADS, int Q>
static cudaError_t launch_search_cfg(const st_tree *t, const void *in, int64_t m, int64_t *out, void *prio,
                                     void *resid, cudaStream_t s, const unsigned int *m_dev, void *prob,
                                     const search_push &push)
{
    const uint32_t n = (uint32_t)t->n;
    const T *H = (const T *)t->heap;
    const int64_t per_block = (int64_t)THREADS * Q;
    // stage the top 64 KiB of the heap (TMA bulk copy) when every CTA has enough queries to amortise it
    uint32_t stage_nodes = 65536u / (uint32_t)sizeof(T);
    const uint64_t two_n = 2ull * n;
    bool stage = m >= 2048;
    if (two_n < stage_nodes) {
        // whole heap fits: round down to a multiple of 16 bytes (the tail is read from global)
        stage_nodes = (uint32_t)(two_n * sizeof(T) / 16 * 16 / sizeof(T));
        if (stage_nodes < 4) stage = false;
    }
    if (stage) {
        const size_t smem = (size_t)stage_nodes * sizeof(T);
        auto kern = search_kernel<T, true, UNIFORM, THREADS, Q>;
        static per_device_once attr_once;   // function attributes are per device
        cudaError_t ea = attr_once.run(t->device, [&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        });
        if (ea != cudaSuccess) return ea;
        unsigned g = grid_for(m, (int)per_block, (int64_t)t->sm_count * 3);
        ++g_kernel_launches;
        kern<<<g, THREADS, smem, s>>>(H, n, t->depth, in, m, out, (T *)prio, (T *)resid, stage_nodes, m_dev,
                                      (typename prob_type<T>::type *)prob, push, sidx_of(t));
    } else {
        unsigned g = grid_for(m, (int)per_block, (int64_t)t->sm_count * 4);
        ++g_kernel_launches;
        search_kernel<T, false, UNIFORM, THREADS, Q><<<g, THREADS, 0, s>>>(H, n, t->depth, in, m, out, (T *)prio,
                                                                          (T *)resid, 0, m_dev,
                                                                          (typename prob_type<T>::type *)prob, push, sidx_ref{});
    }
    return cudaGetLastError();
}

template <typename T, bool UNIFORM>
static cudaError_t launch_search_t(const st_tree *t, const void *in, int64_t m, int64_t *out, void *prio,
                                   void *resid, cudaStream_t s, const unsigned int *m_dev = nullptr, void *prob = nullptr,
                                   const search_push *push = nullptr)
{
    if (m <= 0) return cudaSuccess;
    const search_push none{};
    const search_push &ps = push ? *push : none;
    if (m >= (int64_t)t->sm_count * SEARCH_THREADS_BIG * SEARCH_Q_BIG / 4)
        return launch_search_cfg<T, UNIFORM, SEARCH_THREADS_BIG, SEARCH_Q_BIG>(t, in, m, out, prio, resid, s, m_dev, prob, ps);
    return launch_search_cfg<T, UNIFORM, SEARCH_THREADS_SMALL, SEARCH_Q_SMALL>(t, in, m, out, prio, resid, s, m_dev, prob, ps);
}

cudaError_t launch_search(const st_tree *t, const void *vals, int64_t m, int64_t *out, cudaStream_t s,
                          const unsigned int *m_dev, const search_push *push)
{
    ST_DISPATCH(t->dtype, return (launch_search_t<T, false>(t, vals, m, out, nullptr, nullptr, s, m_dev, nullptr, push)));
    return cudaSuccess;
}

cudaError_t launch_sample(const st_tree *t, const double *u, int64_t m, int64_t *out, void *prio,
                          void *resid, cudaStream_t s, void *prob)
{
    ST_DISPATCH(t->dtype, return (launch_search_t<T, true>(t, u, m, out, prio, resid, s, nullptr, prob)));
    return cudaSuccess;
}

// =========================================================================================
// publish the status word and the root into (mapped pinned) memory the host can read after a sync
// =========================================================================================
template <typename T>
__global__ void publish_kernel(const T *__restrict__ H, const unsigned int *__restrict__ status,
                               unsigned int *dst_status, T *dst_root)
{
    if (threadIdx.x == 0) {
        if (dst_status) *dst_status = status[0] | status[ST_SW_STICKY];
        if (dst_root) *dst_root = H[1];
    }
}

cudaError_t launch_publish(const st_tree *t, void *dst_status, void *dst_root, cudaStream_t s)
{
    ST_DISPATCH(t->dtype, {
        ++g_kernel_launches;
        publish_kernel<T><<<1, 32, 0, s>>>((const T *)t->heap, t->d_status, (unsigned int *)dst_status, (T *)dst_root);
    });
    return cudaGetLastError();
}

// =========================================================================================
// leaf gather
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const T *__restrict__ leaves, int64_t n,
                                                     const int64_t *__restrict__ idx, int64_t m,
                                                     T *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t l = idx[i];
        out[i] = (l >= 0 && l < n) ? __ldcg(leaves + l) : T(0);
    }
}

cudaError_t launch_gather(const st_tree *t, const int64_t *idx, int64_t m, void *out, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    ST_DISPATCH(t->dtype, {
        ++g_kernel_launches;
        gather_kernel<T><<<grid_for(m, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>(
            (const T *)t->heap + t->n, t->n, idx, m, (T *)out);
    });
    return cudaGetLastError();
}

// =========================================================================================
// in-place leaf arithmetic (reference sumtree_array.py:150-173 in-place ufuncs, :202-205 fill):
// leaves = leaves (op) operand, operand a scalar or an array of n elements of the tree dtype.
// One IEEE operation per element in the element type (integers wrap like NumPy's), so the leaves
// are bit-identical to what the host ufunc writes; the caller rebuilds the tree afterwards.
// =========================================================================================
template <typename T, int OP> __device__ __forceinline__ T inplace_apply(T a, T b)
{
    if constexpr (OP == ST_OP_ADD) return t_add<T>(a, b);
    else if constexpr (OP == ST_OP_SUB) return t_sub<T>(a, b);
    else if constexpr (OP == ST_OP_MUL) {
        if constexpr (sizeof(T) == 4 && is_float_t<T>::value) return __fmul_rn(a, b);
        else if constexpr (is_float_t<T>::value) return __dmul_rn(a, b);
        else return (T)(a * b);
    } else if constexpr (OP == ST_OP_DIV) {
        if constexpr (sizeof(T) == 4 && is_float_t<T>::value) return __fdiv_rn(a, b);
        else if constexpr (is_float_t<T>::value) return __ddiv_rn(a, b);
        else return a;   // integer true_divide does not exist in place (NumPy raises); rejected by the launcher
    } else return b;     // ST_OP_FILL
}

template <typename T, int OP, bool ARRAY>
__global__ void __launch_bounds__(256) inplace_kernel(T *__restrict__ x, int64_t n, T scalar, const T *__restrict__ y)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // the leaves start at heap + n: 16-byte aligned only when n * sizeof(T) is; peel to alignment
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    bool vec_ok = true;
    if constexpr (ARRAY) vec_ok = ((reinterpret_cast<uintptr_t>(y + head) & 15) == 0);
    const int64_t nvec = vec_ok ? (n - head) / VEC : 0;
    int4 *x4 = reinterpret_cast<int4 *>(x + head);
    const int4 *y4 = reinterpret_cast<const int4 *>(y + head);
    for (int64_t i = tid; i < nvec; i += stride) {
        int4 q;
        if constexpr (OP != ST_OP_FILL) q = x4[i];
        int4 r;
        if constexpr (ARRAY) r = __ldg(y4 + i);
        T *e = reinterpret_cast<T *>(&q);
        const T *f = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int j = 0; j < VEC; ++j) e[j] = inplace_apply<T, OP>(e[j], ARRAY ? f[j] : scalar);
        x4[i] = q;
    }
    for (int64_t j = tid; j < head; j += stride) x[j] = inplace_apply<T, OP>(x[j], ARRAY ? y[j] : scalar);
    for (int64_t j = head + nvec * VEC + tid; j < n; j += stride) x[j] = inplace_apply<T, OP>(x[j], ARRAY ? y[j] : scalar);
}

template <typename T, int OP>
static cudaError_t launch_inplace_op(st_tree *t, const void *scalar_host, const void *operand, cudaStream_t s)
{
    T *leaves = (T *)t->heap + t->n;
    const unsigned g = grid_for(t->n, 256 * (16 / (int)sizeof(T)) * 4, (int64_t)t->sm_count * 16);
    ++g_kernel_launches;
    if (operand) {
        inplace_kernel<T, OP, true><<<g, 256, 0, s>>>(leaves, t->n, T(0), (const T *)operand);
    } else {
        T sc;
        memcpy(&sc, scalar_host, sizeof(T));
        inplace_kernel<T, OP, false><<<g, 256, 0, s>>>(leaves, t->n, sc, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t launch_inplace(st_tree *t, int op, const void *scalar_host, const void *operand, cudaStream_t s)
{
    const bool is_f = (t->dtype == ST_F32 || t->dtype == ST_F64);
    if (op == ST_OP_DIV && !is_f) return cudaErrorInvalidValue;
    if (op == ST_OP_FILL && operand) return cudaErrorInvalidValue;
    ST_DISPATCH(t->dtype, {
        switch (op) {
            case ST_OP_ADD: return launch_inplace_op<T, ST_OP_ADD>(t, scalar_host, operand, s);
            case ST_OP_SUB: return launch_inplace_op<T, ST_OP_SUB>(t, scalar_host, operand, s);
            case ST_OP_MUL: return launch_inplace_op<T, ST_OP_MUL>(t, scalar_host, operand, s);
            case ST_OP_DIV: return launch_inplace_op<T, ST_OP_DIV>(t, scalar_host, operand, s);
            case ST_OP_FILL: return launch_inplace_op<T, ST_OP_FILL>(t, scalar_host, operand, s);
            default: return cudaErrorInvalidValue;
        }
    });
    return cudaSuccess;
}

// =========================================================================================
// range sums (pyx:131-170) -- the reference's exact subtraction sequence, one thread per output
// =========================================================================================
template <typename T> __device__ __forceinline__ T range_sum(const T *__restrict__ H, int64_t n, int64_t a, int64_t b)
{
    b -= 1;
    if (a == b) return __ldg(H + n + a);
    if (a > b) return T(0);
    a += n;
    b += n;
    T sub = T(0);
    const bool wrapped = a < (a ^ b);  // ends at different depths (pyx:144)
    if (wrapped) {
        if ((b & 1) == 0) sub = t_add<T>(sub, __ldg(H + b + 1));
        b >>= 1;
    }
    while ((a >> 1) != (b >> 1)) {
        if (a & 1) sub = t_add<T>(sub, __ldg(H + a - 1));
        if ((b & 1) == 0) sub = t_add<T>(sub, __ldg(H + b + 1));
        a >>= 1;
        b >>= 1;
    }
    const T cover = wrapped ? __ldg(H + 1) : __ldg(H + (a >> 1));
    return t_sub<T>(cover, sub);
}

template <typename T>
__global__ void __launch_bounds__(256) strided_sum_kernel(const T *__restrict__ H, int64_t n, int64_t stride,
                                                          T *__restrict__ out, int64_t nout)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = stride * i;
        int64_t hi = lo + stride;
        if (hi > n) hi = n;
        out[i] = range_sum<T>(H, n, lo, hi);
    }
}

template <typename T>
__global__ void sum_over_kernel(const T *__restrict__ H, int64_t n, int64_t a, int64_t b, T *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = range_sum<T>(H, n, a, b);
}

cudaError_t launch_sum_over(const st_tree *t, int64_t a, int64_t b, void *out_dev, cudaStream_t s)
{
    ++g_kernel_launches;
    ST_DISPATCH(t->dtype, { sum_over_kernel<T><<<1, 32, 0, s>>>((const T *)t->heap, t->n, a, b, (T *)out_dev); });
    return cudaGetLastError();
}

cudaError_t launch_strided_sum(const st_tree *t, int64_t stride, void *out, int64_t nout, cudaStream_t s)
{
    if (nout <= 0) return cudaSuccess;
    ST_DISPATCH(t->dtype, {
        ++g_kernel_launches;
        strided_sum_kernel<T><<<grid_for(nout, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>(
            (const T *)t->heap, t->n, stride, (T *)out, nout);
    });
    return cudaGetLastError();
}

}  // namespace st
