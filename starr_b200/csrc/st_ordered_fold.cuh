// st_ordered_fold.cuh -- exact, parallel evaluation of an ORDERED floating-point fold
//
//        T <- fl(fl(fl(T + d_0) + d_1) + ... + d_{L-1})
//
// which is what the reference computes at every tree node for the updates below it
// (starr/src/_sumtree_array.pyx:41-49).  A dependent FADD chain costs ~4-12 cycles
// per element on one thread; the root of a 2^20-update batch alone would take milliseconds.
//
// Idea (SURVEY "binade-segmented integer scan"): while T stays inside one binade
// [2^e, 2^(e+1)) its ulp u is constant and T = m*u with an integer m in [2^MB, 2^(MB+1)).
// With x = d/u (exact scaling),  fl(T + d) = u * rne(m + x)  and
//      rne(m + x) = m + a[m & 1]
// where a[0], a[1] are integers that depend only on d and u and differ only when x is an
// exact tie (k + 1/2): then the result is the even neighbour, so it depends on the parity of m.
// Each update is therefore a map on (m) of the form m -> m + a[parity(m)], these maps are closed
// under composition, and composition is associative => a parallel prefix scan gives EVERY
// intermediate m exactly.  The identity fails only when an intermediate value leaves the
// binade (different ulp) or T is not a positive normal number; those positions are detected
// exactly from the scanned prefixes (integer compares), the offending update is applied with a
// real FADD, and the scan restarts behind it in the new binade.  Nothing here is approximate:
// every path reproduces the sequential IEEE result bit for bit.
//
// One CTA of FOLD_THREADS threads processes one segment chunk by chunk (FOLD_CHUNK elements).
#pragma once
#include "st_common.cuh"

namespace st {

constexpr int FOLD_THREADS = 256;
constexpr int FOLD_E = 4;                             // elements per thread per chunk
constexpr int FOLD_CHUNK = FOLD_THREADS * FOLD_E;     // 1024
constexpr int FOLD_SEQ_RUN = 32;  // elements folded sequentially after every restart (bounds the
                                  // cost of inputs that hop across a binade boundary all the time)

template <typename T> struct fp_bits;
template <> struct fp_bits<float> {
    using U = uint32_t;
    using I = int32_t;
    static constexpr int MB = 23;
    static constexpr U EMASK = 0xffu;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ U to_bits(float v) { return __float_as_uint(v); }
    static __device__ __forceinline__ float from_bits(U b) { return __uint_as_float(b); }
};
template <> struct fp_bits<double> {
    using U = uint64_t;
    using I = int64_t;
    static constexpr int MB = 52;
    static constexpr U EMASK = 0x7ffull;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ U to_bits(double v) { return (U)__double_as_longlong(v); }
    static __device__ __forceinline__ double from_bits(U b) { return __longlong_as_double((long long)b); }
};

// m -> m + a[m & 1]
template <typename I> struct pmap {
    I a0, a1;
};
template <typename I> __device__ __forceinline__ pmap<I> pmap_then(const pmap<I> f, const pmap<I> g)
{
    // apply f first, then g
    pmap<I> r;
    r.a0 = f.a0 + ((f.a0 & 1) ? g.a1 : g.a0);
    r.a1 = f.a1 + (((f.a1 + 1) & 1) ? g.a1 : g.a0);
    return r;
}

// One update d seen from a node whose value has exponent field eT (ulp u = 2^(eT - bias - MB)):
//   map  : the parity map of rne(m + d/u)
//   flo  : floor(d/u)      (m + flo inside [2^MB, 2^(MB+1)) <=> the exact sum stays in the binade)
//   big  : |d/u| too large to represent / d is inf or nan  -> always handled by a real FADD
template <typename T> struct upd_map {
    pmap<typename fp_bits<T>::I> map;
    typename fp_bits<T>::I flo;
    bool big;
};

template <typename T> __device__ __forceinline__ upd_map<T> decompose(T d, int eT)
{
    using F = fp_bits<T>;
    using U = typename F::U;
    using I = typename F::I;
    constexpr int MB = F::MB;
    const U b = F::to_bits(d);
    const bool neg = (b >> (F::BITS - 1)) != 0;
    const U ef = (b >> MB) & F::EMASK;
    const U mant = b & ((U(1) << MB) - 1);
    upd_map<T> r;
    r.map.a0 = 0;
    r.map.a1 = 0;
    r.flo = 0;
    r.big = (ef == F::EMASK);
    const U md = ef ? (mant | (U(1) << MB)) : mant;
    if (md == 0 || r.big) return r;  // +-0 changes nothing
    const int ed = ef ? (int)ef : 1;
    const int shift = ed - eT;
    U q, g0, g1;
    bool frac;
    if (shift >= 0) {
        if (shift >= 2) {
            r.big = true;
            return r;
        }
        q = md << shift;
        g0 = g1 = q;
        frac = false;
    } else {
        const int sh = -shift;
        if (sh > MB + 2) {  // |x| < 1/2, not a tie
            q = 0;
            g0 = g1 = 0;
            frac = true;
        } else {
            q = md >> sh;
            const U rem = md & ((U(1) << sh) - 1);
            const U half = U(1) << (sh - 1);
            frac = rem != 0;
            if (rem > half) {
                g0 = g1 = q + 1;
            } else if (rem == half) {  // exact tie: the even neighbour of m + q + 1/2
                g0 = q + (q & 1);
                g1 = q + ((q + 1) & 1);
            } else {
                g0 = g1 = q;
            }
        }
    }
    if (!neg) {
        r.map.a0 = (I)g0;
        r.map.a1 = (I)g1;
        r.flo = (I)q;
    } else {
        r.map.a0 = -(I)g0;
        r.map.a1 = -(I)g1;
        r.flo = -(I)q - (frac ? 1 : 0);
    }
    return r;
}

template <typename T> __device__ __forceinline__ bool scan_ready(T v, int &eT, typename fp_bits<T>::I &m)
{
    using F = fp_bits<T>;
    using U = typename F::U;
    const U b = F::to_bits(v);
    const U ef = (b >> F::MB) & F::EMASK;
    eT = (int)ef;
    m = (typename F::I)((b & ((U(1) << F::MB) - 1)) | (U(1) << F::MB));
    return (b >> (F::BITS - 1)) == 0 && ef >= 1 && ef < F::EMASK;  // positive normal
}

template <typename T> __device__ __forceinline__ T value_of(int eT, typename fp_bits<T>::I m)
{
    using F = fp_bits<T>;
    using U = typename F::U;
    // m in [2^MB, 2^(MB+1)]; m == 2^(MB+1) carries into the exponent, which is exactly 2^(e+1)
    return F::from_bits((U(eT) << F::MB) + ((U)m - (U(1) << F::MB)));
}

struct fold_smem_ctl {
    int first_viol;
    int pos;
    unsigned int slot;
    int pad;
};

// Fold `len` ordered updates src[0..len) into `acc` (exactly as a sequential FADD chain would).
// Must be called by all FOLD_THREADS threads of the CTA; returns the result in every thread.
template <typename T>
__device__ T ordered_fold_cta(T acc, const T *__restrict__ src, int len, T *s_d,
                              typename fp_bits<T>::I (*s_wtot)[2], T *s_acc, fold_smem_ctl *ctl)
{
    using F = fp_bits<T>;
    using I = typename F::I;
    constexpr I LO = I(1) << F::MB;
    constexpr I HI = I(1) << (F::MB + 1);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    if (tid == 0) *s_acc = acc;
    for (int base = 0; base < len; base += FOLD_CHUNK) {
        const int nc = len - base < FOLD_CHUNK ? len - base : FOLD_CHUNK;
        __syncthreads();  // previous chunk fully consumed
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) {
            const int i = tid + k * FOLD_THREADS;
            s_d[i] = (i < nc) ? src[base + i] : T(0);
        }
        if (tid == 0) ctl->pos = 0;
        __syncthreads();
        T dreg[FOLD_E];
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) dreg[k] = s_d[tid * FOLD_E + k];

        for (;;) {
            const int pos = ctl->pos;
            if (pos >= nc) break;
            T cur = *s_acc;
            int eT;
            I m0;
            if (!scan_ready<T>(cur, eT, m0)) {
                // zero / negative / subnormal / inf / nan: plain sequential adds until usable again
                __syncthreads();
                if (tid == 0) {
                    int p = pos;
                    int e2;
                    I m2;
                    do {
                        cur = t_add<T>(cur, s_d[p]);
                        ++p;
                    } while (p < nc && !scan_ready<T>(cur, e2, m2));
                    *s_acc = cur;
                    ctl->pos = p;
                }
                __syncthreads();
                continue;
            }
            // ---- per-thread maps of the still-pending elements ---------------------------------
            upd_map<T> um[FOLD_E];
            pmap<I> local;
            local.a0 = 0;
            local.a1 = 0;
#pragma unroll
            for (int k = 0; k < FOLD_E; ++k) {
                const int i = tid * FOLD_E + k;
                if (i >= pos && i < nc) {
                    um[k] = decompose<T>(dreg[k], eT);
                } else {
                    um[k].map.a0 = 0;
                    um[k].map.a1 = 0;
                    um[k].flo = 0;
                    um[k].big = false;
                }
                local = pmap_then<I>(local, um[k].map);
            }
            // ---- exclusive scan of the maps over the CTA -----------------------------------------
            pmap<I> incl = local;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                pmap<I> o;
                if constexpr (sizeof(I) == 8) {
                    o.a0 = (I)__shfl_up_sync(0xffffffffu, (long long)incl.a0, off);
                    o.a1 = (I)__shfl_up_sync(0xffffffffu, (long long)incl.a1, off);
                } else {
                    o.a0 = __shfl_up_sync(0xffffffffu, incl.a0, off);
                    o.a1 = __shfl_up_sync(0xffffffffu, incl.a1, off);
                }
                if (lane >= off) incl = pmap_then<I>(o, incl);
            }
            if (lane == 31) {
                s_wtot[warp][0] = incl.a0;
                s_wtot[warp][1] = incl.a1;
            }
            pmap<I> excl;  // maps of all earlier lanes of this warp
            if constexpr (sizeof(I) == 8) {
                excl.a0 = (I)__shfl_up_sync(0xffffffffu, (long long)incl.a0, 1);
                excl.a1 = (I)__shfl_up_sync(0xffffffffu, (long long)incl.a1, 1);
            } else {
                excl.a0 = __shfl_up_sync(0xffffffffu, incl.a0, 1);
                excl.a1 = __shfl_up_sync(0xffffffffu, incl.a1, 1);
            }
            if (lane == 0) {
                excl.a0 = 0;
                excl.a1 = 0;
            }
            if (tid == 0) ctl->first_viol = 0x7fffffff;
            __syncthreads();
            pmap<I> pre;
            pre.a0 = 0;
            pre.a1 = 0;
            for (int w = 0; w < warp; ++w) {
                pmap<I> t;
                t.a0 = s_wtot[w][0];
                t.a1 = s_wtot[w][1];
                pre = pmap_then<I>(pre, t);
            }
            pre = pmap_then<I>(pre, excl);
            // ---- walk my elements with the real m, find the first one the identity does not cover --
            I m = m0 + ((m0 & 1) ? pre.a1 : pre.a0);
            int viol = 0x7fffffff;
            I m_at_viol = 0;
#pragma unroll
            for (int k = 0; k < FOLD_E; ++k) {
                const int i = tid * FOLD_E + k;
                if (i >= pos && i < nc && viol == 0x7fffffff) {
                    const I v = m + um[k].flo;
                    if (um[k].big || m < LO || m >= HI || v < LO || v >= HI) {
                        viol = i;
                        m_at_viol = m;
                    } else {
                        m += (m & 1) ? um[k].map.a1 : um[k].map.a0;
                    }
                }
            }
            if (viol != 0x7fffffff) atomicMin(&ctl->first_viol, viol);
            __syncthreads();
            const int fv = ctl->first_viol;
            if (fv == 0x7fffffff) {
                // whole rest of the chunk stayed inside the binade: the owner of the last element
                // holds the final m
                if (tid == (nc - 1) / FOLD_E) {
                    *s_acc = value_of<T>(eT, m);
                    ctl->pos = nc;
                }
            } else if (viol == fv) {
                // every element before fv is exact; apply fv with a real add, then a short
                // sequential run so that boundary-hopping inputs stay cheap
                T v = t_add<T>(value_of<T>(eT, m_at_viol), s_d[fv]);
                int p = fv + 1;
                const int stop = (p + FOLD_SEQ_RUN < nc) ? p + FOLD_SEQ_RUN : nc;
                for (; p < stop; ++p) v = t_add<T>(v, s_d[p]);
                *s_acc = v;
                ctl->pos = p;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    return *s_acc;
}

}  // namespace st
