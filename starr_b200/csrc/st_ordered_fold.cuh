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
// Two drivers share the arithmetic: a warp folds medium segments (256-element chunks, registers
// and shuffles only) and a CTA folds long ones (1024-element chunks, two barriers per chunk).
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
//   flq  : floor(4 d/u).  The identity holds while the exact sum s = m + d/u stays inside
//              [2^MB - 1/4, 2^(MB+1))   <=>   4 (m - 2^MB) + flq >= -1   and   m + (flq >> 2) < 2^(MB+1):
//          inside the binade the rounding grid is u; in the quarter below it (grid u/2) both grids round
//          to 2^MB itself -- s = 2^MB - 1/4 is a tie between 2^MB - 1/2 (odd) and 2^MB (even).  Without
//          that quarter a node sitting EXACTLY on a power of two would fail on every tiny negative update
//          (uniform priorities on a power-of-two tree put the upper nodes there, and the boundary is sticky:
//          from 2^MB small positive updates are absorbed, small negative ones too).
//   big  : |d/u| too large to represent / d is inf or nan  -> always handled by a real FADD
template <typename T> struct upd_map {
    pmap<typename fp_bits<T>::I> map;
    typename fp_bits<T>::I flq;
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
    r.flq = 0;
    r.big = (ef == F::EMASK);
    const U md = ef ? (mant | (U(1) << MB)) : mant;
    if (md == 0 || r.big) return r;  // +-0 changes nothing
    const int ed = ef ? (int)ef : 1;
    const int shift = ed - eT;
    U q, g0, g1, q4;
    bool frac4;
    if (shift >= 0) {
        if (shift >= 2) {
            r.big = true;
            return r;
        }
        q = md << shift;
        g0 = g1 = q;
        q4 = q << 2;
        frac4 = false;
    } else {
        const int sh = -shift;
        if (sh > MB + 2) {  // |x| < 1/4, not a tie
            g0 = g1 = 0;
            q4 = 0;
            frac4 = true;
        } else {
            q = md >> sh;
            const U rem = md & ((U(1) << sh) - 1);
            const U half = U(1) << (sh - 1);
            if (rem > half) {
                g0 = g1 = q + 1;
            } else if (rem == half) {  // exact tie: the even neighbour of m + q + 1/2
                g0 = q + (q & 1);
                g1 = q + ((q + 1) & 1);
            } else {
                g0 = g1 = q;
            }
            if (sh >= 2) {
                q4 = md >> (sh - 2);
                frac4 = (md & ((U(1) << (sh - 2)) - 1)) != 0;
            } else {            // sh == 1: 4x = 2 md, an integer
                q4 = md << 1;
                frac4 = false;
            }
        }
    }
    if (!neg) {
        r.map.a0 = (I)g0;
        r.map.a1 = (I)g1;
        r.flq = (I)q4;
    } else {
        r.map.a0 = -(I)g0;
        r.map.a1 = -(I)g1;
        r.flq = -(I)q4 - (frac4 ? 1 : 0);
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

// ---------------------------------------------------------------------------------------------
// shared pieces
// ---------------------------------------------------------------------------------------------
template <typename I> __device__ __forceinline__ pmap<I> pmap_shfl_up(const pmap<I> v, int off)
{
    pmap<I> o;
    if constexpr (sizeof(I) == 8) {
        o.a0 = (I)__shfl_up_sync(0xffffffffu, (long long)v.a0, off);
        o.a1 = (I)__shfl_up_sync(0xffffffffu, (long long)v.a1, off);
    } else {
        o.a0 = __shfl_up_sync(0xffffffffu, v.a0, off);
        o.a1 = __shfl_up_sync(0xffffffffu, v.a1, off);
    }
    return o;
}
template <typename I> __device__ __forceinline__ I shfl_i(I v, int src)
{
    if constexpr (sizeof(I) == 8) return (I)__shfl_sync(0xffffffffu, (long long)v, src);
    else return __shfl_sync(0xffffffffu, v, src);
}

template <typename I> __device__ __forceinline__ I shfl_xor_i(I v, int mask)
{
    if constexpr (sizeof(I) == 8) return (I)__shfl_xor_sync(0xffffffffu, (long long)v, mask);
    else return __shfl_xor_sync(0xffffffffu, v, mask);
}

// inclusive scan of parity maps across the 32 lanes of a warp
template <typename I> __device__ __forceinline__ pmap<I> pmap_warp_scan(pmap<I> v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const pmap<I> o = pmap_shfl_up<I>(v, off);
        if (lane >= off) v = pmap_then<I>(o, v);
    }
    return v;
}

constexpr int NO_VIOL = 0x7fffffff;

// Walk E consecutive pending updates starting from the exact integer m.  Returns the index of
// the first update the binade identity does not cover (NO_VIOL if none); m is left at the value
// BEFORE that update (or after the last one).
template <typename T, int E>
__device__ __forceinline__ int walk_updates(const upd_map<T> (&um)[E], int first_idx, int pos, int nc,
                                            typename fp_bits<T>::I &m)
{
    using I = typename fp_bits<T>::I;
    constexpr I LO = I(1) << fp_bits<T>::MB;
    constexpr I HI = I(1) << (fp_bits<T>::MB + 1);
    int viol = NO_VIOL;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const int i = first_idx + k;
        if (i >= pos && i < nc && viol == NO_VIOL) {
            const I f = um[k].flq;
            if (um[k].big || m < LO || m >= HI || 4 * (m - LO) + (f < 0 ? f : 0) < -1 || m + (f >> 2) >= HI) viol = i;
            else m += (m & 1) ? um[k].map.a1 : um[k].map.a0;
        }
    }
    return viol;
}

// plain sequential adds from `p` while the value is not a positive normal number (at least one)
template <typename T, typename LOAD>
__device__ __forceinline__ T seq_until_ready(T cur, int &p, int nc, LOAD load)
{
    int e2;
    typename fp_bits<T>::I m2;
    do {
        cur = t_add<T>(cur, load(p));
        ++p;
    } while (p < nc && !scan_ready<T>(cur, e2, m2));
    return cur;
}


// ---------------------------------------------------------------------------------------------
// Huge segments (one node with tens of thousands of updates, e.g. the root): two phases so that
// all SMs work on one segment.
//   phase A (all CTAs): every 1024-update chunk is summarised, relative to the exponent eT of the
//       node's CURRENT value, as a parity map plus the extreme excursions of its exact partial
//       sums for either start parity.  A chunk is "safe" for an incoming m if those excursions keep
//       every intermediate value inside the binade -- then its whole effect is m += a[parity(m)].
//   phase B (one CTA per segment): walks the chunk summaries in order (a few integer ops per
//       chunk); an unsafe chunk, or any chunk after the exponent has changed, is folded exactly
//       by the CTA driver above.
// ---------------------------------------------------------------------------------------------
// Update stream of one node of a tiny (power-of-two) tree seen through a membership mask:
// element i counts for `node` iff node is an ancestor of leaf[i]; everything else reads as 0,
// which is the identity of the fold (T + 0 == T).  Used for the replicated top tree of the
// subtree-sharded layout, where all ranks' differences arrive as ONE batch-ordered vector.
// Sliced variant (slices layout of the sharded tree): the batch is distributed over the ranks, element i of the
// GLOBAL batch lives on rank i / span at local position i % span (positions >= m of a slice are padding and read as 0);
// d[r] / leaf[r] point into rank r's exchange buffer (peer memory), leaf ids are uint8.  Only chunks that must be
// folded one add at a time are ever read through it.
struct slice_ref {
    const void *const *d = nullptr;
    const void *const *leaf = nullptr;
    int64_t span = 0, m = 0;
};

template <typename T> struct masked_src {
    const T *d;
    const void *leaf;   // leaf (shard) id of every element: int64 (leaf_bytes 8) or uint8 (leaf_bytes 1)
    int leaf_bytes;
    int64_t nleaves;
    uint32_t node;
    int rshift;   // leaf level - level(node)
    slice_ref sl{};
    int64_t off = 0;   // sliced: global index of element 0
    __device__ __forceinline__ T operator[](int64_t i) const
    {
        if (sl.d) {
            const int64_t gi = off + i, r = gi / sl.span, lo = gi - r * sl.span;
            if (lo >= sl.m) return T(0);
            const int64_t l = reinterpret_cast<const uint8_t *>(sl.leaf[r])[lo];
            return ((uint32_t)((l + nleaves) >> rshift) == node) ? reinterpret_cast<const T *>(sl.d[r])[lo] : T(0);
        }
        const int64_t l = leaf_bytes == 1 ? (int64_t)__ldg((const uint8_t *)leaf + i) : __ldg((const int64_t *)leaf + i);
        return ((uint32_t)((l + nleaves) >> rshift) == node) ? d[i] : T(0);
    }
    __device__ __forceinline__ masked_src operator+(int64_t o) const
    {
        masked_src r = *this;
        if (sl.d) {
            r.off += o;
        } else {
            r.d += o;
            r.leaf = (const char *)leaf + o * leaf_bytes;
        }
        return r;
    }
};

template <typename I> struct chunk_rec {
    I a0, a1;        // parity map of the whole chunk
    I min0, max0;    // start parity even: min0 = min over k of 4 (m_k - m_in) + min(floor(4 x_k), 0)  (QUARTER units, see
                     // upd_map::flq), max0 = max over k of (m_k - m_in) + max(floor(x_k), 0)
    I min1, max1;    // ... start parity odd
    int ok;          // 0: contains an update the identity never covers, or magnitudes too large
    int pad;
};

template <typename T, typename SRC>
__device__ void chunk_summary_cta(SRC src, int nc, int eT, chunk_rec<typename fp_bits<T>::I> *out,
                                  typename fp_bits<T>::I (*s_wtot)[2], typename fp_bits<T>::I (*s_red)[4],
                                  float *s_abs)
{
    using I = typename fp_bits<T>::I;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FOLD_THREADS / 32;
    upd_map<T> um[FOLD_E];
    pmap<I> local{0, 0};
    bool anybig = false;
    float mag = 0.f;
#pragma unroll
    for (int k = 0; k < FOLD_E; ++k) {
        const int i = tid * FOLD_E + k;
        if (i < nc) um[k] = decompose<T>(src[i], eT);
        else um[k] = upd_map<T>{{0, 0}, 0, false};
        anybig |= um[k].big;
        const I a = um[k].map.a0 < 0 ? -um[k].map.a0 : um[k].map.a0;
        mag += (float)a + 2.f;
        local = pmap_then<I>(local, um[k].map);
    }
    // Every update absorbed (a0 == a1 == 0 for the whole chunk: the node is far larger than the differences, the usual
    // case for the upper levels of a big tree and for the top tree of a sharded one)?  Then the chunk's map is the
    // identity and the trajectory never moves: no scan, the record is two reductions.
    {
        bool moves = false;
        I mnq = 0, mxf = 0;
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) {
            moves |= (um[k].map.a0 | um[k].map.a1) != 0;
            const I f = um[k].flq;
            mnq = f < mnq ? f : mnq;
            const I fx = f > 0 ? (f >> 2) : I(0);
            mxf = fx > mxf ? fx : mxf;
        }
        if (!__syncthreads_or(moves)) {      // also the barrier that frees the shared buffers
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const I a = shfl_xor_i<I>(mnq, off), b = shfl_xor_i<I>(mxf, off);
                mnq = a < mnq ? a : mnq;
                mxf = b > mxf ? b : mxf;
            }
            const int bigs = __syncthreads_or(anybig);
            if (lane == 0) { s_red[warp][0] = mnq; s_red[warp][1] = mxf; }
            __syncthreads();
            if (tid == 0) {
                chunk_rec<I> r;
                r.a0 = 0; r.a1 = 0;
                r.min0 = s_red[0][0]; r.max0 = s_red[0][1];
                for (int w = 1; w < NW; ++w) {
                    r.min0 = s_red[w][0] < r.min0 ? s_red[w][0] : r.min0;
                    r.max0 = s_red[w][1] > r.max0 ? s_red[w][1] : r.max0;
                }
                r.min1 = r.min0; r.max1 = r.max0;
                r.ok = bigs ? 0 : 1;
                r.pad = 0;
                *out = r;
            }
            return;
        }
    }
    const pmap<I> incl = pmap_warp_scan<I>(local, lane);
    pmap<I> excl = pmap_shfl_up<I>(incl, 1);
    if (lane == 0) excl = pmap<I>{0, 0};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mag += __shfl_xor_sync(0xffffffffu, mag, off);
    __syncthreads();   // shared buffers free
    if (lane == 31) {
        s_wtot[warp][0] = incl.a0;
        s_wtot[warp][1] = incl.a1;
    }
    if (lane == 0) s_abs[warp] = mag;
    __syncthreads();
    pmap<I> pre{0, 0}, tot{0, 0};
    float magt = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const pmap<I> t{s_wtot[w][0], s_wtot[w][1]};
        if (w < warp) pre = pmap_then<I>(pre, t);
        tot = pmap_then<I>(tot, t);
        magt += s_abs[w];
    }
    pre = pmap_then<I>(pre, excl);
    // trajectories for both start parities, as offsets from the incoming m
    I off0 = pre.a0, off1 = pre.a1;
    I mn0 = 0, mx0 = 0, mn1 = 0, mx1 = 0;
#pragma unroll
    for (int k = 0; k < FOLD_E; ++k) {
        const int i = tid * FOLD_E + k;
        if (i < nc) {
            const I f = um[k].flq;
            const I lo0 = 4 * off0 + (f < 0 ? f : 0), hi0 = f > 0 ? off0 + (f >> 2) : off0;
            const I lo1 = 4 * off1 + (f < 0 ? f : 0), hi1 = f > 0 ? off1 + (f >> 2) : off1;
            mn0 = lo0 < mn0 ? lo0 : mn0;  mx0 = hi0 > mx0 ? hi0 : mx0;
            mn1 = lo1 < mn1 ? lo1 : mn1;  mx1 = hi1 > mx1 ? hi1 : mx1;
            off0 += (off0 & 1) ? um[k].map.a1 : um[k].map.a0;          // start parity even: m = m_in + off0
            off1 += ((off1 + 1) & 1) ? um[k].map.a1 : um[k].map.a0;    // start parity odd
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const I a = shfl_xor_i<I>(mn0, off), b = shfl_xor_i<I>(mx0, off);
        const I c = shfl_xor_i<I>(mn1, off), d = shfl_xor_i<I>(mx1, off);
        mn0 = a < mn0 ? a : mn0;  mx0 = b > mx0 ? b : mx0;
        mn1 = c < mn1 ? c : mn1;  mx1 = d > mx1 ? d : mx1;
    }
    const int bigs = __syncthreads_or(anybig);
    if (lane == 0) {
        s_red[warp][0] = mn0; s_red[warp][1] = mx0; s_red[warp][2] = mn1; s_red[warp][3] = mx1;
    }
    __syncthreads();
    if (tid == 0) {
        chunk_rec<I> r;
        r.a0 = tot.a0; r.a1 = tot.a1;
        r.min0 = s_red[0][0]; r.max0 = s_red[0][1]; r.min1 = s_red[0][2]; r.max1 = s_red[0][3];
        for (int w = 1; w < NW; ++w) {
            r.min0 = s_red[w][0] < r.min0 ? s_red[w][0] : r.min0;
            r.max0 = s_red[w][1] > r.max0 ? s_red[w][1] : r.max0;
            r.min1 = s_red[w][2] < r.min1 ? s_red[w][2] : r.min1;
            r.max1 = s_red[w][3] > r.max1 ? s_red[w][3] : r.max1;
        }
        // total magnitude below 2^(MB-1): no integer in the scan can have overflowed
        r.ok = (!bigs && magt < (float)(I(1) << (fp_bits<T>::MB - 1))) ? 1 : 0;
        r.pad = 0;
        *out = r;
    }
}

// ---------------------------------------------------------------------------------------------
// warp driver: one warp folds src[0..len) into acc; every lane returns the same result.
// s_w: this warp's private shared staging buffer of FOLDW_PAD elements.
// ---------------------------------------------------------------------------------------------
constexpr int FOLDW_E = 8;
constexpr int FOLDW_CHUNK = 32 * FOLDW_E;                    // 256
constexpr int FOLDW_PAD = FOLDW_CHUNK + FOLDW_CHUNK / 8;     // one pad word per 8: conflict-free

__device__ __forceinline__ int foldw_at(int i) { return i + (i >> 3); }

template <typename T>
__device__ T ordered_fold_warp(T cur, const T *__restrict__ src, int len, T *s_w)
{
    using I = typename fp_bits<T>::I;
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < len; base += FOLDW_CHUNK) {
        const int nc = len - base < FOLDW_CHUNK ? len - base : FOLDW_CHUNK;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < FOLDW_E; ++k) {
            const int i = lane + 32 * k;
            s_w[foldw_at(i)] = (i < nc) ? src[base + i] : T(0);
        }
        __syncwarp();
        T dreg[FOLDW_E];
#pragma unroll
        for (int k = 0; k < FOLDW_E; ++k) dreg[k] = s_w[foldw_at(lane * FOLDW_E + k)];
        auto load = [&](int p) { return s_w[foldw_at(p)]; };
        int pos = 0;
        while (pos < nc) {
            int eT;
            I m0;
            if (!scan_ready<T>(cur, eT, m0)) {
                cur = seq_until_ready<T>(cur, pos, nc, load);   // identical in every lane
                continue;
            }
            upd_map<T> um[FOLDW_E];
            pmap<I> local{0, 0};
#pragma unroll
            for (int k = 0; k < FOLDW_E; ++k) {
                const int i = lane * FOLDW_E + k;
                if (i >= pos && i < nc) um[k] = decompose<T>(dreg[k], eT);
                else um[k] = upd_map<T>{{0, 0}, 0, false};
                local = pmap_then<I>(local, um[k].map);
            }
            const pmap<I> incl = pmap_warp_scan<I>(local, lane);
            pmap<I> excl = pmap_shfl_up<I>(incl, 1);
            if (lane == 0) excl = pmap<I>{0, 0};
            I m = m0 + ((m0 & 1) ? excl.a1 : excl.a0);
            const int viol = walk_updates<T, FOLDW_E>(um, lane * FOLDW_E, pos, nc, m);
            const int fv = __reduce_min_sync(0xffffffffu, viol);
            if (fv == NO_VIOL) {
                const I t0 = shfl_i<I>(incl.a0, 31), t1 = shfl_i<I>(incl.a1, 31);
                cur = value_of<T>(eT, m0 + ((m0 & 1) ? t1 : t0));
                pos = nc;
            } else {
                const I mv = shfl_i<I>(m, fv / FOLDW_E);   // exact m just before update fv
                cur = t_add<T>(value_of<T>(eT, mv), load(fv));
                int p = fv + 1;
                const int stop = (p + FOLD_SEQ_RUN < nc) ? p + FOLD_SEQ_RUN : nc;
                for (; p < stop; ++p) cur = t_add<T>(cur, load(p));
                pos = p;
            }
        }
    }
    return cur;
}

// ---------------------------------------------------------------------------------------------
// CTA driver: FOLD_THREADS threads fold src[0..len) into acc; every thread returns the result.
// Each thread owns FOLD_E consecutive updates of a chunk (loaded straight from global, next
// chunk prefetched); two barriers per chunk.
// ---------------------------------------------------------------------------------------------
struct fold_smem_ctl {
    int first_viol;
    int owner_set;
    unsigned int slot;
    int pad;
};

template <typename T, typename SRC>
__device__ T ordered_fold_cta(T cur, SRC src, int len,
                              typename fp_bits<T>::I (*s_wtot)[2], typename fp_bits<T>::I *s_mviol,
                              fold_smem_ctl *ctl, T *s_seq /* FOLD_CHUNK elements of shared memory */)
{
    using I = typename fp_bits<T>::I;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FOLD_THREADS / 32;
    auto load = [&](int p) { return src[p]; };   // sequential paths read global/L1 directly

    T nxt[FOLD_E];
#pragma unroll
    for (int k = 0; k < FOLD_E; ++k) {
        const int i = tid * FOLD_E + k;
        nxt[k] = (i < len) ? src[i] : T(0);
    }
    for (int base = 0; base < len; base += FOLD_CHUNK) {
        const int nc = len - base < FOLD_CHUNK ? len - base : FOLD_CHUNK;
        T dreg[FOLD_E];
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) dreg[k] = nxt[k];
        // staged for the sequential runs after a restart (read only behind barriers 1 and 2 below;
        // the previous chunk's readers are behind the barrier that ended their round)
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) s_seq[tid * FOLD_E + k] = dreg[k];
#pragma unroll
        for (int k = 0; k < FOLD_E; ++k) {   // prefetch the next chunk
            const int i = base + FOLD_CHUNK + tid * FOLD_E + k;
            nxt[k] = (i < len) ? src[i] : T(0);
        }
        auto loadc = [&](int p) { return load(base + p); };
        int pos = 0;
        int nviol = 0;                        // restarts within this chunk
        while (pos < nc) {                    // pos, cur are uniform across the CTA
            int eT;
            I m0;
            if (!scan_ready<T>(cur, eT, m0)) {
                cur = seq_until_ready<T>(cur, pos, nc, loadc);   // identical in every thread
                continue;
            }
            upd_map<T> um[FOLD_E];
            pmap<I> local{0, 0};
#pragma unroll
            for (int k = 0; k < FOLD_E; ++k) {
                const int i = tid * FOLD_E + k;
                if (i >= pos && i < nc) um[k] = decompose<T>(dreg[k], eT);
                else um[k] = upd_map<T>{{0, 0}, 0, false};
                local = pmap_then<I>(local, um[k].map);
            }
            const pmap<I> incl = pmap_warp_scan<I>(local, lane);
            pmap<I> excl = pmap_shfl_up<I>(incl, 1);
            if (lane == 0) excl = pmap<I>{0, 0};
            if (lane == 31) {
                s_wtot[warp][0] = incl.a0;
                s_wtot[warp][1] = incl.a1;
            }
            if (tid == 0) ctl->first_viol = NO_VIOL;
            __syncthreads();                                           // barrier 1: warp totals
            pmap<I> pre{0, 0}, tot{0, 0};
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const pmap<I> t{s_wtot[w][0], s_wtot[w][1]};
                if (w < warp) pre = pmap_then<I>(pre, t);
                tot = pmap_then<I>(tot, t);
            }
            pre = pmap_then<I>(pre, excl);
            I m = m0 + ((m0 & 1) ? pre.a1 : pre.a0);
            const int viol = walk_updates<T, FOLD_E>(um, tid * FOLD_E, pos, nc, m);
            if (!__syncthreads_or(viol != NO_VIOL)) {                  // barrier 2: any violation?
                cur = value_of<T>(eT, m0 + ((m0 & 1) ? tot.a1 : tot.a0));
                pos = nc;
                continue;
            }
            if (viol != NO_VIOL) atomicMin(&ctl->first_viol, viol);
            __syncthreads();
            const int fv = ctl->first_viol;
            if (viol == fv) *s_mviol = m;      // exact m just before update fv
            __syncthreads();
            cur = t_add<T>(value_of<T>(eT, *s_mviol), s_seq[fv]);
            int p = fv + 1;
            // A value hovering at a binade boundary violates again and again, and every restart costs
            // three barriers: the sequential run grows 32, 128, 512, rest, so a chunk restarts at
            // most four times and its worst case is one plain sequential pass.
            const int run = FOLD_SEQ_RUN << (nviol < 3 ? 2 * nviol : 6);
            ++nviol;
            const int stop = (p + run < nc) ? p + run : nc;
#pragma unroll 16
            for (; p < stop; ++p) cur = t_add<T>(cur, s_seq[p]);   // uniform address: a broadcast, no conflicts
            pos = p;
            __syncthreads();                   // everyone has read s_mviol / first_viol
        }
    }
    return cur;
}

// Plain sequential fold of one chunk (every thread computes the same result).  For chunks whose
// summary says "stays representable but touches the binade boundary": a value hovering there
// defeats the scan (restart after restart), while the dependent-add chain is ~4 cycles per update.
template <typename T, typename SRC>
__device__ T sequential_fold_cta(T cur, SRC src, int nc, T *s_seq)
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < FOLD_E; ++k) {
        const int i = tid * FOLD_E + k;
        s_seq[i] = (i < nc) ? src[i] : T(0);
    }
    __syncthreads();
#pragma unroll 16
    for (int p = 0; p < nc; ++p) cur = t_add<T>(cur, s_seq[p]);
    __syncthreads();
    return cur;
}

}  // namespace st
