// st_update.cu -- batched leaf update with exact ancestor propagation.
//
// Replaces update_sumtree (reference starr/src/_sumtree_array.pyx:22-49):
//
//     for j in 0..ni-1:  d = vals[j % nv] - leaf[idx[j]];  leaf += d;  every ancestor += d
//
// The reference applies the updates one after another, so for floating point the result at
// every node is the ORDERED fold  T = (((T + d_a) + d_b) + ...)  over the updates below
// that node in batch order.  Nodes decouple, which gives the exact parallel
// form used here:
//
//   1. normalised heap keys (leaf position left-aligned to `depth` bits) + stable radix sort
//      by key with the batch position j as payload  -> duplicates of one leaf are adjacent
//      and in batch order;
//   2. leaf pass: one thread per run of equal leaves walks the run sequentially and emits
//      every d_j (this is where `leaf = fl(leaf + fl(val - leaf))` happens);
//   3a. integer dtypes: adds are modular, so order is irrelevant -> every element walks to the
//      root; lanes that share an ancestor are merged with a segmented warp scan and one
//      atomic add per (warp, node) is issued.
//   3b. float dtypes: top-down, one level at a time, the batch is kept ordered by
//      (ancestor at that level, j) -- a stable MSD split (st_partition.cuh, one cooperative
//      launch) -- so every node's updates are one contiguous, batch-ordered segment.  Segments
//      are classified by length and queued:
//        retire (<= 512)   finished for the node AND all deeper levels by ordered atomic adds,
//                          one warp per segment, lane = level (retire_kernel);
//        medium / long     exact parallel ordered fold by a warp / a CTA (fold_scan_kernel,
//                          st_ordered_fold.cuh);
//        huge (> 16384)    chunk summaries on all CTAs (fold_huge_maps_kernel), stitched per
//                          segment inside the fold_scan_kernel launch; segments whose node leaves
//                          its binade get a second pass (two launches, usually empty).
//      Batches of <= 4096 updates run the whole pipeline in one CTA (st_update_small.cuh).
//
// Everything is enqueued on one stream; validation failures set the device status word and
// every later kernel of the batch returns immediately (nothing is written).
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "st_common.cuh"
#include "st_update_defs.cuh"
#include "st_sidx.cuh"
#include "st_ordered_fold.cuh"
#include "st_partition_hybrid.cuh"
#include "st_update_small.cuh"

namespace st {

int64_t update_max_chunk() { return UPDATE_MAX_CHUNK; }

// -----------------------------------------------------------------------------------------
// workspace carving
// -----------------------------------------------------------------------------------------
struct update_ws {
    uint32_t *keys_a, *keys_b, *kj;
    int *j_a, *j_b;
    void *dj, *ds, *lev_d;  // T-typed
    seg_entry *queue, *queue_med, *queue_huge;
    int *huge_off;
    void *huge_recs;
    void *huge_recs2;   // summaries for the binade a segment moved to (second pass)
    int *huge_redo;     // [3][slots]: first chunk of the second pass (-1: none), binade of the first pass, other binade
    int *seg_sa, *seg_ea, *seg_sb, *seg_eb, *zcount;
    uint32_t *lev_key;
    seg_entry *queue_ret;
    unsigned int *tilesum;
    unsigned int *hy_hist, *hy_bintotal;   // hybrid partition (st_partition_hybrid.cuh)
    int *hy_segstart, *hy_seglen;
    unsigned char *hy_segstate;
    unsigned int *bigmap;      // bit per leaf group: more than DEEP_GROUP_MAX updates (ancestors not by deep_red_kernel)
    size_t bigmap_bytes;
    void *cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static update_ws carve(const st_tree *t, int64_t B, char *base)
{
    update_ws w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += align_up(bytes);
        return (void *)p;
    };
    const size_t es = (size_t)t->esize;
    const bool is_f = (t->dtype == ST_F32 || t->dtype == ST_F64);
    w.keys_a = (uint32_t *)take(B * 4);
    w.keys_b = (uint32_t *)take(B * 4);
    w.kj = (uint32_t *)take(B * 4);
    w.j_a = (int *)take(B * 4);
    w.j_b = (int *)take(B * 4);
    w.dj = take(B * es);
    w.ds = take(B * es);
    w.lev_d = is_f ? take((size_t)(t->depth + 1) * B * es) : nullptr;
    w.queue = is_f ? (seg_entry *)take((size_t)(t->depth + 1) * (B / (MEDIUM_SEG + 1) + 2) * sizeof(seg_entry)) : nullptr;
    w.queue_med = is_f ? (seg_entry *)take((size_t)(t->depth + 1) * (B / (SHORT_SEG + 1) + 2) * sizeof(seg_entry)) : nullptr;
    if (is_f) {
        w.queue_huge = (seg_entry *)take((size_t)(t->depth + 1) * (B / (HUGE_SEG + 1) + 2) * sizeof(seg_entry));
        w.huge_off = (int *)take((size_t)(t->depth + 1) * (B / (HUGE_SEG + 1) + 2) * sizeof(int));
        w.huge_recs = take((size_t)(t->depth + 1) * (B / FOLD_CHUNK + 2) * 64);
        w.huge_recs2 = take((size_t)(t->depth + 1) * (B / FOLD_CHUNK + 2) * 64);
        w.huge_redo = (int *)take(3 * (size_t)(t->depth + 1) * (B / (HUGE_SEG + 1) + 2) * sizeof(int));
        w.seg_sa = (int *)take(B * 4);
        w.seg_ea = (int *)take(B * 4);
        w.seg_sb = (int *)take(B * 4);
        w.seg_eb = (int *)take(B * 4);
        w.zcount = (int *)take(B * 4);
        w.lev_key = (uint32_t *)take((size_t)(t->depth + 1) * B * 4);
        w.queue_ret = (seg_entry *)take((size_t)(B + 32) * sizeof(seg_entry));
        w.tilesum = (unsigned int *)take(2 * PART_MAX_TILES * sizeof(unsigned int));
        w.hy_hist = (unsigned int *)take((size_t)HY_MAX_BINS * HY_MAX_TILES * sizeof(unsigned int));
        w.hy_bintotal = (unsigned int *)take(HY_MAX_BINS * sizeof(unsigned int));
        w.hy_segstart = (int *)take(2 * HY_MAX_BINS * sizeof(int));
        w.hy_seglen = (int *)take(2 * HY_MAX_BINS * sizeof(int));
        w.hy_segstate = (unsigned char *)take(2 * HY_MAX_BINS);
        // groups: fewer than 4 * B when the group size is maximal, fewer than 2n >> 10 when it is capped (deep_group_bits)
        w.bigmap_bytes = (size_t)(B / 2 + t->n / 4096 + 256);
        w.bigmap = (unsigned int *)take(w.bigmap_bytes);
    }
    if (t->cub_cap != B) {
        // temp-storage queries are host-side work: do them once per capacity, not once per call
        size_t cb1 = 0, cb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, cb1, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (const int *)nullptr, (int *)nullptr, (int)B, 0, 32, (cudaStream_t)0);
        if (t->dtype == ST_F64)
            cub::DeviceRadixSort::SortPairs(nullptr, cb2, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (const double *)nullptr, (double *)nullptr, (int)B, 0, 32, (cudaStream_t)0);
        else if (t->dtype == ST_F32)
            cub::DeviceRadixSort::SortPairs(nullptr, cb2, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (const float *)nullptr, (float *)nullptr, (int)B, 0, 32, (cudaStream_t)0);
        const_cast<st_tree *>(t)->cub_bytes = cb1 > cb2 ? cb1 : cb2;
        const_cast<st_tree *>(t)->cub_cap = B;
    }
    w.cub_bytes = t->cub_bytes;
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

size_t update_workspace_bytes(const st_tree *t, int64_t batch)
{
    if (batch > UPDATE_MAX_CHUNK) batch = UPDATE_MAX_CHUNK;
    if (batch < 1) batch = 1;
    return carve(t, batch, nullptr).total;
}

// -----------------------------------------------------------------------------------------
// 0. validation (pyx:37-39 negative scan over ALL vals; Cython boundscheck on idx)
// -----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) validate_kernel(const int64_t *__restrict__ idx, int64_t ni,
                                                       const T *__restrict__ vals, int64_t nv, int64_t n,
                                                       unsigned int *status)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;
    if constexpr (is_signed_t<T>::value)
        for (int64_t i = tid; i < nv; i += stride)
            if (vals[i] < T(0)) bad |= STF_NEG_VALS;
    for (int64_t i = tid; i < ni; i += stride) {
        const int64_t l = idx[i];
        if (l < 0 || l >= n) bad |= STF_INDEX;
    }
    if (bad) atomicOr(status + SW_STATUS, bad);
}

__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t *__restrict__ idx, int m, uint32_t n,
                                                        int depth, uint32_t *__restrict__ keys,
                                                        uint32_t *__restrict__ kj, int *__restrict__ jj,
                                                        unsigned int *status)
{
    if (status[SW_STATUS]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) { status[SW_QCOUNT] = 0; status[SW_QTAKE] = 0; status[SW_QCOUNT_MED] = 0; status[SW_QTAKE_MED] = 0;
                  status[SW_QCOUNT_HUGE] = 0; status[SW_HUGE_CHUNKS] = 0; status[SW_QTAKE_HUGE] = 0;
                  status[SW_ANYLONG] = 0; status[SW_ANYLONG + 1] = 0; status[SW_QCOUNT_RET] = 0; status[SW_REDO_ANY] = 0; status[SW_ANYBIG] = 0; }
    if (j >= m) return;
    const uint32_t h = (uint32_t)idx[j] + n;
    const uint32_t key = h << (depth - heap_level(h));
    keys[j] = key;
    kj[j] = key;
    jj[j] = j;
}

// -----------------------------------------------------------------------------------------
// 2. leaf pass: sequential chain per leaf (pyx:43-45), emits d in batch order and sorted order
// -----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) leaf_pass_kernel(T *__restrict__ H, uint32_t n, const uint32_t *__restrict__ keys,
                                                        const int *__restrict__ js, int m,
                                                        const T *__restrict__ vals, uint64_t nv, uint64_t j0,
                                                        T *__restrict__ dj, T *__restrict__ ds,
                                                        unsigned int *status, bool assign)
{
    if (status[SW_STATUS]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t key = keys[i];
    if (i > 0 && keys[i - 1] == key) return;  // not the head of its run
    const uint64_t two_n = 2ull * n;
    const uint32_t h = key_to_heap(key, two_n);
    T cur = H[h];
    int e = i;
    unsigned int nzero = 0;
    do {
        const uint32_t j = (uint32_t)js[e];
        const uint64_t g = j0 + j;  // position in the whole batch (pyx:44: vals[i % nv])
        const T v = vals[g < nv ? g : g % nv];
        const T d = t_sub<T>(v, cur);
        cur = assign ? v : t_add<T>(cur, d);   // assign: the experimental C++ twin stores val itself (cpp:22-23)
        dj[j] = d;
        ds[e] = d;
        nzero += (d == T(0));
        ++e;
    } while (e < m && keys[e] == key);
    H[h] = cur;
    if (nzero) atomicAdd(status + SW_STAT_ZERO_DELTA, nzero);   // bench.py's zero_delta_frac (rare on real batches)
}

// Float path: the sort only has to bring the duplicates of one leaf together in batch order, so the
// low `gb` key bits may be left unsorted.  A group = the updates of the 2^gb neighbouring leaves below one
// level-(depth-gb) node, batch-ordered; its head thread walks it (the running value of the leaf it touched
// last stays in a register, other leaves go through memory: its own earlier stores are visible to its
// later loads).  gb is chosen so that a group holds at most one update on average.
// Groups with more than DEEP_GROUP_MAX updates are flagged in `bigmap` (bit per group): their ancestors are
// left to the partition / fold / retirement machinery; all other groups get theirs from deep_red_kernel.
constexpr int DEEP_GROUP_MAX = 64;   // <= RETIRE_SEG, so that an unflagged group never holds a fold segment

template <typename T>
__global__ void __launch_bounds__(256) leaf_pass_grouped_kernel(T *H, uint32_t n, const uint32_t *__restrict__ keys,
                                                                const int *__restrict__ js, int m,
                                                                const T *__restrict__ vals, uint64_t nv, uint64_t j0,
                                                                T *__restrict__ dj, T *__restrict__ ds, int gb, int depth,
                                                                unsigned int *__restrict__ bigmap, unsigned int *status,
                                                                bool assign)
{
    if (status[SW_STATUS]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t grp = keys[i] >> gb;
    if (i > 0 && (keys[i - 1] >> gb) == grp) return;  // not the head of its group
    const uint64_t two_n = 2ull * n;
    int e = i;
    uint32_t key = keys[i];
    unsigned int nzero = 0;
    uint32_t h_last = 0;      // 0 is never a leaf
    T v_last = T(0);
    do {
        const uint32_t h = key_to_heap(key, two_n);
        const uint32_t j = (uint32_t)js[e];
        const uint64_t g = j0 + j;  // position in the whole batch (pyx:44: vals[i % nv])
        const T v = vals[g < nv ? g : g % nv];
        const T cur = (h == h_last) ? v_last : H[h];
        const T d = t_sub<T>(v, cur);
        v_last = assign ? v : t_add<T>(cur, d);
        h_last = h;
        H[h] = v_last;
        dj[j] = d;
        ds[e] = d;
        nzero += (d == T(0));
        ++e;
        if (e >= m) break;
        key = keys[e];
    } while ((key >> gb) == grp);
    if (bigmap && e - i > DEEP_GROUP_MAX) {
        const uint32_t gi = grp - (1u << (depth - gb));
        atomicOr(bigmap + (gi >> 5), 1u << (gi & 31));
        status[SW_ANYBIG] = 1;
    }
    if (nzero) atomicAdd(status + SW_STAT_ZERO_DELTA, nzero);
}

// Ancestors at the levels depth-gb .. leaf level - 1 ("deep" levels: the ones that do not fit the L2) of every
// group the leaf pass did not flag: all updates of such a node belong to ONE group, so the group's head thread
// adds them in batch order (ordered_add: same-thread same-address atomics keep program order).  Runs in
// leaf-sorted order, so consecutive threads touch neighbouring sectors of every level (DRAM pages are reused),
// which the batch-ordered retirement could not do for these levels.
template <typename T, bool SIDX>
__global__ void __launch_bounds__(256) deep_red_kernel(T *__restrict__ H, uint32_t n, int depth, int gb,
                                                       const uint32_t *__restrict__ keys, const T *__restrict__ ds, int m,
                                                       const unsigned int *__restrict__ status, const sidx_ref sx)
{
    if (status[SW_STATUS]) return;
    const uint64_t two_n = 2ull * n;
    const int top = depth - gb;                          // level of the group's node
    // grid-stride: the launch is kept small on purpose (it shares the SMs with the partition on the main stream)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        uint32_t key = keys[i];
        const uint32_t grp = key >> gb;
        if (i > 0 && (keys[i - 1] >> gb) == grp) continue;  // not the head of its group
        int len = 1;
        while (len <= DEEP_GROUP_MAX && i + len < m && (keys[i + len] >> gb) == grp) ++len;
        if (len > DEEP_GROUP_MAX) continue;                  // flagged group: not ours
        for (int e = i; e < i + len; ++e) {
            key = keys[e];
            const T d = ds[e];
            for (int l = key_leaf_level(key, two_n, depth) - 1; l >= top; --l) {
                const uint32_t v = key >> (depth - l);
                ordered_add<T>(H + v, d);
                if constexpr (SIDX) {
                    if (T *c = sidx_slot<T>(sx, v, l)) ordered_add<T>(c, d);   // the search index copy: same adds, same order
                }
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// 3a. integer propagation: segmented warp scan over equal ancestors + one atomic per segment
// -----------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void atomic_add_wrap(T *p, T v)
{
    if constexpr (sizeof(T) == 4) {
        atomicAdd(reinterpret_cast<unsigned int *>(p), (unsigned int)v);
    } else if constexpr (sizeof(T) == 8) {
        atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
    } else {
        // 8/16-bit lanes inside an aligned 32-bit word: CAS loop
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        unsigned int *w = reinterpret_cast<unsigned int *>(a & ~(uintptr_t)3);
        const unsigned int sh = (unsigned int)(a & 3) * 8;
        const unsigned int mask = (sizeof(T) == 1 ? 0xffu : 0xffffu) << sh;
        unsigned int old = *w, assumed;
        do {
            assumed = old;
            const unsigned int field = (assumed & mask) >> sh;
            const unsigned int nf = ((field + (unsigned int)(typename std::make_unsigned<T>::type)v) << sh) & mask;
            old = atomicCAS(w, assumed, (assumed & ~mask) | nf);
        } while (old != assumed);
    }
}

template <typename T> __device__ __forceinline__ T shfl_up_t(T v, int off)
{
    if constexpr (sizeof(T) == 8) {
        return (T)__shfl_up_sync(0xffffffffu, (unsigned long long)v, off);
    } else {
        return (T)__shfl_up_sync(0xffffffffu, (unsigned int)v, off);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) int_propagate_kernel(T *__restrict__ H, uint32_t n, int depth,
                                                            const uint32_t *__restrict__ keys,
                                                            const T *__restrict__ ds, int m,
                                                            const unsigned int *status, const sidx_ref sx)
{
    if (status[SW_STATUS]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim multiple of 32; whole warps stay
    const int lane = threadIdx.x & 31;
    const uint64_t two_n = 2ull * n;
    const bool live = i < m;
    const uint32_t key = live ? keys[i] : 0u;
    const T d = live ? ds[i] : T(0);
    const int leaf_level = live ? key_leaf_level(key, two_n, depth) : 0;
    for (int l = depth - 1; l >= 0; --l) {
        const uint32_t node = (l < leaf_level) ? (key >> (depth - l)) : 0u;  // 0 = no ancestor here
        T sum = node ? d : T(0);
        // segmented inclusive scan over RUNS of equal nodes (keys need not be sorted: in "apply
        // deltas" mode they arrive in batch order, and equal nodes may be separated by others)
        const uint32_t nprev = __shfl_up_sync(0xffffffffu, node, 1);
        unsigned int head = (lane == 0) || (nprev != node);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const T up = shfl_up_t<T>(sum, off);
            const unsigned int hup = __shfl_up_sync(0xffffffffu, head, off);
            if (lane >= off && !head) {
                sum = (T)(sum + up);
                head |= hup;
            }
        }
        const uint32_t nnext = __shfl_down_sync(0xffffffffu, node, 1);
        const bool tail = (lane == 31) || (nnext != node);
        if (node && tail && sum != T(0)) {
            atomic_add_wrap<T>(H + node, sum);
            if (T *c = sidx_slot<T>(sx, node, l)) atomic_add_wrap<T>(c, sum);   // search index copy (modular adds commute)
        }
    }
}

// -----------------------------------------------------------------------------------------
// 3b. float propagation: ordered segment folds (segments come from st_partition.cuh)
// -----------------------------------------------------------------------------------------
template <typename T, bool SIDX>
__device__ __forceinline__ void retire_warps(T *__restrict__ H, uint32_t n, int depth, const uint32_t *__restrict__ key0,
                                             const uint32_t *__restrict__ lev_key, const T *__restrict__ lev_d,
                                             int64_t lev_stride, const seg_entry *__restrict__ qret, unsigned int total,
                                             unsigned int warp_id, unsigned int warps, const sidx_ref &sx,
                                             int deep_level = 1 << 20, int gb = 0,
                                             const unsigned int *__restrict__ bigmap = nullptr);

// Queued segments: exact parallel ordered folds (st_ordered_fold.cuh).  CTAs first drain the
// queue of long segments (one CTA each), then their warps drain the queue of medium segments
// (one warp each).
template <typename T, bool MASKED, bool REDO>
__device__ void huge_apply_body(T *__restrict__ H, const T *__restrict__ lev_d, int64_t level_stride,
                                const seg_entry *__restrict__ qhuge, const int *__restrict__ huge_off,
                                const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs, unsigned int *status,
                                const void *__restrict__ mleaf, int mleaf_bytes, int64_t mnleaves, int mleaf_level,
                                const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs2, int *__restrict__ redo,
                                int redo_stride, T *s_seq, const sidx_ref &sx, const slice_ref &sl = slice_ref{});

// second-pass stitch of the huge segments, done by the same launch as the long / medium folds
template <typename T> struct huge_params {
    const seg_entry *qhuge;   // nullptr: no huge phase in this launch
    const int *huge_off;
    const chunk_rec<typename fp_bits<T>::I> *recs, *recs2;
    int *redo;
    int redo_stride;
};

// One launch for everything that is left after partition + retirement + the first-pass stitch.  CTA b
// first finishes huge segment b if the first pass handed it over (its node sits at a binade boundary:
// exact folds, the slow part), every other CTA goes straight on to the long and medium queues, so
// the slow stitch overlaps with the rest instead of holding up a kernel of its own.
template <typename T>
__global__ void __launch_bounds__(FOLD_THREADS) fold_scan_kernel(T *__restrict__ H, const T *__restrict__ lev_d,
                                                                 int64_t level_stride,
                                                                 const seg_entry *__restrict__ queue_long,
                                                                 const seg_entry *__restrict__ queue_med,
                                                                 unsigned int *status, retire_params<T> ret,
                                                                 huge_params<T> hp, const sidx_ref sx)
{
    if (status[SW_STATUS]) return;
    using I = typename fp_bits<T>::I;
    __shared__ I s_wtot[FOLD_THREADS / 32][2];
    __shared__ I s_mviol;
    __shared__ fold_smem_ctl ctl;
    __shared__ T s_stage[FOLD_THREADS / 32][FOLDW_PAD];
    if (hp.qhuge)
        huge_apply_body<T, false, true>(H, lev_d, level_stride, hp.qhuge, hp.huge_off, hp.recs, status, nullptr, 0, 0, 0,
                                        hp.recs2, hp.redo, hp.redo_stride, &s_stage[0][0], sx);
    if (ret.qret) {
        // small-batch path: retire the segments the single-CTA kernel queued (st_update_small.cuh)
        retire_warps<T, true>(H, ret.n, ret.depth, ret.key0, ret.lev_key, lev_d, level_stride, ret.qret,
                        status[SW_QCOUNT_RET], blockIdx.x * (FOLD_THREADS >> 5) + (threadIdx.x >> 5),
                        gridDim.x * (FOLD_THREADS >> 5), sx);
    }
    const unsigned int n_long = status[SW_QCOUNT];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) ctl.slot = atomicAdd(status + SW_QTAKE, 1u);
        __syncthreads();
        const unsigned int slot = ctl.slot;
        if (slot >= n_long) break;
        const seg_entry s = queue_long[slot];
        const T *src = lev_d + (int64_t)s.level * level_stride + s.start;
        const T r = ordered_fold_cta<T>(H[s.node], src, s.len, s_wtot, &s_mviol, &ctl, &s_stage[0][0]);
        if (threadIdx.x == 0) sidx_store<T>(H, sx, s.node, s.level, r);
    }
    const unsigned int n_med = status[SW_QCOUNT_MED];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(status + SW_QTAKE_MED, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n_med) break;
        const seg_entry s = queue_med[slot];
        const T *src = lev_d + (int64_t)s.level * level_stride + s.start;
        const T r = ordered_fold_warp<T>(H[s.node], src, s.len, s_stage[warp]);
        if (lane == 0) sidx_store<T>(H, sx, s.node, s.level, r);
    }
}

// Retirement (queued by the partition kernel): a segment of <= RETIRE_SEG updates, batch-ordered, is
// finished for its node AND every deeper level with ordered atomic adds -- no further partitioning,
// no scan.  One warp per segment, LANE = LEVEL: lane k owns level l + k.  The warp streams the
// segment 32 updates at a time (coalesced loads), broadcasts them one by one (shuffles), and every
// lane adds the update to its level's ancestor.  All adds to one node therefore come from one lane
// in batch order (ordered_add); the random node updates of a 2^20 batch spread over every SM at
// full occupancy instead of stalling the cooperative kernel.
template <typename T, bool SIDX>
__device__ __forceinline__ void retire_warps(T *__restrict__ H, uint32_t n, int depth,
                                             const uint32_t *__restrict__ key0,
                                             const uint32_t *__restrict__ lev_key,
                                             const T *__restrict__ lev_d, int64_t lev_stride,
                                             const seg_entry *__restrict__ qret, unsigned int total,
                                             unsigned int warp_id, unsigned int warps, const sidx_ref &sx, int deep_level,
                                             int gb, const unsigned int *__restrict__ bigmap)
{
    // levels >= deep_level belong to deep_red_kernel, except below the groups the leaf pass flagged in bigmap
    // (bigmap == nullptr: no group was flagged in this batch)
    const uint64_t two_n = 2ull * n;
    const int lane = threadIdx.x & 31;
    for (unsigned int w = warp_id; w < total; w += warps) {
        const seg_entry s = qret[w];
        const int l = s.level, len = s.len;
        const uint32_t *kp = (l == 0 ? key0 : lev_key + (int64_t)l * lev_stride) + s.start;
        const T *dp = lev_d + (int64_t)l * lev_stride + s.start;
        const int l2 = l + lane;              // this lane's level
        const bool have_level = l2 < depth;
        const int sh = depth - l2;
        for (int base = 0; base < len; base += 32) {
            uint32_t ky = 0;
            T dv = T(0);
            if (base + lane < len) {
                ky = kp[base + lane];
                dv = dp[base + lane];
            }
            const int cnt = len - base < 32 ? len - base : 32;
            for (int x = 0; x < cnt; ++x) {
                const uint32_t kx = __shfl_sync(0xffffffffu, ky, x);
                T dx;
                if constexpr (sizeof(T) == 8) dx = __longlong_as_double(__shfl_sync(0xffffffffu, __double_as_longlong(dv), x));
                else dx = __shfl_sync(0xffffffffu, dv, x);
                if (have_level && l2 < key_leaf_level(kx, two_n, depth)) {
                    bool mine = l2 < deep_level;
                    if (!mine && bigmap) {
                        const uint32_t gi = (kx >> gb) - (1u << (depth - gb));
                        mine = (bigmap[gi >> 5] >> (gi & 31)) & 1u;
                    }
                    if (mine) {
                        ordered_add<T>(H + (kx >> sh), dx);
                        if constexpr (SIDX) {
                            if (T *c = sidx_slot<T>(sx, kx >> sh, l2)) ordered_add<T>(c, dx);   // search index copy
                        }
                    }
                }
            }
        }
    }
}

template <typename T, bool SIDX>
__global__ void __launch_bounds__(256) retire_kernel(T *__restrict__ H, uint32_t n, int depth,
                                                     const uint32_t *__restrict__ key0,
                                                     const uint32_t *__restrict__ lev_key,
                                                     const T *__restrict__ lev_d, int64_t lev_stride,
                                                     const seg_entry *__restrict__ qret, const unsigned int *status,
                                                     int deep_level, int gb, const unsigned int *__restrict__ bigmap,
                                                     const sidx_ref sx)
{
    if (status[SW_STATUS]) return;
    retire_warps<T, SIDX>(H, n, depth, key0, lev_key, lev_d, lev_stride, qret, status[SW_QCOUNT_RET],
                          blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5), sx, deep_level, gb,
                    status[SW_ANYBIG] ? bigmap : nullptr);
}

// Huge segments, phase A: every CTA summarises chunks (see st_ordered_fold.cuh).
template <typename T, bool MASKED>
__global__ void __launch_bounds__(FOLD_THREADS) fold_huge_maps_kernel(const T *__restrict__ H, const T *__restrict__ lev_d,
                                                                      int64_t level_stride,
                                                                      const seg_entry *__restrict__ qhuge,
                                                                      const int *__restrict__ huge_off,
                                                                      chunk_rec<typename fp_bits<T>::I> *__restrict__ recs,
                                                                      const unsigned int *status,
                                                                      const void *__restrict__ mleaf, int mleaf_bytes, int64_t mnleaves,
                                                                      int mleaf_level, const int *__restrict__ redo_from,
                                                                      int redo_stride)
{
    // redo_from != nullptr: second pass -- only the chunks [redo_from[e], ...) of the segments whose
    // node left the binade its first-pass summaries were made for (H[node] holds the value reached).
    if (status[SW_STATUS]) return;
    if (redo_from && !status[SW_REDO_ANY]) return;
    using I = typename fp_bits<T>::I;
    __shared__ I s_wtot[FOLD_THREADS / 32][2];
    __shared__ I s_red[FOLD_THREADS / 32][4];
    __shared__ float s_abs[FOLD_THREADS / 32];
    __shared__ int s_entry;
    const unsigned int n_huge = status[SW_QCOUNT_HUGE];
    const int total_chunks = (int)status[SW_HUGE_CHUNKS];
    for (int g = blockIdx.x; g < total_chunks; g += gridDim.x) {
        __syncthreads();
        for (unsigned int e = threadIdx.x; e < n_huge; e += FOLD_THREADS) {
            const int off = huge_off[e];
            if (g >= off && g < off + (qhuge[e].len + FOLD_CHUNK - 1) / FOLD_CHUNK) s_entry = (int)e;
        }
        __syncthreads();
        const int e = s_entry;
        const seg_entry s = qhuge[e];
        const int c = g - huge_off[e];
        if (redo_from && (redo_from[e] < 0 || c < redo_from[e] ||
                          redo_from[redo_stride + e] == redo_from[2 * redo_stride + e]))
            continue;   // second pass: only chunks from the hand-over on, and only if there is another binade
        const int nc = s.len - c * FOLD_CHUNK < FOLD_CHUNK ? s.len - c * FOLD_CHUNK : FOLD_CHUNK;
        int eT;
        I m0;
        if (redo_from) {
            eT = redo_from[2 * redo_stride + e];
        } else if (!scan_ready<T>(H[s.node], eT, m0)) {
            if (threadIdx.x == 0) recs[g].ok = 0;
            continue;
        }
        if constexpr (MASKED) {
            const masked_src<T> ms{lev_d, mleaf, mleaf_bytes, mnleaves, s.node, mleaf_level - s.level};
            chunk_summary_cta<T>(ms + (int64_t)c * FOLD_CHUNK, nc, eT, recs + g, s_wtot, s_red, s_abs);
        } else {
            chunk_summary_cta<T>(lev_d + (int64_t)s.level * level_stride + s.start + (int64_t)c * FOLD_CHUNK, nc, eT,
                                 recs + g, s_wtot, s_red, s_abs);
        }
    }
}

// Huge segments, phase B: one CTA per segment stitches the chunk summaries.  First a parallel
// attempt: the chunk maps are composed with a block scan, which gives every chunk's incoming m, and
// all excursion bounds are checked at once (the common case: nothing leaves the binade).  From the
// first chunk that fails the check on, the CTA walks sequentially: safe chunks cost a few integer
// ops, unsafe ones (or any chunk after the exponent changed) are folded exactly.
// Record of chunk c of this segment: recs[rec_base + c * rec_stride].
template <typename T, bool MASKED, bool REDO>
__device__ void huge_apply_body(T *__restrict__ H, const T *__restrict__ lev_d, int64_t level_stride,
                                const seg_entry *__restrict__ qhuge, const int *__restrict__ huge_off,
                                const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs, unsigned int *status,
                                const void *__restrict__ mleaf, int mleaf_bytes, int64_t mnleaves, int mleaf_level,
                                const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs2, int *__restrict__ redo,
                                int redo_stride, T *s_seq /* FOLD_CHUNK elements of shared memory */, const sidx_ref &sx,
                                const slice_ref &sl)
{
    // First pass (REDO false): stitch with the summaries made for the binade eA the node starts in.
    //   redo == nullptr (the masked top fold): chunks that fail the check are folded exactly in place.
    //   redo != nullptr: the first chunk c that fails hands the segment to the second pass -- H[node]
    //   keeps the value in front of chunk c, redo[slot] = c, redo[stride + slot] = eA and
    //   redo[2 * stride + slot] = eB, the binade on the other side of the boundary the chunk ran into.
    // Second pass (REDO true, inside the fold_scan_kernel launch so that its exact folds overlap the
    // long / medium folds): summaries for eB (recs2) next to the old ones (recs); the node may hop
    // between the two binades for free, chunks that straddle the boundary are folded exactly.
    if (status[SW_STATUS]) return;
    if (REDO && !status[SW_REDO_ANY]) return;
    using I = typename fp_bits<T>::I;
    constexpr I LO = I(1) << fp_bits<T>::MB;
    constexpr I HI = I(1) << (fp_bits<T>::MB + 1);
    constexpr int STAGE = FOLD_THREADS / 2;
    __shared__ I s_wtot[FOLD_THREADS / 32][2];
    __shared__ I s_mviol;
    __shared__ fold_smem_ctl ctl;
    __shared__ chunk_rec<I> s_rec[REDO ? 2 : 1][STAGE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_huge = status[SW_QCOUNT_HUGE];
    for (unsigned int it = 0;; ++it) {
        __syncthreads();
        unsigned int slot;
        if constexpr (REDO) {
            slot = blockIdx.x + it * gridDim.x;
        } else {
            if (tid == 0) ctl.slot = atomicAdd(status + SW_QTAKE_HUGE, 1u);
            __syncthreads();
            slot = ctl.slot;
        }
        if (slot >= n_huge) break;
        int c_start = 0;
        int eA = 0, eB = 0;
        if constexpr (REDO) {
            c_start = redo[slot];
            if (c_start < 0) continue;
            eA = redo[redo_stride + slot];
            eB = redo[2 * redo_stride + slot];
        }
        const seg_entry s = qhuge[slot];
        const T *src = lev_d + (int64_t)s.level * level_stride + s.start;
        const int64_t rec_base = MASKED ? (int64_t)(s.node - 1) : (int64_t)huge_off[slot];
        const int64_t rec_stride = MASKED ? (int64_t)(mnleaves - 1) : 1;
        const int nchunks = (s.len + FOLD_CHUNK - 1) / FOLD_CHUNK;
        T cur = H[s.node];
        int e_cur;
        I m;
        bool ready = scan_ready<T>(cur, e_cur, m);
        if constexpr (!REDO) {
            eA = ready ? e_cur : 0;    // summaries were made for this exponent (phase A saw the same value)
            if (redo && tid == 0) redo[slot] = -1;
        }
        if (!REDO && ready) {
            // ---- parallel attempt ---------------------------------------------------------------------
            const int per = (nchunks + FOLD_THREADS - 1) / FOLD_THREADS;
            const int lo = tid * per < nchunks ? tid * per : nchunks;
            const int hi = lo + per < nchunks ? lo + per : nchunks;
            pmap<I> local{0, 0};
            for (int c = lo; c < hi; ++c) {
                const chunk_rec<I> r = recs[rec_base + c * rec_stride];
                local = pmap_then<I>(local, pmap<I>{r.a0, r.a1});
            }
            const pmap<I> incl = pmap_warp_scan<I>(local, lane);
            pmap<I> excl = pmap_shfl_up<I>(incl, 1);
            if (lane == 0) excl = pmap<I>{0, 0};
            if (lane == 31) {
                s_wtot[warp][0] = incl.a0;
                s_wtot[warp][1] = incl.a1;
            }
            if (tid == 0) ctl.first_viol = NO_VIOL;
            __syncthreads();
            pmap<I> pre{0, 0}, tot{0, 0};
#pragma unroll
            for (int w = 0; w < FOLD_THREADS / 32; ++w) {
                const pmap<I> t{s_wtot[w][0], s_wtot[w][1]};
                if (w < warp) pre = pmap_then<I>(pre, t);
                tot = pmap_then<I>(tot, t);
            }
            pre = pmap_then<I>(pre, excl);
            I mm = m + ((m & 1) ? pre.a1 : pre.a0);
            int bad = NO_VIOL;
            I m_bad = 0;
            for (int c = lo; c < hi; ++c) {
                const chunk_rec<I> r = recs[rec_base + c * rec_stride];
                const bool odd = (mm & 1) != 0;
                const I mn = odd ? r.min1 : r.min0, mx = odd ? r.max1 : r.max0;
                if (!(r.ok && 4 * (mm - LO) + mn >= -1 && mm + mx < HI)) {
                    bad = c;
                    m_bad = mm;
                    break;
                }
                mm += odd ? r.a1 : r.a0;
            }
            if (bad != NO_VIOL) atomicMin(&ctl.first_viol, bad);
            __syncthreads();
            const int fb = ctl.first_viol;
            if (fb == NO_VIOL) {
                m += (m & 1) ? tot.a1 : tot.a0;
                c_start = nchunks;
            } else {
                if (bad == fb) s_mviol = m_bad;   // exact m in front of the first chunk that needs care
                __syncthreads();
                m = s_mviol;
                c_start = fb;
            }
        }
        // ---- sequential continuation ---------------------------------------------------------------------
        bool deferred = false;
        int exact_chunks = 0;   // statistics (st_update_stats)
        for (int cb = c_start; cb < nchunks && !deferred; cb += STAGE) {
            __syncthreads();
            if (tid < STAGE && cb + tid < nchunks) {
                s_rec[0][tid] = recs[rec_base + (int64_t)(cb + tid) * rec_stride];
                if constexpr (REDO) s_rec[REDO ? 1 : 0][tid] = recs2[rec_base + (int64_t)(cb + tid) * rec_stride];
            }
            __syncthreads();
            const int lim = nchunks - cb < STAGE ? nchunks - cb : STAGE;
            for (int k = 0; k < lim; ++k) {   // uniform over the CTA
                bool hovering = false;   // small updates, but the chunk touches the binade boundary
                int other = eA;          // the binade on the far side of that boundary
                if (ready) {
                    const int which = e_cur == eA ? 0 : ((REDO && e_cur == eB) ? 1 : -1);
                    if (which >= 0) {
                        const chunk_rec<I> &r = s_rec[REDO ? which : 0][k];
                        const bool odd = (m & 1) != 0;
                        const I mn = odd ? r.min1 : r.min0, mx = odd ? r.max1 : r.max0;
                        if (r.ok && 4 * (m - LO) + mn >= -1 && m + mx < HI) {
                            m += odd ? r.a1 : r.a0;
                            continue;
                        }
                        hovering = r.ok != 0;
                        if (hovering) other = (4 * (m - LO) + mn < -1) ? e_cur - 1 : e_cur + 1;
                    }
                }
                const int c = cb + k;
                if constexpr (!REDO) {
                    if (redo) {
                        // hand over to the second pass, which has the time to fold exactly
                        if (tid == 0) {
                            sidx_store<T>(H, sx, s.node, s.level, ready ? value_of<T>(e_cur, m) : cur);
                            redo[slot] = c;
                            redo[redo_stride + slot] = eA;
                            redo[2 * redo_stride + slot] = other;
                            status[SW_REDO_ANY] = 1u;
                        }
                        deferred = true;
                        break;
                    }
                }
                if (ready) cur = value_of<T>(e_cur, m);
                const int nc = s.len - c * FOLD_CHUNK < FOLD_CHUNK ? s.len - c * FOLD_CHUNK : FOLD_CHUNK;
                if constexpr (MASKED) {
                    const masked_src<T> ms{lev_d, mleaf, mleaf_bytes, mnleaves, s.node, mleaf_level - s.level, sl, 0};
                    cur = hovering ? sequential_fold_cta<T>(cur, ms + (int64_t)c * FOLD_CHUNK, nc, s_seq)
                                   : ordered_fold_cta<T>(cur, ms + (int64_t)c * FOLD_CHUNK, nc, s_wtot, &s_mviol, &ctl, s_seq);
                } else {
                    cur = hovering ? sequential_fold_cta<T>(cur, src + (int64_t)c * FOLD_CHUNK, nc, s_seq)
                                   : ordered_fold_cta<T>(cur, src + (int64_t)c * FOLD_CHUNK, nc, s_wtot, &s_mviol, &ctl, s_seq);
                }
                ready = scan_ready<T>(cur, e_cur, m);
                ++exact_chunks;
            }
        }
        if (!deferred && tid == 0) sidx_store<T>(H, sx, s.node, s.level, ready ? value_of<T>(e_cur, m) : cur);
        if (tid == 0) {
            if (exact_chunks == 0 && !deferred) {
                if (!REDO) atomicAdd(status + SW_STAT_FAST, 1u);
            } else {
                if (!REDO) atomicAdd(status + SW_STAT_CHAINS, 1u);
                atomicAdd(status + SW_STAT_CHAIN_ELEMS, (unsigned int)exact_chunks);
                atomicMax(status + SW_STAT_LONGEST, (unsigned int)exact_chunks);
            }
        }
    }
}

template <typename T, bool MASKED, bool REDO>
__global__ void __launch_bounds__(FOLD_THREADS) fold_huge_apply_kernel(T *__restrict__ H, const T *__restrict__ lev_d,
                                                                       int64_t level_stride,
                                                                       const seg_entry *__restrict__ qhuge,
                                                                       const int *__restrict__ huge_off,
                                                                       const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs,
                                                                       unsigned int *status,
                                                                       const void *__restrict__ mleaf, int mleaf_bytes, int64_t mnleaves,
                                                                       int mleaf_level,
                                                                       const chunk_rec<typename fp_bits<T>::I> *__restrict__ recs2,
                                                                       int *__restrict__ redo, int redo_stride, const sidx_ref sx,
                                                                       const slice_ref sl = slice_ref{})
{
    __shared__ T s_seq[FOLD_CHUNK];
    huge_apply_body<T, MASKED, REDO>(H, lev_d, level_stride, qhuge, huge_off, recs, status, mleaf, mleaf_bytes, mnleaves,
                                     mleaf_level, recs2, redo, redo_stride, s_seq, sx, sl);
}

// Masked variant of phase A for the tiny top tree: records laid out [chunk][node-1] so that a rank's
// chunk range is one contiguous slab (exchanged with one all-gather in the sharded layout).
template <typename T>
__global__ void __launch_bounds__(FOLD_THREADS) fold_top_maps_kernel(const T *__restrict__ H, const T *__restrict__ d,
                                                                     const void *__restrict__ leaf, int leaf_bytes,
                                                                     int64_t nleaves, int leaf_level, int64_t m,
                                                                     int chunk_lo, int chunk_hi,
                                                                     chunk_rec<typename fp_bits<T>::I> *__restrict__ recs,
                                                                     const unsigned int *status)
{
    if (status[SW_STATUS]) return;
    using I = typename fp_bits<T>::I;
    __shared__ I s_wtot[FOLD_THREADS / 32][2];
    __shared__ I s_red[FOLD_THREADS / 32][4];
    __shared__ float s_abs[FOLD_THREADS / 32];
    const int nodes = (int)nleaves - 1;
    const int64_t pairs = (int64_t)(chunk_hi - chunk_lo) * nodes;
    for (int64_t pidx = blockIdx.x; pidx < pairs; pidx += gridDim.x) {
        const int c = chunk_lo + (int)(pidx / nodes);
        const uint32_t node = 1u + (uint32_t)(pidx % nodes);
        chunk_rec<I> *out = recs + (int64_t)c * nodes + (node - 1);
        const int64_t base = (int64_t)c * FOLD_CHUNK;
        int64_t rem = m - base;
        const int nc = rem <= 0 ? 0 : (rem < FOLD_CHUNK ? (int)rem : FOLD_CHUNK);
        int eT;
        I m0;
        __syncthreads();
        if (!scan_ready<T>(H[node], eT, m0)) {
            if (threadIdx.x == 0) out->ok = 0;
            continue;
        }
        const masked_src<T> ms{d, leaf, leaf_bytes, nleaves, node, leaf_level - heap_level(node)};
        chunk_summary_cta<T>(ms + base, nc, eT, out, s_wtot, s_red, s_abs);
    }
}

// -----------------------------------------------------------------------------------------
// host driver
// -----------------------------------------------------------------------------------------
// group size (in key bits) of the float leaf pass / deep pass: the largest gb with m * 2^gb <= n, at most 10,
// and never the whole tree (a group's node stays below the root)
static int deep_group_bits(int m, int64_t n, int depth)
{
    int gb = 0;
    while (gb < 10 && gb + 2 <= depth && ((int64_t)m << (gb + 1)) <= n) ++gb;
    return gb;
}

// ST_TIMELINE=1 (debugging aid): one stderr line per large float batch with the completion times (us after the
// start of the batch) of every stage on the caller's stream and on the side stream; synchronises after every batch.
struct update_timeline {
    bool on = false;
    cudaEvent_t m[10] = {}, sd[3] = {};
    update_timeline()
    {
        on = getenv("ST_TIMELINE") != nullptr;
        if (on) {
            for (auto &e : m) cudaEventCreate(&e);
            for (auto &e : sd) cudaEventCreate(&e);
        }
    }
    void main(int k, cudaStream_t s) { if (on) cudaEventRecord(m[k], s); }
    void side(int k, cudaStream_t s) { if (on) cudaEventRecord(sd[k], s); }
    void report()
    {
        if (!on) return;
        cudaEventSynchronize(m[9]);
        static const char *mn[10] = {"start", "keys", "sort", "leaf", "partition", "maps", "stitch+maps2", "fold_scan", "", "join"};
        float v;
        fprintf(stderr, "[timeline us] main:");
        for (int k = 1; k < 10; ++k)
            if (mn[k][0] && cudaEventElapsedTime(&v, m[0], m[k]) == cudaSuccess) fprintf(stderr, " %s=%.0f", mn[k], v * 1e3f);
        fprintf(stderr, " | side:");
        if (cudaEventElapsedTime(&v, m[0], sd[0]) == cudaSuccess) fprintf(stderr, " deep=%.0f", v * 1e3f);
        if (cudaEventElapsedTime(&v, m[0], sd[1]) == cudaSuccess) fprintf(stderr, " retire=%.0f", v * 1e3f);
        fprintf(stderr, "\n");
    }
};

// hybrid partition (st_partition_hybrid.cuh): top levels by counting sort, the rest per segment in shared memory; the
// cooperative kernel then only runs (from level LT) for batches too skewed for that.  ST_HYBRID=0: A/B switch.
static bool hy_applies(int m, int depth)
{
    static const bool hybrid_off = getenv("ST_HYBRID") && atoi(getenv("ST_HYBRID")) == 0;
    return !hybrid_off && depth >= hy_top_levels(m) + 2 && m > RETIRE_SEG;
}

template <typename T>
static hybrid_params<T> hy_params(const st_tree *t, const update_ws &w, T *lev, int m, int64_t B)
{
    const int LT = hy_top_levels(m);
    hybrid_params<T> hp;
    hp.n = (uint32_t)t->n; hp.depth = t->depth; hp.m = m; hp.ntiles = (m + HY_TILE - 1) / HY_TILE; hp.LT = LT;
    hp.key0 = w.kj; hp.lev_key = w.lev_key; hp.lev_d = lev; hp.lev_stride = B;
    hp.hist = w.hy_hist; hp.bintotal = w.hy_bintotal;
    hp.segstart = w.hy_segstart; hp.seglen = w.hy_seglen; hp.segstate = w.hy_segstate;
    hp.sC = (LT & 1) ? w.seg_sa : w.seg_sb;
    hp.eC = (LT & 1) ? w.seg_ea : w.seg_eb;
    hp.qlong = w.queue; hp.qmed = w.queue_med; hp.qhuge = w.queue_huge; hp.qret = w.queue_ret;
    hp.huge_off = w.huge_off; hp.status = t->d_status;
    return hp;
}

template <typename T>
static cudaError_t update_chunk(st_tree *t, const update_ws &w, const int64_t *idx, int m, const T *vals,
                                int64_t nv, int64_t j0, int64_t B, T *delta_out, const T *deltas_in,
                                bool small, bool assign, cudaStream_t s, int phase = 0)
{
    // phase 0: the whole update.  phase 1: up to and including the leaf pass (the differences d_j exist and
    // delta_out is filled); phase 2: the rest (ancestors), for the batch phase 1 left in the workspace.
    // deltas_in != nullptr: "apply deltas" mode (sharded top tree): the per-update
    // differences are given, the leaves are not touched, only proper ancestors are folded.
    const uint32_t n = (uint32_t)t->n;
    const int depth = t->depth;
    T *H = (T *)t->heap;
    unsigned int *status = t->d_status;
    const unsigned g = (unsigned)((m + 255) / 256);

    auto mark = [&](int k) { if (t->profile) cudaEventRecord(t->prof_ev[k], s); };
    static thread_local update_timeline tl;
    if constexpr (is_float_t<T>::value) {
        if (small && phase == 2) return cudaSuccess;
        if (small) {
            // one CTA does validation, sort, leaf pass and the whole ordered partition (st_update_small.cuh)
            small_params<T> sp;
            sp.H = H; sp.n = n; sp.depth = depth; sp.m = m; sp.idx = idx; sp.vals = vals;
            sp.nv = (uint64_t)nv; sp.j0 = (uint64_t)j0; sp.deltas_in = deltas_in; sp.delta_out = delta_out;
            sp.lev_d = (T *)w.lev_d; sp.lev_stride = B; sp.qlong = w.queue; sp.qmed = w.queue_med; sp.status = status;
            sp.qret = w.queue_ret; sp.key0_out = w.kj; sp.lev_key = w.lev_key; sp.assign = assign;
            g_kernel_launches += 2;
            if (m <= SMALL_THREADS) {
                static per_device_once attr1;   // function attributes are per device
                cudaError_t ea = attr1.run(t->device, [&] {
                    return cudaFuncSetAttribute(update_small_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)small_smem_bytes<T, 1>());
                });
                if (ea != cudaSuccess) return ea;
                update_small_kernel<T, 1><<<1, SMALL_THREADS, small_smem_bytes<T, 1>(), s>>>(sp);
            } else {
                static per_device_once attr4;
                cudaError_t ea = attr4.run(t->device, [&] {
                    return cudaFuncSetAttribute(update_small_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)small_smem_bytes<T, 4>());
                });
                if (ea != cudaSuccess) return ea;
                update_small_kernel<T, 4><<<1, SMALL_THREADS, small_smem_bytes<T, 4>(), s>>>(sp);
            }
            retire_params<T> rp{w.queue_ret, w.kj, w.lev_key, depth, n};
            fold_scan_kernel<T><<<t->sm_count, FOLD_THREADS, 0, s>>>(H, (const T *)w.lev_d, B, w.queue, w.queue_med,
                                                                    status, rp, huge_params<T>{nullptr, nullptr, nullptr, nullptr, nullptr, 0}, sidx_of(t));
            return cudaGetLastError();
        }
    }
    size_t cb = w.cub_bytes;
    cudaError_t e = cudaSuccess;
    const T *d_batch = deltas_in;           // d in batch order
    if (phase == 2) d_batch = (const T *)w.dj;
    if (phase != 2) {
    mark(0);
    tl.main(0, s);
    ++g_kernel_launches;
    make_keys_kernel<<<g, 256, 0, s>>>(idx, m, n, depth, w.keys_a, w.kj, w.j_a, status);
    tl.main(1, s);
    if constexpr (is_float_t<T>::value) {
        // The histogram and the scan of the hybrid partition need nothing but the batch-ordered keys: they run on the
        // tree's side stream beside the sort and the leaf pass (ST_HY_AHEAD=0: behind them, on the caller's stream).
        t->hy_ahead = 0;
        static const bool overlap_off0 = getenv("ST_OVERLAP") && atoi(getenv("ST_OVERLAP")) == 0;
        static const bool hy_ahead_off = getenv("ST_HY_AHEAD") && atoi(getenv("ST_HY_AHEAD")) == 0;
        if (!deltas_in && !overlap_off0 && !hy_ahead_off && !t->profile && hy_applies(m, depth)) {
            if (!t->side) {
                e = cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking);
                if (e != cudaSuccess) return e;
            }
            e = cudaEventRecord(t->ev_fork, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(t->side, t->ev_fork, 0);
            if (e != cudaSuccess) return e;
            const hybrid_params<T> hp = hy_params<T>(t, w, (T *)w.lev_d, m, B);
            g_kernel_launches += 2;
            hy_hist_kernel<<<hp.ntiles, HY_THREADS, 0, t->side>>>(w.kj, m, depth, hp.LT, w.hy_hist, status);
            hy_scan_kernel<T><<<1 << hp.LT, HY_THREADS, 0, t->side>>>(hp);
            e = cudaEventRecord(t->ev_hy, t->side);
            if (e != cudaSuccess) return e;
            t->hy_ahead = 1;
        }
    }
    }
    if (!deltas_in && phase != 2) {
        // floats only need the duplicates of a leaf adjacent and batch-ordered: the low gb key bits stay unsorted,
        // gb as large as keeps the groups this creates (2^gb neighbouring leaves) at <= 1 update on average
        int gb = 0;
        if constexpr (is_float_t<T>::value) gb = deep_group_bits(m, (int64_t)n, depth);
        t->pending_gb = gb;
        e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.keys_a, w.keys_b, w.j_a, w.j_b, m, gb, depth + 1, s);
        if (e != cudaSuccess) return e;
        mark(1);
        tl.main(2, s);
        ++g_kernel_launches;
        if (gb) {
            e = cudaMemsetAsync(w.bigmap, 0, (((size_t)1 << (depth - gb)) + 7) / 8 + 4, s);
            if (e != cudaSuccess) return e;
            leaf_pass_grouped_kernel<T><<<g, 256, 0, s>>>(H, n, w.keys_b, w.j_b, m, vals, (uint64_t)nv, (uint64_t)j0,
                                                         (T *)w.dj, (T *)w.ds, gb, depth, w.bigmap, status, assign);
        } else {
            leaf_pass_kernel<T><<<g, 256, 0, s>>>(H, n, w.keys_b, w.j_b, m, vals, (uint64_t)nv, (uint64_t)j0,
                                                  (T *)w.dj, (T *)w.ds, status, assign);
        }
        d_batch = (const T *)w.dj;
        if (delta_out) {
            e = cudaMemcpyAsync(delta_out, w.dj, (size_t)m * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return e;
        }
    }
    if (phase == 1) return cudaGetLastError();

    if constexpr (!is_float_t<T>::value) {
        ++g_kernel_launches;
        if (deltas_in)   // unsorted keys: adjacent equal ancestors still merge, the rest use separate atomics
            int_propagate_kernel<T><<<g, 256, 0, s>>>(H, n, depth, w.kj, deltas_in, m, status, sidx_of(t));
        else
            int_propagate_kernel<T><<<g, 256, 0, s>>>(H, n, depth, w.keys_b, (const T *)w.ds, m, status, sidx_of(t));
        return cudaGetLastError();
    } else {
        // Level 0 order = batch order: (kj, dj).  Level l+1 order = stable sort of level l by
        // the top l+1 key bits.  lev_d[l] keeps every level's ordered d for the long folds.
        T *lev = (T *)w.lev_d;
        mark(2);
        tl.main(3, s);
        // Side stream.  Three groups of nodes are written by three independent parts of the pipeline, all of which
        // only read the batch arrays: (1) the deep levels below the leaf groups -- deep_red_kernel, in leaf-sorted
        // order, needs nothing but the leaf pass; (2) the nodes of retired segments -- retire_kernel, needs the
        // partition; (3) the nodes of longer segments -- the folds.  (1) and (2) are bound by the L2 / HBM atomic
        // units, the partition and (3) by shared memory and integer math, so (1) + (2) run on the tree's side stream
        // (retirement with half of every SM's thread slots) beside the partition and the folds on the caller's.
        static const bool overlap_off = getenv("ST_OVERLAP") && atoi(getenv("ST_OVERLAP")) == 0;   // A/B switch
        if (!overlap_off && !t->profile && !t->side) {   // created by the first large float batch (the caller holds the tree's lock)
            e = cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        const bool overlap = !overlap_off && !t->profile && t->side;
        const int gbd = deltas_in ? 0 : t->pending_gb;            // leaf-group bits (0: no deep pass)
        const int deep_level = gbd ? depth - gbd : (1 << 20);     // retirement stops here, except under flagged groups
        cudaStream_t rs = overlap ? t->side : s;
        const sidx_ref sx = sidx_of(t);
        // ST_DEEP_EARLY=1 starts deep_red_kernel beside the PARTITION instead of behind it.  Measured: the partition
        // then takes 212 us instead of 118 (its shared-memory kernels stall behind the atomics' L2 misses), so the
        // pairing is: partition alone, then deep pass + retirement (atomics) beside the folds (integer math).
        static const bool deep_early = getenv("ST_DEEP_EARLY") && atoi(getenv("ST_DEEP_EARLY")) != 0;
        auto launch_deep = [&](cudaStream_t ds, unsigned grid) {
            if (!gbd) return;
            ++g_kernel_launches;
            if (sx.base) deep_red_kernel<T, true><<<grid, 256, 0, ds>>>(H, n, depth, gbd, w.keys_b, (const T *)w.ds, m, status, sx);
            else deep_red_kernel<T, false><<<grid, 256, 0, ds>>>(H, n, depth, gbd, w.keys_b, (const T *)w.ds, m, status, sx);
            tl.side(0, ds);
        };
        static const int deep_ctas = getenv("ST_DEEP_CTAS") ? atoi(getenv("ST_DEEP_CTAS")) : 4;   // per SM (tuning)
        if (overlap && deep_early) {
            e = cudaEventRecord(t->ev_fork, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(t->side, t->ev_fork, 0);
            if (e != cudaSuccess) return e;
            launch_deep(rs, t->sm_count * 2);
        }
        e = cudaMemcpyAsync(lev, d_batch, (size_t)m * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        if (!t->coop_ok) return cudaErrorCooperativeLaunchTooLarge;
        part_params<T> pp;
        pp.H = H; pp.n = n; pp.depth = depth; pp.m = m;
        pp.ntiles = (m + PART_TILE - 1) / PART_TILE;
        pp.key0 = w.kj; pp.lev_key = w.lev_key;
        pp.lev_d = lev; pp.lev_stride = B;
        pp.sA = w.seg_sa; pp.eA = w.seg_ea; pp.sB = w.seg_sb; pp.eB = w.seg_eb;
        pp.Z = w.zcount; pp.tilesum = w.tilesum;
        pp.qlong = w.queue; pp.qmed = w.queue_med; pp.qhuge = w.queue_huge; pp.huge_off = w.huge_off;
        pp.qret = w.queue_ret;
        pp.status = status;
        pp.l0 = 0;
        // hybrid partition: top levels by counting sort, the rest per segment in shared memory; the cooperative
        // kernel below then only runs (from level LT) for batches too skewed for that
        if (hy_applies(m, depth)) {
            const hybrid_params<T> hp = hy_params<T>(t, w, lev, m, B);
            const size_t seg_smem = hy_segment_smem<T>();
            static per_device_once hy_attr;   // function attributes are per device
            e = hy_attr.run(t->device, [&] {
                return cudaFuncSetAttribute(hy_segment_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem);
            });
            if (e != cudaSuccess) return e;
            if (t->hy_ahead) {       // histogram + scan ran ahead on the side stream (first half of this batch)
                e = cudaStreamWaitEvent(s, t->ev_hy, 0);
                if (e != cudaSuccess) return e;
                t->hy_ahead = 0;
            } else {
                g_kernel_launches += 2;
                hy_hist_kernel<<<hp.ntiles, HY_THREADS, 0, s>>>(w.kj, m, depth, hp.LT, w.hy_hist, status);
                hy_scan_kernel<T><<<1 << hp.LT, HY_THREADS, 0, s>>>(hp);
            }
            g_kernel_launches += 2;
            hy_scatter_kernel<T><<<hp.ntiles, HY_THREADS, 0, s>>>(hp);
            hy_segment_kernel<T><<<1 << hp.LT, HY_B_THREADS, seg_smem, s>>>(hp);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            pp.l0 = hp.LT;
        }
        if (t->part_occ < 1) {   // per tree (one dtype, one device); the caller holds the tree's lock
            int occ = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, partition_levels_kernel<T>, PART_THREADS, 0);
            if (e != cudaSuccess) return e;
            if (occ < 1) return cudaErrorCooperativeLaunchTooLarge;
            t->part_occ = occ;
        }
        const int occ = t->part_occ;
        // balanced persistent grid: every CTA gets the same number of tiles (the grid barriers wait
        // for the slowest CTA)
        const int max_res = occ * t->sm_count;
        const int per_cta = (pp.ntiles + max_res - 1) / max_res;
        const int grid = (pp.ntiles + per_cta - 1) / per_cta;
        void *args[] = {&pp};
        ++g_kernel_launches;
        e = cudaLaunchCooperativeKernel((const void *)partition_levels_kernel<T>, dim3(grid), dim3(PART_THREADS),
                                        args, 0, s);
        if (e != cudaSuccess) return e;
        using REC = chunk_rec<typename fp_bits<T>::I>;
        mark(3);
        tl.main(4, s);
        ++g_kernel_launches;
        // (retirement: see the note on the side stream above the partition)
        if (overlap) {   // the side stream (already busy with deep_red_kernel) goes on once the partition is complete
            e = cudaEventRecord(t->ev_fork, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(t->side, t->ev_fork, 0);
            if (e != cudaSuccess) return e;
        }
        // Order on the side stream: retirement (L2-resident atomics) first, beside the huge-segment summaries and the
        // stitch; the deep pass (344 MB of scattered DRAM read-modify-write) behind it, beside the scan-bound folds.
        // The other way round (ST_SIDE_ORDER=0) the summaries took 108 us beside the deep pass instead of 50: 572 vs 541 us.
        static const bool retire_first = !(getenv("ST_SIDE_ORDER") && atoi(getenv("ST_SIDE_ORDER")) == 0);
        if (!(overlap && deep_early) && !retire_first) launch_deep(rs, overlap ? t->sm_count * deep_ctas : g);
        static const int retire_ctas = getenv("ST_RETIRE_CTAS") ? atoi(getenv("ST_RETIRE_CTAS")) : 4;   // per SM beside the folds (tuning)
        {
            const unsigned rg = t->sm_count * (overlap ? retire_ctas : 8);
            if (sx.base)
                retire_kernel<T, true><<<rg, 256, 0, rs>>>(H, n, depth, w.kj, w.lev_key, lev, B, w.queue_ret, status, deep_level, gbd,
                                                           w.bigmap, sx);
            else
                retire_kernel<T, false><<<rg, 256, 0, rs>>>(H, n, depth, w.kj, w.lev_key, lev, B, w.queue_ret, status, deep_level, gbd,
                                                            w.bigmap, sx);
        }
        tl.side(1, rs);
        if (!(overlap && deep_early) && retire_first) launch_deep(rs, overlap ? t->sm_count * deep_ctas : g);
        if (overlap) {
            e = cudaEventRecord(t->ev_join, t->side);
            if (e != cudaSuccess) return e;
        }
        g_kernel_launches += 4;
        mark(4);
        // huge segments: chunk summaries -> first-pass stitch (pure map composition; a segment whose
        // node runs into a binade boundary is handed over) -> summaries for the far side of that
        // boundary (usually nothing to do) -> second-pass stitch inside the fold launch
        const int redo_stride = (int)((t->depth + 1) * (B / (HUGE_SEG + 1) + 2));
        fold_huge_maps_kernel<T, false><<<t->sm_count * 8, FOLD_THREADS, 0, s>>>(
            H, lev, B, w.queue_huge, w.huge_off, (REC *)w.huge_recs, status, nullptr, 0, 0, 0, nullptr, 0);
        mark(5);
        tl.main(5, s);
        fold_huge_apply_kernel<T, false, false><<<t->sm_count, FOLD_THREADS, 0, s>>>(
            H, lev, B, w.queue_huge, w.huge_off, (const REC *)w.huge_recs, status, nullptr, 0, 0, 0, nullptr, w.huge_redo,
            redo_stride, sidx_of(t));
        fold_huge_maps_kernel<T, false><<<t->sm_count * 8, FOLD_THREADS, 0, s>>>(
            H, lev, B, w.queue_huge, w.huge_off, (REC *)w.huge_recs2, status, nullptr, 0, 0, 0, w.huge_redo, redo_stride);
        mark(6);
        tl.main(6, s);
        fold_scan_kernel<T><<<t->sm_count * 8, FOLD_THREADS, 0, s>>>(
            H, lev, B, w.queue, w.queue_med, status, retire_params<T>{nullptr, nullptr, nullptr, 0, 0},
            huge_params<T>{w.queue_huge, w.huge_off, (const REC *)w.huge_recs, (const REC *)w.huge_recs2, w.huge_redo,
                           redo_stride},
            sidx_of(t));
        tl.main(7, s);
        if (overlap) {
            e = cudaStreamWaitEvent(s, t->ev_join, 0);
            if (e != cudaSuccess) return e;
        }
        tl.main(9, s);
        tl.report();
        mark(7);
        if (t->profile) {
            // profiling mode is synchronous by design (bench.py's roofline leg only); its phases run one after another
            cudaEventSynchronize(t->prof_ev[7]);
            for (int k = 0; k < 7; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, t->prof_ev[k], t->prof_ev[k + 1]);
                t->prof_ms[k] += ms;
            }
            t->prof_ms[7] += 1.0;
        }
        return cudaGetLastError();
    }
}

// Tiny power-of-two tree (the replicated top tree of the sharded layout): fold ALL m differences, in
// batch order, into every internal node that is an ancestor of their leaf.  Every node is one
// "huge segment" over the whole vector with non-members masked to 0 -- no sort, no partition.
// Phase A (chunk summaries, any chunk range -> can be split over ranks) and phase B (stitch) are
// separate entry points; records are [chunk][node-1], 64 bytes apart at most.
__global__ void top_entries_kernel(seg_entry *q, unsigned int *status, int n, int m)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;   // node 1 .. n-1
    if (h == 0) {
        status[SW_QCOUNT_HUGE] = (unsigned int)(n - 1);
        status[SW_QTAKE_HUGE] = 0;
    }
    if (h >= 1 && h < n) q[h - 1] = seg_entry{heap_level((uint32_t)h), 0, m, (uint32_t)h};
}

static bool fold_by_leaf_ok(const st_tree *t, int64_t m)
{
    return (t->dtype == ST_F32 || t->dtype == ST_F64) && t->n <= 64 && (t->n & (t->n - 1)) == 0 && m <= 0x7fffffffLL;
}

size_t fold_rec_bytes(const st_tree *t) { return t->dtype == ST_F64 ? sizeof(chunk_rec<int64_t>) : sizeof(chunk_rec<int32_t>); }

cudaError_t launch_fold_by_leaf_maps(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                     int chunk_lo, int chunk_hi, void *recs, cudaStream_t s)
{
    if (!fold_by_leaf_ok(t, m) || (leaf_bytes != 1 && leaf_bytes != 8)) return cudaErrorInvalidValue;
    if (chunk_hi <= chunk_lo) return cudaSuccess;
    const int64_t pairs = (int64_t)(chunk_hi - chunk_lo) * (t->n - 1);
    const unsigned g = (unsigned)(pairs < (int64_t)t->sm_count * 8 ? pairs : (int64_t)t->sm_count * 8);
    ++g_kernel_launches;
    if (t->dtype == ST_F32)
        fold_top_maps_kernel<float><<<g, FOLD_THREADS, 0, s>>>((const float *)t->heap, (const float *)deltas, leaf,
                                                               leaf_bytes, t->n, t->depth, m, chunk_lo, chunk_hi,
                                                               (chunk_rec<int32_t> *)recs, t->d_status);
    else
        fold_top_maps_kernel<double><<<g, FOLD_THREADS, 0, s>>>((const double *)t->heap, (const double *)deltas, leaf,
                                                                leaf_bytes, t->n, t->depth, m, chunk_lo, chunk_hi,
                                                                (chunk_rec<int64_t> *)recs, t->d_status);
    return cudaGetLastError();
}

cudaError_t launch_fold_by_leaf_apply(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                      const void *recs, cudaStream_t s, const void *const *slice_d,
                                      const void *const *slice_leaf, int64_t slice_span, int64_t slice_m)
{
    if (!fold_by_leaf_ok(t, m) || (leaf_bytes != 1 && leaf_bytes != 8)) return cudaErrorInvalidValue;
    if (m <= 0) return cudaSuccess;
    const size_t need = 256 + (size_t)t->n * sizeof(seg_entry);
    if (need > t->ws.bytes) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        if (t->ws.base) cudaFree(t->ws.base);
        t->ws.base = nullptr; t->ws.bytes = 0; t->reserved_batch = 0; t->cub_cap = -1;
        e = cudaMalloc(&t->ws.base, need < 4096 ? 4096 : need);
        if (e != cudaSuccess) return e;
        t->ws.bytes = need < 4096 ? 4096 : need;
    }
    seg_entry *q = (seg_entry *)t->ws.base;
    slice_ref sl{};
    if (slice_d) { sl.d = slice_d; sl.leaf = slice_leaf; sl.span = slice_span; sl.m = slice_m; }
    g_kernel_launches += 2;
    top_entries_kernel<<<1, 64, 0, s>>>(q, t->d_status, (int)t->n, (int)m);
    if (t->dtype == ST_F32)
        fold_huge_apply_kernel<float, true, false><<<(unsigned)(t->n - 1), FOLD_THREADS, 0, s>>>(
            (float *)t->heap, (const float *)deltas, 0, q, nullptr, (const chunk_rec<int32_t> *)recs, t->d_status, leaf,
            leaf_bytes, t->n, t->depth, nullptr, nullptr, 0, sidx_ref{}, sl);
    else
        fold_huge_apply_kernel<double, true, false><<<(unsigned)(t->n - 1), FOLD_THREADS, 0, s>>>(
            (double *)t->heap, (const double *)deltas, 0, q, nullptr, (const chunk_rec<int64_t> *)recs, t->d_status, leaf,
            leaf_bytes, t->n, t->depth, nullptr, nullptr, 0, sidx_ref{}, sl);
    return cudaGetLastError();
}

// both phases on this device (single-process use); the records live in the staging area
cudaError_t launch_fold_by_leaf(st_tree *t, const int64_t *leaf, const void *deltas, int64_t m, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    if (!fold_by_leaf_ok(t, m)) return cudaErrorInvalidValue;
    const int nchunks = (int)((m + FOLD_CHUNK - 1) / FOLD_CHUNK);
    const size_t need = (size_t)nchunks * (t->n - 1) * fold_rec_bytes(t);
    if (need > t->scratch.bytes) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        if (t->scratch.base) cudaFree(t->scratch.base);
        t->scratch.base = nullptr; t->scratch.bytes = 0;
        e = cudaMalloc(&t->scratch.base, need);
        if (e != cudaSuccess) return e;
        t->scratch.bytes = need;
    }
    cudaError_t e = launch_fold_by_leaf_maps(t, leaf, 8, deltas, m, 0, nchunks, t->scratch.base, s);
    if (e != cudaSuccess) return e;
    return launch_fold_by_leaf_apply(t, leaf, 8, deltas, m, t->scratch.base, s);
}

cudaError_t launch_update(st_tree *t, const int64_t *idx, int64_t ni, const void *vals, int64_t nv,
                          void *delta_out, const void *deltas_in, cudaStream_t s, bool assign, int phase)
{
    if (phase == 2) {
        // second half of a split update (st_update_begin_device / st_update_finish_device)
        if (!t->pending_split) return cudaSuccess;
        t->pending_split = 0;
        const int64_t cap2 = t->reserved_batch;
        const int m2 = (int)t->pending_ni;
        ST_DISPATCH(t->dtype, {
            update_ws w = carve(t, cap2, (char *)t->ws.base);
            return update_chunk<T>(t, w, nullptr, m2, nullptr, 0, 0, cap2, nullptr, nullptr, false, false, s, 2);
        });
        return cudaSuccess;
    }
    t->pending_split = 0;
    if (ni <= 0 && nv <= 0) return cudaSuccess;
    if (ni <= 0) {
        // nothing to apply, but the reference still scans ALL vals for negatives (pyx:37-39)
        cudaError_t eb = launch_batch_begin(t, s);
        if (eb != cudaSuccess) return eb;
        ST_DISPATCH(t->dtype, {
            ++g_kernel_launches;
            validate_kernel<T><<<grid_for(nv, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>(
                idx, 0, (const T *)vals, nv, t->n, t->d_status);
        });
        return cudaGetLastError();
    }
    // chunk capacity: the (grow-only) reserved batch size, capped at UPDATE_MAX_CHUNK.  Larger
    // batches are applied chunk after chunk, which is exact because the reference itself is
    // sequential in the batch position.
    int64_t cap = ni > t->reserved_batch ? ni : t->reserved_batch;
    if (cap > UPDATE_MAX_CHUNK) cap = UPDATE_MAX_CHUNK;
    const size_t need = carve(t, cap, nullptr).total;
    if (need > t->ws.bytes) {
        // never happens once st_reserve() covered the batch size (e.g. before graph capture)
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        if (t->ws.base) cudaFree(t->ws.base);
        t->ws.base = nullptr;
        t->ws.bytes = 0;
        e = cudaMalloc(&t->ws.base, need);
        if (e != cudaSuccess) return e;
        t->ws.bytes = need;
    }
    t->reserved_batch = cap;
    t->stat_updates += ni;
    ST_DISPATCH(t->dtype, {
        const bool small = is_float_t<T>::value && !t->profile && ni <= SMALL_M && (deltas_in || nv <= SMALL_M);
        if (!small) {
            cudaError_t eb = launch_batch_begin(t, s);
            if (eb != cudaSuccess) return eb;
            ++g_kernel_launches;
            validate_kernel<T><<<grid_for(ni > nv ? ni : nv, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>(
                idx, ni, (const T *)vals, deltas_in ? 0 : nv, t->n, t->d_status);
        }
        update_ws w = carve(t, cap, (char *)t->ws.base);
        // a split update (phase 1) covers a batch that is one chunk of the multi-kernel pipeline; anything else
        // is applied completely now and the second half has nothing left to do
        const bool split = phase == 1 && !small && ni <= cap && !t->profile;
        for (int64_t j0 = 0; j0 < ni; j0 += cap) {
            const int m = (int)(ni - j0 < cap ? ni - j0 : cap);
            cudaError_t e = update_chunk<T>(t, w, idx + j0, m, (const T *)vals, nv, j0, cap,
                                            delta_out ? (T *)delta_out + j0 : nullptr,
                                            deltas_in ? (const T *)deltas_in + j0 : nullptr, small, assign, s,
                                            split ? 1 : 0);
            if (e != cudaSuccess) return e;
        }
        if (split) {
            t->pending_split = 1;
            t->pending_ni = ni;
        }
    });
    return cudaGetLastError();
}

}  // namespace st

#ifdef ST_SMALL_TIMING
extern "C" __attribute__((visibility("default"))) int st_debug_ticks(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, st::g_part_ticks, 256 * sizeof(unsigned long long));
}
#endif
