// st_peer.cu -- peer-memory exchange area of the subtree-sharded tree.
//
// The reference is single-process (no collective anywhere), so nothing is replaced
// here; what is pinned is the RESULT: the sharded tree stays bit-identical to one reference tree.
// This file is only the transport: every rank of one node owns one device buffer
//
//     [ flag page: flags[channel][source rank] (uint32) | payload ... ]
//
// whose CUDA IPC handle is exchanged once (through torch.distributed, by the Python layer); after
// st_peer_connect every rank holds a device-visible pointer to every other rank's buffer (NVLink /
// NVSwitch peer access).  Kernels of rank r then store their results STRAIGHT into the buffers of
// all ranks at the batch-order position of every element, so no all-gather, no slab layout and no
// "back to batch order" pass is needed:
//
//   scatter    payload_q[off + pos[k]] = src[k]  for every rank q   (update: the differences d_j)
//   broadcast  payload_q[off .. off+bytes) = payload_r[off ..)      (top-fold records, shard roots)
//   search     st_search_push_device: the local descent writes the GLOBAL leaf index of query k to
//              payload_q[off + 8 * pos[k]] on every rank                (sample)
//   signal / wait   a release store of a step counter into flags[channel][r] of every rank, and a
//              spin (acquire loads, bounded by a timeout) until all sources reached the step.
// All of it is stream-ordered; nothing synchronises with the host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "st_common.cuh"

namespace st {

constexpr int PEER_MAX = 64;
constexpr int PEER_CHANNELS = 16;
constexpr size_t PEER_FLAG_BYTES = 8192;   // 16 channels x 64 sources x 4 B, rounded up

struct peer_ptrs {
    char *base[PEER_MAX];   // start of every rank's buffer (flag page first)
    int n;
};

}  // namespace st

struct st_peer {
    int device = 0, rank = 0, world = 1;
    size_t payload_bytes = 0;
    char *local = nullptr;
    bool opened[st::PEER_MAX] = {};
    st::peer_ptrs ptrs{};
    unsigned int *d_err = nullptr;   // bit 0: a wait timed out
    unsigned int *h_err = nullptr;
    unsigned int *h_mat = nullptr;   // pinned: host copy of the slices count matrix (st_slices_matrix_fetch / _wait)
    cudaEvent_t mat_ev = nullptr;
    const void **d_tab = nullptr;    // device: 2 parities x 2 arrays (differences, shard ids) x PEER_MAX slice pointers
};

namespace st {

__device__ __forceinline__ unsigned long long gtimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void peer_signal_kernel(peer_ptrs pp, int rank, int channel, unsigned int value)
{
    const int q = threadIdx.x;
    if (q >= pp.n) return;
    __threadfence_system();   // everything this stream stored before (earlier kernels) is ordered in front of the flag
    unsigned int *f = reinterpret_cast<unsigned int *>(pp.base[q]) + channel * PEER_MAX + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
}

__global__ void peer_wait_kernel(const unsigned int *flags, int channel, unsigned int value, int world,
                                 unsigned long long timeout_ns, unsigned int *err)
{
    const int q = threadIdx.x;
    if (q >= world) return;
    const unsigned int *f = flags + channel * PEER_MAX + q;
    const unsigned long long t0 = gtimer_ns();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - value) >= 0) break;
        if (gtimer_ns() - t0 > timeout_ns) {   // a rank died or never launched: do not hang the GPU
            atomicOr(err, 1u);
            break;
        }
        __nanosleep(200);
    }
}

// payload_q[off .. off + bytes) = local payload[off ..) for every q != rank (V = widest vector the range allows)
template <typename V>
__global__ void __launch_bounds__(256) peer_broadcast_kernel(peer_ptrs pp, int rank, size_t off, size_t bytes)
{
    const size_t nvec = bytes / sizeof(V);
    const V *src = reinterpret_cast<const V *>(pp.base[rank] + off);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const V v = src[i];
        for (int q = 0; q < pp.n; ++q)
            if (q != rank) reinterpret_cast<V *>(pp.base[q] + off)[i] = v;
    }
}

// payload_q[off + pos[k] * sizeof(T)] = src[k] for EVERY q (the local copy included)
template <typename T>
__global__ void __launch_bounds__(256) peer_scatter_kernel(peer_ptrs pp, size_t off, const T *__restrict__ src,
                                                           const int32_t *__restrict__ pos, int64_t count,
                                                           const unsigned int *__restrict__ count_dev)
{
    if (count_dev) {
        const int64_t c = (int64_t)*count_dev;
        if (c < count) count = c;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const T v = src[k];
        const size_t at = off + (size_t)pos[k] * sizeof(T);
        for (int q = 0; q < pp.n; ++q) *reinterpret_cast<T *>(pp.base[q] + at) = v;
    }
}

// payload_q[off + k * sizeof(T)] = src[k], k < count, for EVERY rank q: consecutive lanes store consecutive
// elements, so every warp store is one full 128-byte (or 256-byte) write per peer -- the scattered 4 / 8-byte
// stores of peer_scatter_kernel cost a whole NVLink packet each.  The element count goes to a header word.
template <typename T>
__global__ void __launch_bounds__(256) peer_put_kernel(peer_ptrs pp, size_t off, const T *__restrict__ src, int64_t count,
                                                       const unsigned int *__restrict__ count_dev, int rank, int64_t count_off)
{
    if (count_dev) {
        const int64_t c = (int64_t)*count_dev;
        if (c < count) count = c;
    }
    if (count_off >= 0 && blockIdx.x == 0 && threadIdx.x < pp.n)
        *reinterpret_cast<unsigned int *>(pp.base[threadIdx.x] + count_off) = (unsigned int)count;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const T v = src[k];
        for (int q = 0; q < pp.n; ++q) reinterpret_cast<T *>(pp.base[q] + off)[k] = v;
    }
    (void)rank;
}

// out[pos] = q * leaves_per_shard + idx for the (pos, idx) pairs every rank q has put into this rank's buffer
__global__ void __launch_bounds__(256) peer_scatter_pairs_kernel(const char *__restrict__ local, size_t pairs_off,
                                                                 size_t slab_bytes, size_t counts_off, int world,
                                                                 int64_t leaves_per_shard, int64_t *__restrict__ out, int64_t m)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int q = 0; q < world; ++q) {
        const int64_t cnt = (int64_t)*reinterpret_cast<const unsigned int *>(local + counts_off + 4 * (size_t)q);
        const int2 *pairs = reinterpret_cast<const int2 *>(local + pairs_off + (size_t)q * slab_bytes);
        for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += stride) {
            const int2 pr = pairs[k];
            if (pr.x >= 0 && pr.x < m) out[pr.x] = (int64_t)q * leaves_per_shard + pr.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Slices layout: the batch is DISTRIBUTED over the ranks (rank r holds elements [r m, (r+1) m) of the global batch, as
// data-parallel learners produce it), so no rank ever touches more than its own 1/G of the batch plus the elements it
// owns.  One exchange = count matrix first (M[r][q] = elements of rank r's slice that belong to shard q, plus rank r's
// validation flags in column `world`), then every kernel derives its offsets from the matrix on the device:
//   send    slice partitioned by owner (st_shard_partition_device) -> owner q's receive area at sum_{r' < r} M[r'][q]:
//           the receive area is the owner's elements in GLOBAL batch order (rank-major = batch order), contiguous
//   return  the owner's results (differences d_j / found leaves) in receive order -> source r's return area at
//           sum_{q' < q} M[r][q'] = the position the element had in r's partitioned slice
//   unpermute   result of slice element j = ret[start[shard[j]] + slot[j]]
// ---------------------------------------------------------------------------------------------
__global__ void slices_row_kernel(peer_ptrs pp, int rank, size_t mat_off, const unsigned int *__restrict__ row,
                                  int channel, unsigned int value)
{
    // one CTA: row r of the matrix goes to every rank, then the release flag
    const int w = pp.n;
    for (int i = threadIdx.x; i < w * (w + 1); i += blockDim.x) {
        const int q = i / (w + 1), c = i - q * (w + 1);
        reinterpret_cast<unsigned int *>(pp.base[q] + mat_off)[rank * (w + 1) + c] = row[c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < w) {
        unsigned int *f = reinterpret_cast<unsigned int *>(pp.base[threadIdx.x]) + channel * PEER_MAX + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
    }
}

struct slices_arr {
    const void *src[2];
    size_t dst_off[2];
    int esize[2];
    int n;
};

template <typename V> __device__ __forceinline__ void copy_elem(char *dst, const void *src, int64_t di, int64_t si)
{
    reinterpret_cast<V *>(dst)[di] = reinterpret_cast<const V *>(src)[si];
}
__device__ __forceinline__ void copy_sized(char *dst, const void *src, int64_t di, int64_t si, int es)
{
    switch (es) {
    case 1: copy_elem<uint8_t>(dst, src, di, si); break;
    case 2: copy_elem<uint16_t>(dst, src, di, si); break;
    case 4: copy_elem<uint32_t>(dst, src, di, si); break;
    default: copy_elem<uint64_t>(dst, src, di, si); break;
    }
}

// forward: element k of the partitioned slice (k in [0, sum_q M[rank][q])) belongs to destination q with
// start[q] <= k < start[q + 1]; it lands at element base[q] + (k - start[q]) of q's receive area.
__global__ void __launch_bounds__(256) slices_send_kernel(peer_ptrs pp, int rank, size_t mat_off, slices_arr a,
                                                          unsigned int *__restrict__ total_out)
{
    __shared__ unsigned int s_start[PEER_MAX + 1], s_base[PEER_MAX];
    const int w = pp.n;
    const unsigned int *mat = reinterpret_cast<const unsigned int *>(pp.base[rank] + mat_off);
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int q = 0; q < w; ++q) { s_start[q] = run; run += mat[rank * (w + 1) + q]; }
        s_start[w] = run;
    }
    if (threadIdx.x < w) {
        unsigned int b = 0;
        for (int r = 0; r < rank; ++r) b += mat[r * (w + 1) + threadIdx.x];
        s_base[threadIdx.x] = b;
    }
    if (blockIdx.x == 0 && threadIdx.x == 32 && total_out) {
        unsigned int tot = 0;
        for (int r = 0; r < w; ++r) tot += mat[r * (w + 1) + rank];
        *total_out = tot;                       // elements this rank receives (it owns)
    }
    __syncthreads();
    const int64_t total = s_start[w];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        int q = 0;
        while (k >= (int64_t)s_start[q + 1]) ++q;
        const int64_t di = (int64_t)s_base[q] + (k - (int64_t)s_start[q]);
        for (int i = 0; i < a.n; ++i) copy_sized(pp.base[q] + a.dst_off[i], a.src[i], di, k, a.esize[i]);
    }
}

// backward: element k of the owner's receive order came from source r with rstart[r] <= k < rstart[r + 1]; its
// result goes to element sum_{q' < rank} M[r][q'] + (k - rstart[r]) of r's return area.
__global__ void __launch_bounds__(256) slices_return_kernel(peer_ptrs pp, int rank, size_t mat_off, const void *__restrict__ src,
                                                            int esize, size_t dst_off)
{
    __shared__ unsigned int s_rstart[PEER_MAX + 1], s_base[PEER_MAX];
    const int w = pp.n;
    const unsigned int *mat = reinterpret_cast<const unsigned int *>(pp.base[rank] + mat_off);
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int r = 0; r < w; ++r) { s_rstart[r] = run; run += mat[r * (w + 1) + rank]; }
        s_rstart[w] = run;
    }
    if (threadIdx.x < w) {
        unsigned int b = 0;
        for (int q = 0; q < rank; ++q) b += mat[threadIdx.x * (w + 1) + q];
        s_base[threadIdx.x] = b;
    }
    __syncthreads();
    const int64_t total = s_rstart[w];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        int r = 0;
        while (k >= (int64_t)s_rstart[r + 1]) ++r;
        copy_sized(pp.base[r] + dst_off, src, (int64_t)s_base[r] + (k - (int64_t)s_rstart[r]), k, esize);
    }
}

// result of slice element j = ret[start[shard[j]] + slot[j]], start from this rank's own row of the matrix.
// INDEX: ret holds int32 local leaves, out[j] = shard * leaves_per_shard + leaf (int64); otherwise a plain copy.
template <typename V, bool INDEX>
__global__ void __launch_bounds__(256) slices_unpermute_kernel(const unsigned int *__restrict__ row, int world,
                                                               const uint8_t *__restrict__ shard8,
                                                               const int32_t *__restrict__ slot, int64_t m,
                                                               const V *__restrict__ ret, int64_t leaves_per_shard,
                                                               void *__restrict__ out, uint8_t *__restrict__ shard_copy)
{
    __shared__ unsigned int s_start[PEER_MAX];
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int q = 0; q < world; ++q) { s_start[q] = run; run += row[q]; }
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        const int32_t sl = slot[j];
        const int q = shard8[j];
        if constexpr (INDEX) {
            reinterpret_cast<int64_t *>(out)[j] =
                sl >= 0 ? (int64_t)q * leaves_per_shard + (int64_t)ret[(size_t)s_start[q] + (size_t)sl] : (int64_t)-1;
        } else {
            V v{};
            if (sl >= 0) v = ret[(size_t)s_start[q] + (size_t)sl];
            reinterpret_cast<V *>(out)[j] = v;
            if (shard_copy) shard_copy[j] = (uint8_t)q;   // next to the differences, where the other ranks can read it
        }
    }
}

// Slices layout, sample: route this rank's OWN queries on the top tree and push every left-over prefix straight into
// its owner's receive region (one region per source rank, so the append position is a local counter; any order: the
// slice position travels with the prefix).  Lanes of a warp that go to the same owner take consecutive places, so the
// remote stores of a warp are a few full-width runs.  Region of source r on the owner: element [r * cap + k].
template <typename T>
__global__ void __launch_bounds__(256) route_push_kernel(const T *__restrict__ H, int nleaves, int depth,
                                                         const double *__restrict__ u, int64_t m, int rank, peer_ptrs pp,
                                                         size_t resid_off, size_t pos_off, int64_t cap,
                                                         unsigned int *__restrict__ cnt)
{
    constexpr int K = 4;                               // queries per thread and round
    __shared__ T top[2 * PEER_MAX];
    __shared__ unsigned int s_cnt[PEER_MAX], s_base[PEER_MAX], s_pre[PEER_MAX + 1];
    __shared__ T s_res[256 * K];
    __shared__ int32_t s_pos[256 * K];
    __shared__ unsigned char s_dst[256 * K];
    for (int i = threadIdx.x; i < 2 * nleaves; i += blockDim.x) top[i] = H[i];
    if (threadIdx.x < PEER_MAX) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const T root = top[1];
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (m + stride * K - 1) / (stride * K);
    for (int64_t r = 0; r < rounds; ++r) {
        int q[K];
        T p[K];
        unsigned int off[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int64_t i = (r * K + k) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            q[k] = -1;
            p[k] = T(0);
            off[k] = 0;
            if (i < m) {
                // (root * u).astype(dtype), sumtree_array.py:420
                if constexpr (sizeof(T) == 4 && is_float_t<T>::value) p[k] = __double2float_rn(__dmul_rn((double)root, u[i]));
                else if constexpr (is_float_t<T>::value) p[k] = __dmul_rn(root, u[i]);
                else p[k] = (T)__dmul_rn((double)root, u[i]);
                uint32_t h = 1;
                for (int lev = 0; lev < depth; ++lev) {
                    const T lv = top[2 * h];
                    if (p[k] >= lv) { p[k] = t_sub<T>(p[k], lv); h = 2 * h + 1; } else { h = 2 * h; }
                }
                q[k] = (int)h - nleaves;
            }
            // place inside this CTA's run for owner q: one shared-memory atomic per (warp, owner)
            const unsigned int peers = __match_any_sync(0xffffffffu, q[k]);
            if (q[k] >= 0) {
                const int leader = __ffs(peers) - 1;
                unsigned int b = 0;
                if (lane == leader) b = atomicAdd(&s_cnt[q[k]], (unsigned int)__popc(peers));
                off[k] = __shfl_sync(peers, b, leader) + __popc(peers & ((1u << lane) - 1u));
            }
        }
        __syncthreads();
        // ONE global atomic per (CTA, owner, round): a counter bumped by every warp serialises ~2^15 atomics per 2^20 queries
        if (threadIdx.x < nleaves) {
            const unsigned int c = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = c ? atomicAdd(cnt + threadIdx.x, c) : 0u;
        }
        if (threadIdx.x == 0) {
            unsigned int run = 0;
            for (int d = 0; d < nleaves; ++d) { s_pre[d] = run; run += s_cnt[d]; }
            s_pre[nleaves] = run;
        }
        __syncthreads();
        // stage the round's elements grouped by owner, so that consecutive threads store consecutive elements of ONE
        // owner's run (with 8 owners a warp's own lanes would give 16-byte remote stores)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (q[k] >= 0) {
                const unsigned int sl = s_pre[q[k]] + off[k];
                s_res[sl] = p[k];
                s_pos[sl] = (int32_t)((r * K + k) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
                s_dst[sl] = (unsigned char)q[k];
            }
        }
        __syncthreads();
        const unsigned int total = s_pre[nleaves];
        for (unsigned int e = threadIdx.x; e < total; e += blockDim.x) {
            const int d = s_dst[e];
            const size_t at = (size_t)rank * (size_t)cap + s_base[d] + (e - s_pre[d]);
            reinterpret_cast<T *>(pp.base[d] + resid_off)[at] = s_res[e];
            reinterpret_cast<int32_t *>(pp.base[d] + pos_off)[at] = s_pos[e];
        }
        if (threadIdx.x < nleaves) s_cnt[threadIdx.x] = 0;
        __syncthreads();   // the staging arrays and counters are reused by the next round
    }
}

// the counts of what route_push_kernel appended go to the owners (cnt_off + 4 * rank on owner q) and, for the scatter
// of the results, into this rank's own buffer (sent_off + 4 * q); then the release flag
__global__ void slices_counts_kernel(peer_ptrs pp, int rank, const unsigned int *__restrict__ cnt, size_t cnt_off,
                                     size_t sent_off, int channel, unsigned int value)
{
    const int q = threadIdx.x;
    if (q < pp.n) {
        const unsigned int c = cnt[q];
        reinterpret_cast<unsigned int *>(pp.base[q] + cnt_off)[rank] = c;
        reinterpret_cast<unsigned int *>(pp.base[rank] + sent_off)[q] = c;
    }
    __threadfence_system();
    __syncthreads();
    if (q < pp.n) {
        unsigned int *f = reinterpret_cast<unsigned int *>(pp.base[q]) + channel * PEER_MAX + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
    }
}

// Top-tree routing of a replicated query batch (the first g levels of the reference descent,
// pyx:96-113, on the replicated top tree) fused with the selection of this rank's queries:
// the queries whose descent ends in top leaf `rank` are appended (any order: every result is
// addressed by its batch position later) with the prefix that is left over for the local descent.
template <typename T>
__global__ void __launch_bounds__(256) route_compact_kernel(const T *__restrict__ H, int nleaves, int depth,
                                                            const double *__restrict__ u, int64_t m, int rank,
                                                            T *__restrict__ own_resid, int32_t *__restrict__ own_pos,
                                                            unsigned int *__restrict__ own_count)
{
    __shared__ T top[2 * PEER_MAX];
    __shared__ unsigned int s_wcount[8], s_base;
    for (int i = threadIdx.x; i < 2 * nleaves; i += blockDim.x) top[i] = H[i];
    __syncthreads();
    const T root = top[1];
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int K = 4;                               // queries per thread and round
    const int64_t rounds = (m + stride * K - 1) / (stride * K);
    const int warp = threadIdx.x >> 5;
    for (int64_t r = 0; r < rounds; ++r) {
        bool mine[K];
        T p[K];
        unsigned int bal[K], wc = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int64_t i = (r * K + k) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            mine[k] = false;
            p[k] = T(0);
            if (i < m) {
                // (root * u).astype(dtype), sumtree_array.py:420
                if constexpr (sizeof(T) == 4 && is_float_t<T>::value) p[k] = __double2float_rn(__dmul_rn((double)root, u[i]));
                else if constexpr (is_float_t<T>::value) p[k] = __dmul_rn(root, u[i]);
                else p[k] = (T)__dmul_rn((double)root, u[i]);
                uint32_t h = 1;
                for (int lev = 0; lev < depth; ++lev) {
                    const T lv = top[2 * h];
                    if (p[k] >= lv) { p[k] = t_sub<T>(p[k], lv); h = 2 * h + 1; } else { h = 2 * h; }
                }
                mine[k] = ((int)h - nleaves == rank);
            }
            bal[k] = __ballot_sync(0xffffffffu, mine[k]);
            wc += __popc(bal[k]);
        }
        // append: ONE atomic per CTA and round (a counter bumped by every warp serialised 2^16 atomics per 2^21 queries)
        if (lane == 0) s_wcount[warp] = wc;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
            for (int w = 0; w < 8; ++w) { const unsigned int c = s_wcount[w]; s_wcount[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(own_count, tot) : 0u;
        }
        __syncthreads();
        unsigned int at = s_base + s_wcount[warp];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (mine[k]) {
                const unsigned int a = at + __popc(bal[k] & ((1u << lane) - 1u));
                own_resid[a] = p[k];
                own_pos[a] = (int32_t)((r * K + k) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
            }
            at += __popc(bal[k]);
        }
        __syncthreads();   // s_wcount / s_base are reused by the next round
    }
}

cudaError_t launch_route_compact(const st_tree *top, const double *u, int64_t m, int rank, void *own_resid,
                                 int32_t *own_pos, unsigned int *own_count, cudaStream_t s)
{
    if (top->n > PEER_MAX || (top->n & (top->n - 1))) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(own_count, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess || m <= 0) return e;
    ST_DISPATCH(top->dtype, {
        ++g_kernel_launches;
        route_compact_kernel<T><<<grid_for(m, 256 * 4, (int64_t)top->sm_count * 8), 256, 0, s>>>(
            (const T *)top->heap, (int)top->n, top->depth, u, m, rank, (T *)own_resid, own_pos, own_count);
    });
    return cudaGetLastError();
}

}  // namespace st

namespace {
thread_local char g_perr[256] = "";
int pfail(int status, const char *msg, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(g_perr, sizeof(g_perr), "%s: %s", msg, cudaGetErrorString(e));
    else snprintf(g_perr, sizeof(g_perr), "%s", msg);
    return status;
}
#define PCU(call)                                                     \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return pfail(ST_ERR_CUDA, #call, e__); \
    } while (0)

struct dev_guard {
    int prev = -1;
    explicit dev_guard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); }
    ~dev_guard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

const char *st_peer_last_error(void) { return g_perr; }

int st_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int st_peer_create(int device, int rank, int world, int64_t payload_bytes, st_peer **out, unsigned char *handle_out)
{
    if (!out || !handle_out || world < 1 || world > st::PEER_MAX || rank < 0 || rank >= world || payload_bytes < 0)
        return pfail(ST_ERR_ARG, "st_peer_create: bad argument");
    *out = nullptr;
    dev_guard g(device);
    st_peer *p = new (std::nothrow) st_peer();
    if (!p) return pfail(ST_ERR_ARG, "st_peer_create: out of host memory");
    p->device = device; p->rank = rank; p->world = world;
    p->payload_bytes = ((size_t)payload_bytes + 255) & ~(size_t)255;
    const size_t total = st::PEER_FLAG_BYTES + p->payload_bytes;
    cudaError_t e = cudaMalloc((void **)&p->local, total);      // plain cudaMalloc: IPC-exportable
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, total);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_err, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->d_err, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&p->h_err, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&p->h_mat, sizeof(unsigned int) * st::PEER_MAX * (st::PEER_MAX + 1));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->mat_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_tab, sizeof(void *) * 4 * st::PEER_MAX);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        st_peer_destroy(p);
        return pfail(ST_ERR_CUDA, "st_peer_create", e);
    }
    memcpy(handle_out, &h, sizeof(h));
    p->ptrs.n = world;
    p->ptrs.base[rank] = p->local;
    *out = p;
    return ST_OK;
}

int st_peer_connect(st_peer *p, const unsigned char *handles)
{
    if (!p || !handles) return pfail(ST_ERR_ARG, "st_peer_connect: bad argument");
    dev_guard g(p->device);
    for (int q = 0; q < p->world; ++q) {
        if (q == p->rank || p->ptrs.base[q]) continue;   // own buffer / adopted (same-process) peer
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        PCU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->ptrs.base[q] = (char *)ptr;
        p->opened[q] = true;
    }
    return ST_OK;
}

int st_peer_adopt(st_peer *p, int q, void *base)
{
    if (!p || q < 0 || q >= p->world || q == p->rank || !base) return pfail(ST_ERR_ARG, "st_peer_adopt: bad argument");
    p->ptrs.base[q] = (char *)base;
    return ST_OK;
}

void *st_peer_base(const st_peer *p) { return p ? p->local : nullptr; }

int st_peer_disconnect(st_peer *p)
{
    if (!p) return ST_OK;
    dev_guard g(p->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < p->world; ++q) {
        if (p->opened[q]) cudaIpcCloseMemHandle(p->ptrs.base[q]);
        p->opened[q] = false;
        if (q != p->rank) p->ptrs.base[q] = nullptr;
    }
    return ST_OK;
}

int st_peer_destroy(st_peer *p)
{
    if (!p) return ST_OK;
    st_peer_disconnect(p);
    dev_guard g(p->device);
    if (p->local) cudaFree(p->local);
    if (p->d_err) cudaFree(p->d_err);
    if (p->h_err) cudaFreeHost(p->h_err);
    if (p->h_mat) cudaFreeHost(p->h_mat);
    if (p->mat_ev) cudaEventDestroy(p->mat_ev);
    if (p->d_tab) cudaFree(p->d_tab);
    delete p;
    return ST_OK;
}

void *st_peer_payload(const st_peer *p) { return p ? p->local + st::PEER_FLAG_BYTES : nullptr; }
int64_t st_peer_payload_bytes(const st_peer *p) { return p ? (int64_t)p->payload_bytes : -1; }

int st_peer_broadcast(st_peer *p, int64_t offset, int64_t bytes, void *stream)
{
    if (!p || offset < 0 || bytes < 0 || (offset & 3) || (bytes & 3) || (size_t)(offset + bytes) > p->payload_bytes)
        return pfail(ST_ERR_ARG, "st_peer_broadcast: bad argument (offset and size must be multiples of 4 inside the payload)");
    if (bytes == 0 || p->world == 1) return ST_OK;
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    const size_t off = st::PEER_FLAG_BYTES + (size_t)offset;
    const int64_t both = offset | bytes;
    cudaStream_t s = (cudaStream_t)stream;
    if (!(both & 15))
        st::peer_broadcast_kernel<int4><<<st::grid_for(bytes / 16, 256, 592), 256, 0, s>>>(p->ptrs, p->rank, off, (size_t)bytes);
    else if (!(both & 7))
        st::peer_broadcast_kernel<int2><<<st::grid_for(bytes / 8, 256, 592), 256, 0, s>>>(p->ptrs, p->rank, off, (size_t)bytes);
    else
        st::peer_broadcast_kernel<int><<<st::grid_for(bytes / 4, 256, 592), 256, 0, s>>>(p->ptrs, p->rank, off, (size_t)bytes);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_peer_scatter(st_peer *p, const void *src_dev, const int32_t *pos_dev, int64_t count, const uint32_t *count_dev,
                    int esize, int64_t offset, int64_t span_elems, void *stream)
{
    if (!p || count < 0 || offset < 0 || (offset % 8) || (esize != 1 && esize != 2 && esize != 4 && esize != 8) ||
        (count > 0 && (!src_dev || !pos_dev)) || (size_t)(offset + span_elems * esize) > p->payload_bytes)
        return pfail(ST_ERR_ARG, "st_peer_scatter: bad argument");
    if (count == 0) return ST_OK;
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    const unsigned grid = st::grid_for(count, 256, 1184);
    const size_t off = st::PEER_FLAG_BYTES + (size_t)offset;
    cudaStream_t s = (cudaStream_t)stream;
    switch (esize) {
    case 1: st::peer_scatter_kernel<uint8_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint8_t *)src_dev, pos_dev, count, count_dev); break;
    case 2: st::peer_scatter_kernel<uint16_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint16_t *)src_dev, pos_dev, count, count_dev); break;
    case 4: st::peer_scatter_kernel<uint32_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint32_t *)src_dev, pos_dev, count, count_dev); break;
    default: st::peer_scatter_kernel<uint64_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint64_t *)src_dev, pos_dev, count, count_dev); break;
    }
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_peer_put(st_peer *p, const void *src_dev, int64_t count, const uint32_t *count_dev, int esize, int64_t dst_offset,
                int64_t count_offset, void *stream)
{
    if (!p || count < 0 || dst_offset < 0 || (dst_offset % esize) || (esize != 4 && esize != 8) || (count > 0 && !src_dev) ||
        (size_t)(dst_offset + count * esize) > p->payload_bytes ||
        (count_offset >= 0 && ((count_offset & 3) || (size_t)count_offset + 4 > p->payload_bytes)))
        return pfail(ST_ERR_ARG, "st_peer_put: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    const unsigned grid = st::grid_for(count > 0 ? count : 1, 256, 1184);
    const size_t off = st::PEER_FLAG_BYTES + (size_t)dst_offset;
    const int64_t coff = count_offset >= 0 ? (int64_t)st::PEER_FLAG_BYTES + count_offset : -1;
    cudaStream_t s = (cudaStream_t)stream;
    if (esize == 4)
        st::peer_put_kernel<uint32_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint32_t *)src_dev, count, count_dev, p->rank, coff);
    else
        st::peer_put_kernel<uint64_t><<<grid, 256, 0, s>>>(p->ptrs, off, (const uint64_t *)src_dev, count, count_dev, p->rank, coff);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_peer_scatter_pairs(st_peer *p, int64_t pairs_offset, int64_t slab_bytes, int64_t counts_offset,
                          int64_t leaves_per_shard, int64_t *out_dev, int64_t m, void *stream)
{
    if (!p || pairs_offset < 0 || (pairs_offset & 7) || slab_bytes < 0 || (slab_bytes & 7) || counts_offset < 0 ||
        (counts_offset & 3) || !out_dev || m < 0 ||
        (size_t)(pairs_offset + slab_bytes * p->world) > p->payload_bytes ||
        (size_t)(counts_offset + 4 * p->world) > p->payload_bytes)
        return pfail(ST_ERR_ARG, "st_peer_scatter_pairs: bad argument");
    if (m == 0) return ST_OK;
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    st::peer_scatter_pairs_kernel<<<st::grid_for(m / p->world + 1, 256, 1184), 256, 0, (cudaStream_t)stream>>>(
        p->local, st::PEER_FLAG_BYTES + (size_t)pairs_offset, (size_t)slab_bytes, st::PEER_FLAG_BYTES + (size_t)counts_offset,
        p->world, leaves_per_shard, out_dev, m);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_search_pairs_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                           const int32_t *pos_dev, int32_t *pairs_out_dev, void *stream)
{
    if (!t || m_max < 0 || !m_dev || (m_max > 0 && (!vals_dev || !pos_dev || !pairs_out_dev)))
        return pfail(ST_ERR_ARG, "st_search_pairs_device: bad argument");
    dev_guard g(t->device);
    st::search_push push;
    push.pos = pos_dev;
    push.pairs = pairs_out_dev;
    PCU(st::launch_search(t, vals_dev, m_max, nullptr, (cudaStream_t)stream, m_dev, &push));
    return ST_OK;
}

// ---- slices layout (batch distributed over the ranks) ---------------------------------------------------------------
namespace {
bool region_ok(const st_peer *p, int64_t off, int64_t bytes, int align)
{
    return off >= 0 && bytes >= 0 && !(off % align) && (size_t)(off + bytes) <= p->payload_bytes;
}
int64_t matrix_bytes(const st_peer *p) { return (int64_t)p->world * (p->world + 1) * 4; }
}  // namespace

int st_slices_rows(st_peer *p, const uint32_t *row_dev, int64_t matrix_offset, int channel, unsigned int value, void *stream)
{
    if (!p || !row_dev || channel < 0 || channel >= st::PEER_CHANNELS || !region_ok(p, matrix_offset, matrix_bytes(p), 4))
        return pfail(ST_ERR_ARG, "st_slices_rows: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    st::slices_row_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p->ptrs, p->rank, st::PEER_FLAG_BYTES + (size_t)matrix_offset,
                                                              row_dev, channel, value);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_slices_matrix_fetch(st_peer *p, int64_t matrix_offset, void *stream)
{
    if (!p || !region_ok(p, matrix_offset, matrix_bytes(p), 4)) return pfail(ST_ERR_ARG, "st_slices_matrix_fetch: bad argument");
    dev_guard g(p->device);
    PCU(cudaMemcpyAsync(p->h_mat, p->local + st::PEER_FLAG_BYTES + matrix_offset, (size_t)matrix_bytes(p), cudaMemcpyDeviceToHost,
                        (cudaStream_t)stream));
    PCU(cudaEventRecord(p->mat_ev, (cudaStream_t)stream));
    return ST_OK;
}

int st_slices_matrix_wait(st_peer *p, uint32_t *matrix_out_host)
{
    if (!p || !matrix_out_host) return pfail(ST_ERR_ARG, "st_slices_matrix_wait: bad argument");
    dev_guard g(p->device);
    PCU(cudaEventSynchronize(p->mat_ev));     // the fetch only, not what was enqueued behind it
    memcpy(matrix_out_host, p->h_mat, (size_t)matrix_bytes(p));
    return ST_OK;
}

int st_slices_send(st_peer *p, int64_t matrix_offset, const void *src0_dev, int esize0, int64_t dst_offset0, const void *src1_dev,
                   int esize1, int64_t dst_offset1, int64_t m_max, uint32_t *total_out_dev, void *stream)
{
    auto es_ok = [](int e) { return e == 1 || e == 2 || e == 4 || e == 8; };
    if (!p || m_max < 0 || !region_ok(p, matrix_offset, matrix_bytes(p), 4) || !src0_dev || !es_ok(esize0) ||
        !region_ok(p, dst_offset0, 0, 8) || (src1_dev && (!es_ok(esize1) || !region_ok(p, dst_offset1, 0, 8))))
        return pfail(ST_ERR_ARG, "st_slices_send: bad argument");
    dev_guard g(p->device);
    st::slices_arr a{};
    a.n = src1_dev ? 2 : 1;
    a.src[0] = src0_dev; a.esize[0] = esize0; a.dst_off[0] = st::PEER_FLAG_BYTES + (size_t)dst_offset0;
    a.src[1] = src1_dev; a.esize[1] = esize1; a.dst_off[1] = st::PEER_FLAG_BYTES + (size_t)dst_offset1;
    ++st::g_kernel_launches;
    st::slices_send_kernel<<<st::grid_for(m_max > 0 ? m_max : 1, 256, 1184), 256, 0, (cudaStream_t)stream>>>(
        p->ptrs, p->rank, st::PEER_FLAG_BYTES + (size_t)matrix_offset, a, total_out_dev);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_slices_return(st_peer *p, int64_t matrix_offset, const void *src_dev, int esize, int64_t dst_offset, int64_t total_max,
                     void *stream)
{
    if (!p || total_max < 0 || !region_ok(p, matrix_offset, matrix_bytes(p), 4) || !src_dev ||
        (esize != 1 && esize != 2 && esize != 4 && esize != 8) || !region_ok(p, dst_offset, 0, 8))
        return pfail(ST_ERR_ARG, "st_slices_return: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    st::slices_return_kernel<<<st::grid_for(total_max > 0 ? total_max : 1, 256, 1184), 256, 0, (cudaStream_t)stream>>>(
        p->ptrs, p->rank, st::PEER_FLAG_BYTES + (size_t)matrix_offset, src_dev, esize, st::PEER_FLAG_BYTES + (size_t)dst_offset);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_slices_unpermute(st_peer *p, int64_t matrix_offset, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                        int64_t ret_offset, int esize, int64_t leaves_per_shard, void *out_dev, int64_t shard_copy_offset,
                        void *stream)
{
    if (!p || m < 0 || (shard_copy_offset >= 0 && !region_ok(p, shard_copy_offset, m, 1)) || !region_ok(p, matrix_offset, matrix_bytes(p), 4) || !region_ok(p, ret_offset, m * (leaves_per_shard > 0 ? 4 : esize), 8) ||
        (m > 0 && (!shard_dev || !slot_dev || !out_dev)) || (esize != 1 && esize != 2 && esize != 4 && esize != 8))
        return pfail(ST_ERR_ARG, "st_slices_unpermute: bad argument");
    if (m == 0) return ST_OK;
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    const unsigned int *row = reinterpret_cast<const unsigned int *>(p->local + st::PEER_FLAG_BYTES + matrix_offset) +
                              (size_t)p->rank * (p->world + 1);
    const char *ret = p->local + st::PEER_FLAG_BYTES + ret_offset;
    uint8_t *sc = shard_copy_offset >= 0 ? (uint8_t *)(p->local + st::PEER_FLAG_BYTES + shard_copy_offset) : nullptr;
    const unsigned grid = st::grid_for(m, 256, 1184);
    cudaStream_t s = (cudaStream_t)stream;
    if (leaves_per_shard > 0)
        st::slices_unpermute_kernel<int32_t, true><<<grid, 256, 0, s>>>(row, p->world, shard_dev, slot_dev, m, (const int32_t *)ret, leaves_per_shard, out_dev, nullptr);
    else if (esize == 1)
        st::slices_unpermute_kernel<uint8_t, false><<<grid, 256, 0, s>>>(row, p->world, shard_dev, slot_dev, m, (const uint8_t *)ret, 0, out_dev, sc);
    else if (esize == 2)
        st::slices_unpermute_kernel<uint16_t, false><<<grid, 256, 0, s>>>(row, p->world, shard_dev, slot_dev, m, (const uint16_t *)ret, 0, out_dev, sc);
    else if (esize == 4)
        st::slices_unpermute_kernel<uint32_t, false><<<grid, 256, 0, s>>>(row, p->world, shard_dev, slot_dev, m, (const uint32_t *)ret, 0, out_dev, sc);
    else
        st::slices_unpermute_kernel<uint64_t, false><<<grid, 256, 0, s>>>(row, p->world, shard_dev, slot_dev, m, (const uint64_t *)ret, 0, out_dev, sc);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_slices_table(st_peer *p, int parity, int64_t d_offset, int64_t leaf_offset)
{
    if (!p || (parity != 0 && parity != 1) || !region_ok(p, d_offset, 0, 8) || !region_ok(p, leaf_offset, 0, 8))
        return pfail(ST_ERR_ARG, "st_slices_table: bad argument");
    dev_guard g(p->device);
    const void *tab[2 * st::PEER_MAX] = {};
    for (int q = 0; q < p->world; ++q) {
        if (!p->ptrs.base[q]) return pfail(ST_ERR_ARG, "st_slices_table: peers are not connected yet");
        tab[q] = p->ptrs.base[q] + st::PEER_FLAG_BYTES + d_offset;
        tab[st::PEER_MAX + q] = p->ptrs.base[q] + st::PEER_FLAG_BYTES + leaf_offset;
    }
    PCU(cudaMemcpy(p->d_tab + (size_t)parity * 2 * st::PEER_MAX, tab, sizeof(tab), cudaMemcpyHostToDevice));
    return ST_OK;
}

int st_top_fold_maps_slice_device(st_tree *t, const uint8_t *leaf_dev, const void *deltas_dev, int64_t m_slice,
                                  int64_t record_first, int64_t nchunks, void *recs_dev, void *stream)
{
    if (!t || m_slice < 0 || record_first < 0 || nchunks < 0 || !recs_dev || (m_slice > 0 && (!leaf_dev || !deltas_dev)) ||
        nchunks * 1024 < m_slice || (record_first + nchunks) * 1024 > 0x7fffffffLL)
        return pfail(ST_ERR_ARG, "st_top_fold_maps_slice_device: bad argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu);
    dev_guard g(t->device);
    // chunk c of the GLOBAL batch = chunk c - record_first of this slice: shift the array origins accordingly
    const int64_t shift = record_first * 1024;
    const char *leaf0 = (const char *)leaf_dev - shift;
    const char *d0 = (const char *)deltas_dev - shift * t->esize;
    PCU(st::launch_fold_by_leaf_maps(t, leaf0, 1, d0, shift + m_slice, (int)record_first, (int)(record_first + nchunks), recs_dev,
                                     (cudaStream_t)stream));
    return ST_OK;
}

int st_top_fold_apply_slices_device(st_tree *t, st_peer *p, int parity, int64_t span, int64_t m_slice, const void *recs_dev,
                                    void *stream)
{
    if (!t || !p || (parity != 0 && parity != 1) || span <= 0 || (span % 1024) || m_slice < 0 || m_slice > span || !recs_dev ||
        span * p->world > 0x7fffffffLL)
        return pfail(ST_ERR_ARG, "st_top_fold_apply_slices_device: bad argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu);
    dev_guard g(t->device);
    const void *const *tab = p->d_tab + (size_t)parity * 2 * st::PEER_MAX;
    PCU(st::launch_fold_by_leaf_apply(t, nullptr, 1, nullptr, span * p->world, recs_dev, (cudaStream_t)stream, tab,
                                      tab + st::PEER_MAX, span, m_slice));
    return ST_OK;
}

int st_search_leaf32_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev, int32_t *out_dev,
                            void *stream)
{
    if (!t || m_max < 0 || !m_dev || (m_max > 0 && (!vals_dev || !out_dev)))
        return pfail(ST_ERR_ARG, "st_search_leaf32_device: bad argument");
    dev_guard g(t->device);
    st::search_push push;
    push.leaf32 = out_dev;
    PCU(st::launch_search(t, vals_dev, m_max, nullptr, (cudaStream_t)stream, m_dev, &push));
    return ST_OK;
}

int st_route_push_device(const st_tree *top, const double *u_dev, int64_t m, st_peer *p, int64_t resid_offset, int64_t pos_offset,
                         int64_t cap, uint32_t *cnt_dev, void *stream)
{
    if (!top || !p || m < 0 || m > cap || cap > 0x7fffffffLL || !cnt_dev || (m > 0 && !u_dev) || top->n != p->world ||
        !region_ok(p, resid_offset, (int64_t)p->world * cap * top->esize, 8) || !region_ok(p, pos_offset, (int64_t)p->world * cap * 4, 8))
        return pfail(ST_ERR_ARG, "st_route_push_device: bad argument");
    if (top->n > st::PEER_MAX || (top->n & (top->n - 1))) return pfail(ST_ERR_ARG, "st_route_push_device: not a top tree");
    dev_guard g(top->device);
    cudaStream_t s = (cudaStream_t)stream;
    PCU(cudaMemsetAsync(cnt_dev, 0, sizeof(unsigned int) * p->world, s));
    if (m == 0) return ST_OK;
    ST_DISPATCH(top->dtype, {
        ++st::g_kernel_launches;
        st::route_push_kernel<T><<<st::grid_for(m, 256 * 4, (int64_t)top->sm_count * 8), 256, 0, s>>>(
            (const T *)top->heap, (int)top->n, top->depth, u_dev, m, p->rank, p->ptrs, st::PEER_FLAG_BYTES + (size_t)resid_offset,
            st::PEER_FLAG_BYTES + (size_t)pos_offset, cap, cnt_dev);
    });
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_slices_counts(st_peer *p, const uint32_t *cnt_dev, int64_t cnt_offset, int64_t sent_offset, int channel, unsigned int value,
                     void *stream)
{
    if (!p || !cnt_dev || channel < 0 || channel >= st::PEER_CHANNELS || !region_ok(p, cnt_offset, 4 * p->world, 4) ||
        !region_ok(p, sent_offset, 4 * p->world, 4))
        return pfail(ST_ERR_ARG, "st_slices_counts: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    st::slices_counts_kernel<<<1, st::PEER_MAX, 0, (cudaStream_t)stream>>>(p->ptrs, p->rank, cnt_dev, st::PEER_FLAG_BYTES + (size_t)cnt_offset,
                                                                          st::PEER_FLAG_BYTES + (size_t)sent_offset, channel, value);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_search_segments_push_device(const st_tree *t, st_peer *p, int64_t resid_offset, int64_t pos_offset, int64_t cnt_offset,
                                   int64_t cap, int64_t back_offset, void *stream)
{
    if (!t || !p || cap <= 0 || cap > 0x7fffffffLL || !region_ok(p, resid_offset, (int64_t)p->world * cap * t->esize, 8) ||
        !region_ok(p, pos_offset, (int64_t)p->world * cap * 4, 8) || !region_ok(p, cnt_offset, 4 * p->world, 4) ||
        !region_ok(p, back_offset, (int64_t)p->world * cap * 8, 8))
        return pfail(ST_ERR_ARG, "st_search_segments_push_device: bad argument");
    dev_guard g(t->device);
    st::search_push push;
    push.n = p->world;
    // result k of source r's segment goes to r's back region of THIS rank: [rank * cap + k] pairs of (slice position, leaf)
    for (int q = 0; q < p->world; ++q)
        push.base[q] = p->ptrs.base[q] + st::PEER_FLAG_BYTES + back_offset + (size_t)p->rank * (size_t)cap * 8;
    push.pos = reinterpret_cast<const int32_t *>(p->local + st::PEER_FLAG_BYTES + pos_offset);
    push.seg_cnt = reinterpret_cast<const unsigned int *>(p->local + st::PEER_FLAG_BYTES + cnt_offset);
    push.seg_stride = cap;
    PCU(st::launch_search(t, p->local + st::PEER_FLAG_BYTES + resid_offset, (int64_t)p->world * cap, nullptr, (cudaStream_t)stream,
                          nullptr, &push));
    return ST_OK;
}

int st_peer_signal(st_peer *p, int channel, unsigned int value, void *stream)
{
    if (!p || channel < 0 || channel >= st::PEER_CHANNELS) return pfail(ST_ERR_ARG, "st_peer_signal: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    st::peer_signal_kernel<<<1, st::PEER_MAX, 0, (cudaStream_t)stream>>>(p->ptrs, p->rank, channel, value);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_peer_wait(st_peer *p, int channel, unsigned int value, void *stream)
{
    if (!p || channel < 0 || channel >= st::PEER_CHANNELS) return pfail(ST_ERR_ARG, "st_peer_wait: bad argument");
    dev_guard g(p->device);
    ++st::g_kernel_launches;
    static const unsigned long long timeout_ns = [] {   // STARR_B200_WAIT_TIMEOUT_S: debugging aid (default 20 s)
        const char *e = getenv("STARR_B200_WAIT_TIMEOUT_S");
        const double sec = e ? atof(e) : 20.0;
        return (unsigned long long)((sec > 0 ? sec : 20.0) * 1e9);
    }();
    st::peer_wait_kernel<<<1, st::PEER_MAX, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned int *>(p->local), channel,
                                                                      value, p->world, timeout_ns, p->d_err);
    PCU(cudaGetLastError());
    return ST_OK;
}

int st_peer_status(st_peer *p, void *stream)
{
    if (!p) return pfail(ST_ERR_ARG, "st_peer_status: bad argument");
    dev_guard g(p->device);
    PCU(cudaMemcpyAsync(p->h_err, p->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PCU(cudaStreamSynchronize((cudaStream_t)stream));
    if (*p->h_err) return pfail(ST_ERR_CUDA, "st_peer_wait timed out: a rank of the sharded tree did not reach the exchange");
    return ST_OK;
}

int st_route_compact_device(const st_tree *top, const double *u_dev, int64_t m, int rank, void *own_resid_out_dev,
                            int32_t *own_pos_out_dev, uint32_t *own_count_out_dev, void *stream)
{
    if (!top || m < 0 || rank < 0 || !own_count_out_dev || (m > 0 && (!u_dev || !own_resid_out_dev || !own_pos_out_dev)) ||
        m > 0x7fffffffLL)
        return pfail(ST_ERR_ARG, "st_route_compact_device: bad argument");
    dev_guard g(top->device);
    PCU(st::launch_route_compact(top, u_dev, m, rank, own_resid_out_dev, own_pos_out_dev, own_count_out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_search_push_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                          const int32_t *pos_dev, int64_t index_offset, st_peer *p, int64_t payload_offset,
                          int64_t span_elems, void *stream)
{
    if (!t || !p || m_max < 0 || !m_dev || payload_offset < 0 || (payload_offset % 8) ||
        (size_t)(payload_offset + span_elems * 8) > p->payload_bytes || (m_max > 0 && (!vals_dev || !pos_dev)))
        return pfail(ST_ERR_ARG, "st_search_push_device: bad argument");
    dev_guard g(t->device);
    st::search_push push;
    push.n = p->ptrs.n;
    for (int q = 0; q < p->ptrs.n; ++q) push.base[q] = p->ptrs.base[q] + st::PEER_FLAG_BYTES + payload_offset;
    push.pos = pos_dev;
    push.index_offset = index_offset;
    PCU(st::launch_search(t, vals_dev, m_max, nullptr, (cudaStream_t)stream, m_dev, &push));
    return ST_OK;
}

}  // extern "C"
