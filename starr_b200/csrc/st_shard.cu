// st_shard.cu -- helpers of the subtree-sharded layout (no reference counterpart, the
// reference is single-process).  Every rank sees the whole batch (ids of all shards); these
// kernels turn it into
//   * the owning shard of every element (uint8) and its stable position ("slot") among the
//     elements of that shard, for ALL shards -- so that per-shard results exchanged as compact
//     slabs (one all-gather) can be put back into batch order by every rank, and
//   * this rank's own elements, compacted in batch order (ids made shard-local, payload gathered).
// Three launches: per-tile counts per shard, scan over tiles per shard, slots + compaction.
// Streaming, HBM-bound: ~20 B per element.
#include <cub/block/block_scan.cuh>

#include "st_common.cuh"
#include "st_update_defs.cuh"

namespace st {

constexpr int SH_THREADS = 256;
constexpr int SH_E = 8;
constexpr int SH_TILE = SH_THREADS * SH_E;
constexpr int SH_MAX = 64;          // shards (the top tree is limited to 64 leaves, fold_by_leaf_ok)
constexpr int SH_WARPS = SH_THREADS / 32;
constexpr uint8_t SH_NONE = 255;    // id outside every shard (the batch is flagged and skipped)

// Eight 16-bit counters (one per shard of the current group of eight shards) packed in 128 bits:
// a tile holds 2048 elements, so no field can overflow into its neighbour.
struct pk128 {
    unsigned long long lo, hi;
};
struct pk_add {
    __device__ __forceinline__ pk128 operator()(const pk128 &a, const pk128 &b) const
    {
        return pk128{a.lo + b.lo, a.hi + b.hi};
    }
};
__device__ __forceinline__ void pk_inc(pk128 &v, int k)   // k in 0..7
{
    const unsigned long long one = 1ull << (16 * (k & 3));
    if (k < 4) v.lo += one; else v.hi += one;
}
__device__ __forceinline__ unsigned int pk_get(const pk128 &v, int k)
{
    return (unsigned int)(((k < 4 ? v.lo : v.hi) >> (16 * (k & 3))) & 0xffffu);
}

// Thread t of a tile owns the SH_E = 8 consecutive elements [tile*2048 + 8t, +8): their shard ids
// travel as one packed 64-bit word.
__device__ __forceinline__ unsigned long long load_shard8(const uint8_t *__restrict__ shard8, int64_t j0, int64_t m)
{
    if (j0 + SH_E <= m) return *(const unsigned long long *)(shard8 + j0);   // 8-byte aligned: j0 % 8 == 0
    unsigned long long w = 0;
    for (int k = 0; k < SH_E; ++k)
        w |= (unsigned long long)(j0 + k < m ? shard8[j0 + k] : SH_NONE) << (8 * k);
    return w;
}

template <typename T, bool CHECK_NEG>
__global__ void __launch_bounds__(SH_THREADS) shard_count_kernel(const int64_t *__restrict__ ids, int64_t m, int shift,
                                                                 int shards, const T *__restrict__ payload,
                                                                 int64_t npayload, uint8_t *__restrict__ shard_out,
                                                                 unsigned int *__restrict__ tile_counts, int64_t ntiles,
                                                                 unsigned int *flags)
{
    __shared__ pk128 s_warp[SH_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * SH_TILE + (int64_t)tid * SH_E;
    unsigned int bad = 0;
    unsigned long long word = 0;
    int64_t id[SH_E];
    if (j0 + SH_E <= m && ((uintptr_t)ids & 15) == 0) {
        const longlong2 *p = (const longlong2 *)(ids + j0);
#pragma unroll
        for (int k = 0; k < SH_E / 2; ++k) {
            const longlong2 v = p[k];
            id[2 * k] = v.x;
            id[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SH_E; ++k) id[k] = j0 + k < m ? ids[j0 + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SH_E; ++k) {
        unsigned long long b = SH_NONE;
        if (j0 + k < m) {
            const int64_t q = id[k] >> shift;
            if (id[k] < 0 || q >= shards) bad |= STF_INDEX;
            else b = (unsigned long long)q;
        }
        word |= b << (8 * k);
    }
    if (j0 + SH_E <= m) {
        *(unsigned long long *)(shard_out + j0) = word;
    } else {
        for (int k = 0; k < SH_E && j0 + k < m; ++k) shard_out[j0 + k] = (uint8_t)(word >> (8 * k));
    }
    if constexpr (CHECK_NEG && is_signed_t<T>::value) {
        // the reference rejects a negative entry anywhere in vals, used or not (pyx:37-39)
        const int64_t stride = (int64_t)gridDim.x * SH_THREADS;
        for (int64_t k = (int64_t)blockIdx.x * SH_THREADS + tid; k < npayload; k += stride)
            if (payload[k] < T(0)) bad |= STF_NEG_VALS;
    }
    if (bad) atomicOr(flags, bad);
    for (int g0 = 0; g0 < shards; g0 += 8) {
        pk128 c{0, 0};
#pragma unroll
        for (int k = 0; k < SH_E; ++k) {
            const int b = (int)((word >> (8 * k)) & 0xff) - g0;
            if (b >= 0 && b < 8) pk_inc(c, b);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            c.lo += __shfl_xor_sync(0xffffffffu, c.lo, off);
            c.hi += __shfl_xor_sync(0xffffffffu, c.hi, off);
        }
        if (lane == 0) s_warp[warp] = c;
        __syncthreads();
        if (tid < 8 && g0 + tid < shards) {
            unsigned int sum = 0;
#pragma unroll
            for (int w = 0; w < SH_WARPS; ++w) sum += pk_get(s_warp[w], tid);
            tile_counts[(int64_t)(g0 + tid) * ntiles + blockIdx.x] = sum;
        }
        __syncthreads();
    }
}

// one CTA per shard: exclusive scan of its per-tile counts, total -> counts[shard];
// counts[shards] receives the validation flags of the count kernel
__global__ void __launch_bounds__(SH_THREADS) shard_scan_kernel(unsigned int *__restrict__ tile_counts, int64_t ntiles,
                                                                unsigned int *__restrict__ counts,
                                                                const unsigned int *__restrict__ flags)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[gridDim.x] = *flags;
    typedef cub::BlockScan<unsigned int, SH_THREADS> scan_t;
    __shared__ typename scan_t::TempStorage tmp;
    unsigned int *row = tile_counts + (int64_t)blockIdx.x * ntiles;
    unsigned int running = 0;
    for (int64_t b = 0; b < ntiles; b += SH_THREADS * 4) {
        const int64_t i0 = b + threadIdx.x * 4;
        unsigned int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = i0 + k < ntiles ? row[i0 + k] : 0u;
        unsigned int total;
        scan_t(tmp).ExclusiveSum(v, v, total);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < ntiles) row[i0 + k] = v[k] + running;
        running += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = running;
}

template <typename T>
__global__ void __launch_bounds__(SH_THREADS, 3) shard_slot_kernel(const int64_t *__restrict__ ids, int64_t m, int shift,
                                                                   int shards, int rank, const T *__restrict__ payload,
                                                                   int64_t npayload, const uint8_t *__restrict__ shard8,
                                                                   const unsigned int *__restrict__ tile_off,
                                                                   int64_t ntiles, int32_t *__restrict__ slot_out,
                                                                   int64_t *__restrict__ own_ids, T *__restrict__ own_payload,
                                                                   int32_t *__restrict__ own_pos)
{
    typedef cub::BlockScan<pk128, SH_THREADS> scan_t;
    __shared__ typename scan_t::TempStorage tmp;
    __shared__ unsigned int s_off[SH_MAX];            // elements of each shard before this tile
    __shared__ int32_t s_slot[SH_TILE];               // staged so that the global stores are coalesced
    __shared__ unsigned short s_own[SH_TILE];         // this rank's elements of the tile, in slot order
    __shared__ unsigned int s_nown;
    const int tid = threadIdx.x;
    if (tid < SH_MAX) s_off[tid] = tid < shards ? tile_off[(int64_t)tid * ntiles + blockIdx.x] : 0u;
    const int64_t tile0 = (int64_t)blockIdx.x * SH_TILE;
    const int64_t j0 = tile0 + (int64_t)tid * SH_E;
    const unsigned long long word = load_shard8(shard8, j0, m);
    // ids outside every shard (and the padding of a ragged last tile) keep slot -1
#pragma unroll
    for (int k = 0; k < SH_E; ++k) s_slot[k * SH_THREADS + tid] = -1;
    __syncthreads();
    for (int g0 = 0; g0 < shards; g0 += 8) {
        pk128 c{0, 0};
#pragma unroll
        for (int k = 0; k < SH_E; ++k) {
            const int b = (int)((word >> (8 * k)) & 0xff) - g0;
            if (b >= 0 && b < 8) pk_inc(c, b);
        }
        pk128 pre, total;
        scan_t(tmp).ExclusiveScan(c, pre, pk128{0, 0}, pk_add(), total);
        if (tid == 0 && rank >= g0 && rank < g0 + 8) s_nown = pk_get(total, rank - g0);
#pragma unroll
        for (int k = 0; k < SH_E; ++k) {
            const int b = (int)((word >> (8 * k)) & 0xff) - g0;
            if (b >= 0 && b < 8) {
                const unsigned int within = pk_get(pre, b);
                s_slot[tid * SH_E + k] = (int32_t)(s_off[b + g0] + within);
                if (b + g0 == rank) s_own[within] = (unsigned short)(tid * SH_E + k);
                pk_inc(pre, b);
            }
        }
        __syncthreads();   // tmp is reused by the next group; s_slot / s_own / s_nown are complete after the last one
    }
#pragma unroll
    for (int r = 0; r < SH_E; ++r) {
        const int i = r * SH_THREADS + tid;
        if (tile0 + i < m) slot_out[tile0 + i] = s_slot[i];
    }
    // compaction of this rank's elements: consecutive threads write consecutive slots
    if (!own_ids && !own_payload && !own_pos) return;
    const int64_t origin = (int64_t)rank << shift;
    const unsigned int nown = s_nown, first = s_off[rank];
    for (unsigned int i = tid; i < nown; i += SH_THREADS) {
        const int64_t j = tile0 + s_own[i];
        if (own_ids) own_ids[first + i] = ids[j] - origin;
        if (own_payload) own_payload[first + i] = payload[j < npayload ? j : j % npayload];
        if (own_pos) own_pos[first + i] = (int32_t)j;   // batch position: where this element's results go on every rank
    }
}

// Slices layout (the batch is DISTRIBUTED over the ranks, every rank holds 1/G of it): the same counts and slots, but
// instead of compacting one shard the kernel writes the STABLE G-way partition of the whole slice -- the elements of
// shard 0 in batch order, then those of shard 1, ... -- which is the order the slice travels in (one contiguous run
// per destination rank) and the order its results come back in.  part_ids: shard-local ids; part_payload: payload.
// Consecutive threads store consecutive elements of one destination run.
template <typename T>
__global__ void __launch_bounds__(SH_THREADS, 3) shard_partition_kernel(const int64_t *__restrict__ ids, int64_t m, int shift,
                                                                        int shards, const T *__restrict__ payload,
                                                                        const uint8_t *__restrict__ shard8,
                                                                        const unsigned int *__restrict__ tile_off,
                                                                        int64_t ntiles, const unsigned int *__restrict__ counts,
                                                                        int32_t *__restrict__ slot_out,
                                                                        int64_t *__restrict__ part_ids, T *__restrict__ part_payload)
{
    typedef cub::BlockScan<pk128, SH_THREADS> scan_t;
    __shared__ typename scan_t::TempStorage tmp;
    __shared__ unsigned int s_off[SH_MAX];            // elements of each shard before this tile
    __shared__ unsigned int s_start[SH_MAX];          // first element of each shard's run in the partition
    __shared__ unsigned int s_tpre[SH_MAX + 1];       // this tile: elements of the shards in front of shard q
    __shared__ int32_t s_slot[SH_TILE];
    __shared__ unsigned short s_order[SH_TILE];       // tile elements sorted by (shard, batch position)
    const int tid = threadIdx.x;
    if (tid < SH_MAX) s_off[tid] = tid < shards ? tile_off[(int64_t)tid * ntiles + blockIdx.x] : 0u;
    if (tid == 0) {
        unsigned int run = 0;
        for (int q = 0; q < shards; ++q) { s_start[q] = run; run += counts[q]; }
        s_tpre[0] = 0;
    }
    const int64_t tile0 = (int64_t)blockIdx.x * SH_TILE;
    const int64_t j0 = tile0 + (int64_t)tid * SH_E;
    const unsigned long long word = load_shard8(shard8, j0, m);
#pragma unroll
    for (int k = 0; k < SH_E; ++k) s_slot[k * SH_THREADS + tid] = -1;
    __syncthreads();
    unsigned int within_reg[SH_E];
    for (int g0 = 0; g0 < shards; g0 += 8) {
        pk128 c{0, 0};
#pragma unroll
        for (int k = 0; k < SH_E; ++k) {
            const int b = (int)((word >> (8 * k)) & 0xff) - g0;
            if (b >= 0 && b < 8) pk_inc(c, b);
        }
        pk128 pre, total;
        scan_t(tmp).ExclusiveScan(c, pre, pk128{0, 0}, pk_add(), total);
        if (tid == 0)
            for (int b = 0; b < 8 && g0 + b < shards; ++b) s_tpre[g0 + b + 1] = s_tpre[g0 + b] + pk_get(total, b);
#pragma unroll
        for (int k = 0; k < SH_E; ++k) {
            const int b = (int)((word >> (8 * k)) & 0xff) - g0;
            if (b >= 0 && b < 8) {
                within_reg[k] = pk_get(pre, b);
                s_slot[tid * SH_E + k] = (int32_t)(s_off[b + g0] + within_reg[k]);
                pk_inc(pre, b);
            }
        }
        __syncthreads();   // tmp is reused by the next group; s_tpre[.. g0 + 8] complete
    }
#pragma unroll
    for (int k = 0; k < SH_E; ++k) {
        const int b = (int)((word >> (8 * k)) & 0xff);
        if (b < shards) s_order[s_tpre[b] + within_reg[k]] = (unsigned short)(tid * SH_E + k);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SH_E; ++r) {
        const int i = r * SH_THREADS + tid;
        if (tile0 + i < m) slot_out[tile0 + i] = s_slot[i];
    }
    if (!part_ids && !part_payload) return;
    const unsigned int nvalid = s_tpre[shards];
    for (unsigned int i = tid; i < nvalid; i += SH_THREADS) {
        const unsigned int e = s_order[i];
        const int64_t j = tile0 + e;
        const int q = shard8[j];
        const size_t at = (size_t)s_start[q] + (size_t)s_slot[e];
        if (part_ids) part_ids[at] = ids[j] - ((int64_t)q << shift);
        if (part_payload) part_payload[at] = payload[j];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) shard_expand_values_kernel(const uint8_t *__restrict__ shard8,
                                                                  const int32_t *__restrict__ slot, int64_t m,
                                                                  const T *__restrict__ slabs, int64_t stride,
                                                                  T *__restrict__ out)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += step) {
        const int32_t s = slot[j];
        out[j] = s >= 0 ? slabs[(int64_t)shard8[j] * stride + s] : T(0);
    }
}

struct shard_starts { int64_t v[SH_MAX]; };

// out[j] = cat[starts[shard[j]] + slot[j]]: the slabs of all shards back to back (peer exchange, no padding)
template <typename T>
__global__ void __launch_bounds__(256) shard_expand_cat_kernel(const uint8_t *__restrict__ shard8,
                                                               const int32_t *__restrict__ slot, int64_t m,
                                                               const T *__restrict__ cat, const shard_starts starts,
                                                               T *__restrict__ out)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += step) {
        const int32_t s = slot[j];
        out[j] = s >= 0 ? cat[starts.v[shard8[j]] + s] : T(0);
    }
}

__global__ void __launch_bounds__(256) shard_expand_indices_kernel(const uint8_t *__restrict__ shard8,
                                                                   const int32_t *__restrict__ slot, int64_t m,
                                                                   const int32_t *__restrict__ slabs, int64_t stride,
                                                                   int64_t leaves_per_shard, int64_t *__restrict__ out)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += step) {
        const int32_t s = slot[j];
        const int64_t sh = shard8[j];
        out[j] = s >= 0 ? sh * leaves_per_shard + slabs[sh * stride + s] : (int64_t)-1;
    }
}

// -----------------------------------------------------------------------------------------
static int64_t shard_tiles(int64_t m) { return (m + SH_TILE - 1) / SH_TILE; }

size_t shard_plan_scratch_bytes(int64_t m, int shards)   // per-tile counts of every shard + one flags word
{
    return ((size_t)shards * (size_t)(shard_tiles(m) > 0 ? shard_tiles(m) : 1) + 1) * sizeof(unsigned int);
}

cudaError_t launch_shard_plan(st_tree *t, const int64_t *ids, int64_t m, int shift, int shards, int rank,
                              const void *payload, int64_t npayload, int check_negative, uint8_t *shard_out,
                              int32_t *slot_out, int64_t *own_ids_out, void *own_payload_out, void *scratch,
                              unsigned int *counts_out, cudaStream_t s, int32_t *own_pos_out)
{
    if (shards < 1 || shards > SH_MAX || rank < 0 || rank >= shards || shift < 0 || shift > 62)
        return cudaErrorInvalidValue;
    const int64_t ntiles = shard_tiles(m);
    if (ntiles <= 0 || ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    unsigned int *tile_counts = (unsigned int *)scratch;
    unsigned int *flags = tile_counts + (size_t)shards * ntiles;
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    g_kernel_launches += 3;
    ST_DISPATCH(t->dtype, {
        if (check_negative)
            shard_count_kernel<T, true><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, npayload, shard_out, tile_counts, ntiles, flags);
        else
            shard_count_kernel<T, false><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, npayload, shard_out, tile_counts, ntiles, flags);
        shard_scan_kernel<<<(unsigned)shards, SH_THREADS, 0, s>>>(tile_counts, ntiles, counts_out, flags);
        shard_slot_kernel<T><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(ids, m, shift, shards, rank, (const T *)payload,
                                                                    npayload, shard_out, tile_counts, ntiles, slot_out,
                                                                    own_ids_out, (T *)own_payload_out, own_pos_out);
    });
    return cudaGetLastError();
}

cudaError_t launch_shard_partition(st_tree *t, const int64_t *ids, int64_t m, int shift, int shards, const void *payload,
                                   int check_negative, uint8_t *shard_out, int32_t *slot_out, int64_t *part_ids_out,
                                   void *part_payload_out, void *scratch, unsigned int *row_out, cudaStream_t s)
{
    if (shards < 1 || shards > SH_MAX || shift < 0 || shift > 62) return cudaErrorInvalidValue;
    const int64_t ntiles = shard_tiles(m);
    if (ntiles <= 0 || ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    unsigned int *tile_counts = (unsigned int *)scratch;
    unsigned int *flags = tile_counts + (size_t)shards * ntiles;
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    g_kernel_launches += 3;
    ST_DISPATCH(t->dtype, {
        if (check_negative)
            shard_count_kernel<T, true><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, m, shard_out, tile_counts, ntiles, flags);
        else
            shard_count_kernel<T, false><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, m, shard_out, tile_counts, ntiles, flags);
        shard_scan_kernel<<<(unsigned)shards, SH_THREADS, 0, s>>>(tile_counts, ntiles, row_out, flags);
        shard_partition_kernel<T><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(ids, m, shift, shards, (const T *)payload, shard_out,
                                                                         tile_counts, ntiles, row_out, slot_out, part_ids_out,
                                                                         (T *)part_payload_out);
    });
    return cudaGetLastError();
}

cudaError_t launch_shard_expand_values(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                       const void *slabs, int64_t slab_stride, void *out, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    ++g_kernel_launches;
    const unsigned g = grid_for(m, 256, (int64_t)t->sm_count * 16);
    switch (t->esize) {
    case 1: shard_expand_values_kernel<uint8_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint8_t *)slabs, slab_stride, (uint8_t *)out); break;
    case 2: shard_expand_values_kernel<uint16_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint16_t *)slabs, slab_stride, (uint16_t *)out); break;
    case 4: shard_expand_values_kernel<uint32_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint32_t *)slabs, slab_stride, (uint32_t *)out); break;
    default: shard_expand_values_kernel<uint64_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint64_t *)slabs, slab_stride, (uint64_t *)out); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_shard_expand_cat(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m, const void *cat,
                                    const int64_t *starts_host, int shards, void *out, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    if (shards < 1 || shards > SH_MAX) return cudaErrorInvalidValue;
    shard_starts st{};
    for (int q = 0; q < shards; ++q) st.v[q] = starts_host[q];
    ++g_kernel_launches;
    const unsigned g = grid_for(m, 256, (int64_t)t->sm_count * 16);
    switch (t->esize) {
    case 1: shard_expand_cat_kernel<uint8_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint8_t *)cat, st, (uint8_t *)out); break;
    case 2: shard_expand_cat_kernel<uint16_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint16_t *)cat, st, (uint16_t *)out); break;
    case 4: shard_expand_cat_kernel<uint32_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint32_t *)cat, st, (uint32_t *)out); break;
    default: shard_expand_cat_kernel<uint64_t><<<g, 256, 0, s>>>(shard, slot, m, (const uint64_t *)cat, st, (uint64_t *)out); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_shard_expand_indices(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                        const int32_t *slabs, int64_t slab_stride, int64_t leaves_per_shard,
                                        int64_t *out, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    ++g_kernel_launches;
    shard_expand_indices_kernel<<<grid_for(m, 256, (int64_t)t->sm_count * 16), 256, 0, s>>>(
        shard, slot, m, slabs, slab_stride, leaves_per_shard, out);
    return cudaGetLastError();
}

}  // namespace st
