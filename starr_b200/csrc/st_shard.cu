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

template <typename T, bool CHECK_NEG>
__global__ void __launch_bounds__(SH_THREADS) shard_count_kernel(const int64_t *__restrict__ ids, int64_t m, int shift,
                                                                 int shards, const T *__restrict__ payload,
                                                                 int64_t npayload, uint8_t *__restrict__ shard_out,
                                                                 unsigned int *__restrict__ tile_counts, int64_t ntiles,
                                                                 unsigned int *status_a, unsigned int *status_b)
{
    __shared__ unsigned int hist[SH_MAX];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < SH_MAX) hist[tid] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SH_TILE;
    unsigned int bad = 0;
#pragma unroll
    for (int r = 0; r < SH_E; ++r) {
        const int64_t j = base + r * SH_THREADS + tid;
        int key = 0x100 + lane;   // unique: matches nobody
        if (j < m) {
            const int64_t id = ids[j];
            const int64_t q = id >> shift;
            if (id < 0 || q >= shards) {
                bad |= STF_INDEX;
                shard_out[j] = SH_NONE;
            } else {
                shard_out[j] = (uint8_t)q;
                key = (int)q;
            }
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, key);
        if (key < 0x100 && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&hist[key], __popc(peers));
    }
    if constexpr (CHECK_NEG && is_signed_t<T>::value) {
        // the reference rejects a negative entry anywhere in vals, used or not (pyx:37-39)
        const int64_t stride = (int64_t)gridDim.x * SH_THREADS;
        for (int64_t k = (int64_t)blockIdx.x * SH_THREADS + tid; k < npayload; k += stride)
            if (payload[k] < T(0)) bad |= STF_NEG_VALS;
    }
    if (bad) {
        atomicOr(status_a + SW_STATUS, bad);
        if (status_b) atomicOr(status_b + SW_STATUS, bad);
    }
    __syncthreads();
    if (tid < shards) tile_counts[(int64_t)tid * ntiles + blockIdx.x] = hist[tid];
}

// one CTA per shard: exclusive scan of its per-tile counts, total -> counts[shard]
__global__ void __launch_bounds__(SH_THREADS) shard_scan_kernel(unsigned int *__restrict__ tile_counts, int64_t ntiles,
                                                                unsigned int *__restrict__ counts)
{
    typedef cub::BlockScan<unsigned int, SH_THREADS> scan_t;
    __shared__ typename scan_t::TempStorage tmp;
    unsigned int *row = tile_counts + (int64_t)blockIdx.x * ntiles;
    const int64_t per = (ntiles + SH_THREADS - 1) / SH_THREADS;
    const int64_t lo = per * threadIdx.x;
    const int64_t hi = lo + per < ntiles ? lo + per : ntiles;
    unsigned int sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += row[i];
    unsigned int off, total;
    scan_t(tmp).ExclusiveSum(sum, off, total);
    for (int64_t i = lo; i < hi; ++i) {
        const unsigned int c = row[i];
        row[i] = off;
        off += c;
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(SH_THREADS) shard_slot_kernel(const int64_t *__restrict__ ids, int64_t m, int shift,
                                                                int shards, int rank, const T *__restrict__ payload,
                                                                int64_t npayload, const uint8_t *__restrict__ shard8,
                                                                const unsigned int *__restrict__ tile_off,
                                                                int64_t ntiles, int32_t *__restrict__ slot_out,
                                                                int64_t *__restrict__ own_ids, T *__restrict__ own_payload)
{
    __shared__ unsigned int s_base[SH_MAX];           // elements of each shard before the current row
    __shared__ unsigned int s_wc[SH_WARPS][SH_MAX];   // per-warp counts of the current row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < SH_MAX) s_base[tid] = tid < shards ? tile_off[(int64_t)tid * ntiles + blockIdx.x] : 0u;
    for (int i = tid; i < SH_WARPS * SH_MAX; i += SH_THREADS) (&s_wc[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SH_TILE;
    const int64_t origin = (int64_t)rank << shift;
    for (int r = 0; r < SH_E; ++r) {
        const int64_t j = base + r * SH_THREADS + tid;
        const int sh = j < m ? (int)shard8[j] : (int)SH_NONE;
        const bool valid = sh < shards;
        const unsigned int peers = __match_any_sync(0xffffffffu, valid ? sh : 0x100 + lane);
        const int within = __popc(peers & ((1u << lane) - 1));
        if (valid && within == 0) s_wc[warp][sh] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned int before = s_base[sh];
            for (int w = 0; w < warp; ++w) before += s_wc[w][sh];
            const unsigned int slot = before + within;
            slot_out[j] = (int32_t)slot;
            if (sh == rank) {
                if (own_ids) own_ids[slot] = ids[j] - origin;
                if (own_payload) own_payload[slot] = payload[j < npayload ? j : j % npayload];
            }
        } else if (j < m) {
            slot_out[j] = -1;
        }
        __syncthreads();
        if (tid < shards) {
            unsigned int sum = 0;
#pragma unroll
            for (int w = 0; w < SH_WARPS; ++w) {
                sum += s_wc[w][tid];
                s_wc[w][tid] = 0;
            }
            s_base[tid] += sum;
        }
        __syncthreads();
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

size_t shard_plan_scratch_bytes(int64_t m, int shards)
{
    return (size_t)shards * (size_t)(shard_tiles(m) > 0 ? shard_tiles(m) : 1) * sizeof(unsigned int);
}

cudaError_t launch_shard_plan(st_tree *local, st_tree *top, const int64_t *ids, int64_t m, int shift, int shards,
                              int rank, const void *payload, int64_t npayload, int check_negative, uint8_t *shard_out,
                              int32_t *slot_out, int64_t *own_ids_out, void *own_payload_out, void *scratch,
                              unsigned int *counts_out, cudaStream_t s)
{
    if (shards < 1 || shards > SH_MAX || rank < 0 || rank >= shards || shift < 0 || shift > 62)
        return cudaErrorInvalidValue;
    const int64_t ntiles = shard_tiles(m);
    if (ntiles <= 0 || ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    unsigned int *tile_counts = (unsigned int *)scratch;
    unsigned int *status_b = top ? top->d_status : nullptr;
    g_kernel_launches += 3;
    ST_DISPATCH(local->dtype, {
        if (check_negative)
            shard_count_kernel<T, true><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, npayload, shard_out, tile_counts, ntiles, local->d_status,
                status_b);
        else
            shard_count_kernel<T, false><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(
                ids, m, shift, shards, (const T *)payload, npayload, shard_out, tile_counts, ntiles, local->d_status,
                status_b);
        shard_scan_kernel<<<(unsigned)shards, SH_THREADS, 0, s>>>(tile_counts, ntiles, counts_out);
        shard_slot_kernel<T><<<(unsigned)ntiles, SH_THREADS, 0, s>>>(ids, m, shift, shards, rank, (const T *)payload,
                                                                    npayload, shard_out, tile_counts, ntiles, slot_out,
                                                                    own_ids_out, (T *)own_payload_out);
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
