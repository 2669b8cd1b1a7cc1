// st_mintree.cu -- device-resident MIN tree.
//
// Replaces the pure-Python helper of the reference's PER buffer, starr/experimental/slow.py:53-78
// (MinTree: `_data[idx] = val`, then every ancestor = min(left, right), recomputed from the children)
// and is what CircularMinTree (slow.py:91-98) appends into.  Same heap layout as the sum tree:
// heap[h], 1 <= h < n internal, heap[n + i] = leaf i, children 2h / 2h+1, untouched slots = +inf
// (integers: the largest value of the type).
//
// Batched set, `heap[n + idx[j]] = vals[j % nv]` for j in batch order:
//   1. claim    atomicMax(tag[leaf], j): the LAST write to a leaf wins, as in a sequential loop;
//   2. write    the winner stores its value and releases the tag;
//   3. ancestors level by level from the deepest level up: every element recomputes its ancestor
//               at that level as min(left, right).  min is idempotent, so elements that share an
//               ancestor store the same value; a level starts after the level below is complete
//               (kernel boundary, or __syncthreads when one CTA holds the whole batch).
// Recomputing from the children is the reference's own formulation, so the result is the
// reference's for any batch order and any dtype (no rounding is involved).
#include "st_common.cuh"

namespace st {

template <typename T> __device__ __forceinline__ T min_ref(T l, T r) { return r < l ? r : l; }   // Python's min(l, r)

template <typename T> __host__ __device__ inline T min_identity()
{
    if constexpr (is_float_t<T>::value) {
#ifdef __CUDA_ARCH__
        return sizeof(T) == 4 ? (T)__int_as_float(0x7f800000) : (T)__longlong_as_double(0x7ff0000000000000LL);
#else
        return (T)__builtin_inf();
#endif
    } else if constexpr (is_signed_t<T>::value) {
        return (T)((~0ull >> 1) >> (64 - 8 * sizeof(T)));
    } else {
        return (T)~(T)0;
    }
}

template <typename T> __global__ void __launch_bounds__(256) min_fill_kernel(T *H, int64_t count)
{
    const T inf = min_identity<T>();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) H[i] = inf;
}

__global__ void __launch_bounds__(256) tag_fill_kernel(int *tag, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) tag[i] = -1;
}

__global__ void __launch_bounds__(256) min_claim_kernel(const int64_t *__restrict__ idx, int m, int64_t n, int *tag,
                                                        unsigned int *status)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t l = idx[j];
    if (l < 0 || l >= n) {
        atomicOr(status, STF_INDEX);
        return;
    }
    atomicMax(tag + l, j);
}

template <typename T>
__global__ void __launch_bounds__(256) min_write_kernel(T *__restrict__ H, const int64_t *__restrict__ idx, int m, int64_t n,
                                                        const T *__restrict__ vals, int64_t nv, int *tag,
                                                        const unsigned int *status)
{
    // a failed batch writes nothing, but the tags it claimed must still be released
    const bool failed = *status != 0;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t l = idx[j];
    if (l < 0 || l >= n) return;
    if (tag[l] != j) return;
    if (!failed) H[n + l] = vals[j < nv ? j : j % nv];
    tag[l] = -1;
}

// ancestor of heap leaf h at level l (root = level 0), or 0 if l is not above the leaf
__device__ __forceinline__ uint32_t ancestor_at(uint32_t h, int l)
{
    const int hl = heap_level(h);
    return l < hl ? h >> (hl - l) : 0u;
}

template <typename T>
__global__ void __launch_bounds__(256) min_level_kernel(T *__restrict__ H, const int64_t *__restrict__ idx, int m, int64_t n,
                                                        int level, const unsigned int *status)
{
    if (*status) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t p = ancestor_at((uint32_t)(idx[j] + n), level);
    if (p) H[p] = min_ref<T>(H[2 * p], H[2 * p + 1]);
}

// whole batch in one CTA (PER-sized batches): claim, write and every level separated by __syncthreads
constexpr int MIN_SMALL_THREADS = 1024;
constexpr int MIN_SMALL_E = 8;
template <typename T>
__global__ void __launch_bounds__(MIN_SMALL_THREADS) min_small_kernel(T *H, const int64_t *__restrict__ idx, int m, int64_t n,
                                                                      int depth, const T *__restrict__ vals, int64_t nv,
                                                                      int *tag, unsigned int *status)
{
    uint32_t h[MIN_SMALL_E];
    unsigned int bad = 0;
#pragma unroll
    for (int k = 0; k < MIN_SMALL_E; ++k) {
        const int j = threadIdx.x + k * MIN_SMALL_THREADS;
        h[k] = 0;
        if (j < m) {
            const int64_t l = idx[j];
            if (l < 0 || l >= n) bad = STF_INDEX;
            else h[k] = (uint32_t)(l + n);
        }
    }
    if (__syncthreads_or(bad != 0)) {
        if (threadIdx.x == 0) atomicOr(status, STF_INDEX);
        return;
    }
#pragma unroll
    for (int k = 0; k < MIN_SMALL_E; ++k)
        if (h[k]) atomicMax(tag + (h[k] - (uint32_t)n), threadIdx.x + k * MIN_SMALL_THREADS);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MIN_SMALL_E; ++k) {
        const int j = threadIdx.x + k * MIN_SMALL_THREADS;
        if (h[k] && tag[h[k] - (uint32_t)n] == j) {
            H[h[k]] = vals[j < nv ? j : j % nv];
            tag[h[k] - (uint32_t)n] = -1;
        }
    }
    for (int level = depth - 1; level >= 0; --level) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MIN_SMALL_E; ++k) {
            if (!h[k]) continue;
            const uint32_t p = ancestor_at(h[k], level);
            if (p) H[p] = min_ref<T>(H[2 * p], H[2 * p + 1]);
        }
    }
}

// full rebuild from the resident leaves: node = min(left, right) for h = n-1 .. 1, one level per launch
template <typename T>
__global__ void __launch_bounds__(256) min_build_level_kernel(T *__restrict__ H, uint32_t first, uint32_t last)
{
    for (uint64_t p = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < last; p += (uint64_t)gridDim.x * blockDim.x)
        H[p] = min_ref<T>(H[2 * p], H[2 * p + 1]);
}

cudaError_t launch_min_init(st_tree *t, cudaStream_t s)
{
    ST_DISPATCH(t->dtype, {
        g_kernel_launches += 2;
        min_fill_kernel<T><<<grid_for(2 * t->n, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>((T *)t->heap, 2 * t->n);
        tag_fill_kernel<<<grid_for(t->n, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>((int *)t->aux, t->n);
    });
    return cudaGetLastError();
}

cudaError_t launch_min_build(st_tree *t, cudaStream_t s)
{
    const uint32_t n = (uint32_t)t->n;
    ST_DISPATCH(t->dtype, {
        // levels of internal nodes, deepest first: [2^l, min(2^(l+1), n))
        for (int l = heap_level(n - 1); l >= 0; --l) {
            const uint32_t first = 1u << l;
            const uint32_t last = (2ull << l) < n ? (2u << l) : n;
            if (last <= first) continue;
            ++g_kernel_launches;
            min_build_level_kernel<T><<<grid_for(last - first, 256, (int64_t)t->sm_count * 8), 256, 0, s>>>((T *)t->heap, first, last);
        }
    });
    return cudaGetLastError();
}

cudaError_t launch_min_update(st_tree *t, const int64_t *idx, int64_t ni, const void *vals, int64_t nv, cudaStream_t s)
{
    if (ni <= 0) return cudaSuccess;
    if (ni > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int m = (int)ni;
    cudaError_t e = launch_batch_begin(t, s);
    if (e != cudaSuccess) return e;
    ST_DISPATCH(t->dtype, {
        T *H = (T *)t->heap;
        int *tag = (int *)t->aux;
        if (m <= MIN_SMALL_THREADS * MIN_SMALL_E) {
            ++g_kernel_launches;
            min_small_kernel<T><<<1, MIN_SMALL_THREADS, 0, s>>>(H, idx, m, t->n, t->depth, (const T *)vals, nv, tag, t->d_status);
        } else {
            const unsigned g = (unsigned)((m + 255) / 256);
            g_kernel_launches += 2 + t->depth;
            min_claim_kernel<<<g, 256, 0, s>>>(idx, m, t->n, tag, t->d_status);
            min_write_kernel<T><<<g, 256, 0, s>>>(H, idx, m, t->n, (const T *)vals, nv, tag, t->d_status);
            for (int level = t->depth - 1; level >= 0; --level)
                min_level_kernel<T><<<g, 256, 0, s>>>(H, idx, m, t->n, level, t->d_status);
        }
    });
    return cudaGetLastError();
}

}  // namespace st
