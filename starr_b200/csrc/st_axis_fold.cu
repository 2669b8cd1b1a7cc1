// st_axis_fold.cu -- sum over a NON-trailing axis of the leaves viewed as [A][R][C] (C-contiguous).
//
// Replaces the reference's fallback `return self.view(np.ndarray).sum(axis=..)` (sumtree_array.py:514)
// for the axis combinations the tree cannot answer (sumtree_array.py:497-512).  NumPy reduces a non-trailing
// axis row by row -- out[a][c] = (((x[a][0][c] + x[a][1][c]) + x[a][2][c]) + ...) -- so every output column is
// one SEQUENTIAL chain of R adds starting from +0; that order is kept for floats (bit-exact with ndarray.sum,
// tests/test_oracle.py::test_axis_fold_is_numpy_order).  Integer sums are associative (NumPy widens small
// integers to 64 bit; the 64-bit adds wrap), so their R range is additionally split over CTAs.
//
// The kernel is an HBM stream: s*A*R*C bytes in, s_out*A*C out.  Each CTA owns a strip of BC columns: a
// producer thread pulls the strip through a 6-stage ring of 16-KiB shared-memory tiles with TMA tensor
// copies (cp.async.bulk.tensor.3d + mbarrier complete_tx; SASS UTMALDG; full / empty barriers per slot), a
// consumer warp (lane = column) runs the dependent-add chains.  96 KiB per CTA, 192 KiB per SM in flight.
// Measured (tools/debug/axis_probe.cu, 4096 x 4096 float32): the tensor copies alone deliver 4.4 TB/s, the
// dependent-add chain (~9 cycles per row with the mbarrier wait) is what bounds a float column at ~20 us.
#include <cuda.h>

#include <mutex>

#include "st_common.cuh"

namespace st {

constexpr int AF_TILE_BYTES = 16384;   // one ring slot (a tile is BC columns x up to 256 rows)
constexpr int AF_STAGES = 6;          // 96 KiB per CTA, two CTAs per SM

// ---- generic fallback: one thread per output column (any alignment) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
axis_fold_kernel(const T *__restrict__ x, int64_t A, int64_t R, int64_t C,
                 typename fold_acc<T>::type *__restrict__ out)
{
    using ACC = typename fold_acc<T>::type;
    const int64_t total = A * C;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = o / C, c = o - a * C;
        const T *p = x + a * R * C + c;
        ACC acc = ACC(0);      // NumPy's add.reduce starts from the identity (+0)
        int64_t r = 0;
        constexpr int U = 16;
        for (; r + U <= R; r += U) {
            T v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) v[k] = __ldg(p + (r + k) * C);
#pragma unroll
            for (int k = 0; k < U; ++k) acc = t_add<ACC>(acc, (ACC)v[k]);
        }
        for (; r < R; ++r) acc = t_add<ACC>(acc, (ACC)__ldg(p + r * C));
        out[o] = acc;
    }
}

// ---- TMA-fed strip kernel ------------------------------------------------------------------------------
template <typename T, int BC> __host__ __device__ constexpr int af_rows()   // rows per tile (a box dimension is <= 256)
{
    return AF_TILE_BYTES / (BC * (int)sizeof(T)) > 256 ? 256 : AF_TILE_BYTES / (BC * (int)sizeof(T));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <typename T, int BC>
__global__ void __launch_bounds__(64)
axis_fold_tma_kernel(const __grid_constant__ CUtensorMap tm, int64_t R, int64_t C, int strips, int rsplit,
                     int64_t rows_per_part, typename fold_acc<T>::type *__restrict__ out)
{
    using ACC = typename fold_acc<T>::type;
    constexpr int BR = af_rows<T, BC>();
    constexpr int TILE = BR * BC * (int)sizeof(T);   // bytes one tensor copy delivers (<= AF_TILE_BYTES, the slot size)
    extern __shared__ __align__(128) unsigned char af_smem[];
    T *tiles = reinterpret_cast<T *>(af_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(af_smem + AF_STAGES * AF_TILE_BYTES);
    uint64_t *empty = full + AF_STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    int64_t b = blockIdx.x;
    const int part = (int)(b % rsplit); b /= rsplit;
    const int strip = (int)(b % strips);
    const int a = (int)(b / strips);
    const int64_t r0 = part * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < R ? r0 + rows_per_part : R;
    if (r0 >= r1) return;
    const int c0 = strip * BC;
    const int ntiles = (int)((r1 - r0 + BR - 1) / BR);

    if (threadIdx.x == 0) {
        for (int s = 0; s < AF_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(empty + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 1) {
        // ---- producer: one thread keeps the ring full (a slot is refilled when the consumer has released it)
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % AF_STAGES;
                if (t >= AF_STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(t / AF_STAGES - 1) & 1u);
                const uint32_t bar = smem_u32(full + s);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                    ::"r"(smem_u32(af_smem + (size_t)s * AF_TILE_BYTES)), "l"(&tm), "r"(c0), "r"((int)(r0 + (int64_t)t * BR)),
                      "r"(a), "r"(bar)
                    : "memory");
            }
        }
        return;
    }

    // ---- consumer warp: lane = column; the chain of adds runs straight out of the tile (the loads of the next
    // rows are issued between the dependent adds; staging blocks of rows in registers first measured slower),
    // the slot goes back to the producer after its last row.  Integer sums are associative: four chains per lane.
    const bool live = lane < BC && c0 + lane < C;
    const int col = lane < BC ? lane : 0;
    // NumPy's add.reduce starts from the identity: acc = +0, then row 0, 1, ... (an all -0.0 column sums to +0.0)
    ACC acc = ACC(0), acc1 = ACC(0), acc2 = ACC(0), acc3 = ACC(0);
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % AF_STAGES;
        mbar_wait(smem_u32(full + s), (uint32_t)(t / AF_STAGES) & 1u);
        const T *src = tiles + (size_t)s * (AF_TILE_BYTES / sizeof(T)) + col;
        const int64_t left = r1 - (r0 + (int64_t)t * BR);       // rows of this tile that exist
        if (left >= BR) {
            if constexpr (is_float_t<T>::value) {
#pragma unroll 32
                for (int k = 0; k < BR; ++k) acc = t_add<ACC>(acc, (ACC)src[k * BC]);
            } else {
                static_assert(BR % 4 == 0, "tile rows");
#pragma unroll 8
                for (int k = 0; k < BR; k += 4) {
                    acc += (ACC)src[k * BC];
                    acc1 += (ACC)src[(k + 1) * BC];
                    acc2 += (ACC)src[(k + 2) * BC];
                    acc3 += (ACC)src[(k + 3) * BC];
                }
            }
        } else {
            for (int k = 0; k < (int)left; ++k) acc = t_add<ACC>(acc, (ACC)src[k * BC]);
        }
        __syncwarp();   // every lane has read slot s
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
    }
    if constexpr (!is_float_t<T>::value) acc = (ACC)(acc + acc1 + acc2 + acc3);
    if (live) {
        ACC *o = out + (int64_t)a * C + c0 + lane;
        if (rsplit == 1) *o = acc;
        else if constexpr (!is_float_t<T>::value) atomicAdd(reinterpret_cast<unsigned long long *>(o), (unsigned long long)acc);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled()
{
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    });
    return fn;
}

template <typename T, int BC>
static cudaError_t launch_tma(const st_tree *t, const T *x, int64_t A, int64_t R, int64_t C, void *out, cudaStream_t s,
                              bool *done)
{
    using ACC = typename fold_acc<T>::type;
    constexpr int BR = af_rows<T, BC>();
    *done = false;
    encode_tiled_fn enc = encode_tiled();
    if (!enc) return cudaSuccess;
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)A};
    const cuuint64_t strides[2] = {(cuuint64_t)C * sizeof(T), (cuuint64_t)R * (cuuint64_t)C * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)BC, (cuuint32_t)BR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(T) == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                 : sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    if (enc(&tm, dt, 3, (void *)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            BC * sizeof(T) >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaSuccess;   // shape the descriptor cannot express: the generic kernel takes it
    const int strips = (int)((C + BC - 1) / BC);
    int rsplit = 1;
    if constexpr (!is_float_t<T>::value) {
        // associative: split the rows until every SM has a few CTAs (parts stay >= 4 tiles long)
        const int64_t want = (3LL * t->sm_count + A * strips - 1) / (A * strips);
        const int64_t most = R / (4 * BR);
        rsplit = (int)(want < 1 ? 1 : (want > most ? (most < 1 ? 1 : most) : want));
    }
    int64_t rows_per_part = (R + rsplit - 1) / rsplit;
    rows_per_part = (rows_per_part + BR - 1) / BR * BR;
    rsplit = (int)((R + rows_per_part - 1) / rows_per_part);
    const int64_t grid = A * strips * rsplit;
    if (grid > 0x7fffffffLL) return cudaSuccess;
    const size_t smem = (size_t)AF_STAGES * AF_TILE_BYTES + 2 * AF_STAGES * sizeof(uint64_t);
    auto kern = axis_fold_tma_kernel<T, BC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (rsplit > 1) {
        e = cudaMemsetAsync(out, 0, (size_t)(A * C) * sizeof(ACC), s);
        if (e != cudaSuccess) return e;
    }
    ++g_kernel_launches;
    kern<<<(unsigned)grid, 64, smem, s>>>(tm, R, C, strips, rsplit, rows_per_part, (ACC *)out);
    *done = true;
    return cudaGetLastError();
}

cudaError_t launch_axis_fold(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out, cudaStream_t s)
{
    if (A * C <= 0) return cudaSuccess;
    static const int force_bc = getenv("ST_AXIS_BC") ? atoi(getenv("ST_AXIS_BC")) : 0;   // tuning / tests: 0 auto, -1 generic
    ST_DISPATCH(t->dtype, {
        using ACC = typename fold_acc<T>::type;
        const T *x = (const T *)t->heap + t->n;
        const bool tma_ok = force_bc >= 0 && ((uintptr_t)x % 16 == 0) && ((C * (int64_t)sizeof(T)) % 16 == 0) &&
                            C < (1LL << 31) && R < (1LL << 31) && A < (1LL << 31) && R >= 64 &&
                            (C * (int64_t)sizeof(T)) < (1LL << 40) && (R * C * (int64_t)sizeof(T)) < (1LL << 40);
        if (tma_ok) {
            // widest strip that still gives every SM a CTA; a strip row is never narrower than one 32-byte sector
            int bc = 32;
            if constexpr (is_float_t<T>::value)   // (integer sums split the rows instead, see launch_tma)
                while (bc > 8 && bc / 2 * (int)sizeof(T) >= 32 && A * ((C + bc - 1) / bc) < t->sm_count) bc /= 2;
            if (force_bc == 32 || force_bc == 16 || force_bc == 8) bc = force_bc;
            if (bc * (int)sizeof(T) < 16) bc = 16;
            bool done = false;
            cudaError_t e = cudaSuccess;
            if (bc == 32) e = launch_tma<T, 32>(t, x, A, R, C, out, s, &done);
            else if (bc == 16) e = launch_tma<T, 16>(t, x, A, R, C, out, s, &done);
            else if constexpr (sizeof(T) >= 2) e = launch_tma<T, 8>(t, x, A, R, C, out, s, &done);
            if (e != cudaSuccess || done) return e;
        }
        ++g_kernel_launches;
        axis_fold_kernel<T><<<grid_for(A * C, 128, (int64_t)t->sm_count * 16), 128, 0, s>>>(x, A, R, C, (ACC *)out);
    });
    return cudaGetLastError();
}

}  // namespace st
