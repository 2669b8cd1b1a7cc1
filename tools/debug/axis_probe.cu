// axis_probe.cu -- where does the strip kernel of st_axis_fold.cu lose its time?  Variants of the same
// column-strip fold over a 4096 x 4096 float matrix: TMA tensor tiles vs LDGSTS (cp.async 16 B) producers,
// with and without the add chain, different tile sizes / stage counts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/debug/axis_probe tools/debug/axis_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// ---------------- TMA variant: BC columns, BR rows per tile, STAGES slots, consumer optional ----------------
template <int BC, int BR, int STAGES, bool ADD>
__global__ void __launch_bounds__(32 + (BC > 32 ? BC : 32)) tma_kernel(const __grid_constant__ CUtensorMap tm, int R, int C, float *out)
{
    constexpr int TILE = BC * BR * 4;
    constexpr int NCW = BC > 32 ? BC / 32 : 1;
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = (uint64_t *)(sm + STAGES * TILE), *empty = full + STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * BC;
    const int ntiles = R / BR;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + s)), "r"(NCW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NCW) {
        if (lane == 0)
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % STAGES;
                if (t >= STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(t / STAGES - 1) & 1u);
                const uint32_t bar = smem_u32(full + s);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(sm + (size_t)s * TILE)), "l"(&tm), "r"(c0), "r"(t * BR), "r"(bar) : "memory");
            }
        return;
    }
    float acc = 0.f;
    const float *tiles = (const float *)sm;
    const int col = (warp * 32 + lane) % BC;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(smem_u32(full + s), (uint32_t)(t / STAGES) & 1u);
        const float *src = tiles + (size_t)s * (TILE / 4) + col;
        if (ADD) {
#pragma unroll 32
            for (int k = 0; k < BR; ++k) acc = __fadd_rn(acc, src[k * BC]);
        } else {
            acc += src[0];
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
    }
    if (warp * 32 + lane < BC) out[c0 + warp * 32 + lane] = acc;
}

// ---------------- LDGSTS variant: PW producer warps copy 16 B per lane, mbarrier-tracked ----------------
template <int BC, int BR, int STAGES, int PW, bool ADD>
__global__ void __launch_bounds__(32 + 32 * PW) ldgsts_kernel(const float *__restrict__ x, int R, int C, float *out)
{
    constexpr int TILE = BC * BR * 4;
    constexpr int VEC_PER_ROW = BC / 4;                 // 16-byte pieces per tile row
    constexpr int VECS = BR * VEC_PER_ROW;              // per tile
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = (uint64_t *)(sm + STAGES * TILE), *empty = full + STAGES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * BC;
    const int ntiles = R / BR;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(32 * PW));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(empty + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= 1) {
        const int pt = (warp - 1) * 32 + lane;          // producer thread id
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % STAGES;
            if (t >= STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(t / STAGES - 1) & 1u);
            for (int v = pt; v < VECS; v += 32 * PW) {
                const int r = v / VEC_PER_ROW, cv = v % VEC_PER_ROW;
                const float *g = x + (size_t)(t * BR + r) * C + c0 + cv * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + (size_t)s * TILE + (size_t)v * 16)), "l"(g) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + s)) : "memory");
        }
        return;
    }
    float acc = 0.f;
    const float *tiles = (const float *)sm;
    const int col = lane % BC;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(smem_u32(full + s), (uint32_t)(t / STAGES) & 1u);
        const float *src = tiles + (size_t)s * (TILE / 4) + col;
        if (ADD) {
#pragma unroll 32
            for (int k = 0; k < BR; ++k) acc = __fadd_rn(acc, src[k * BC]);
        } else {
            acc += src[0];
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
    }
    if (lane < BC) out[c0 + lane] = acc;
}

static float *g_flush; static size_t g_flush_n = 256u << 20;
static const float *g_x[4];
template <typename F> static float timeit(F f)   // f(which matrix): 8 launches cycling over 4 x 64 MiB (> L2) per timing
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<float> v;
    for (int i = 0; i < 5; ++i) {
        CK(cudaMemset(g_flush, i, g_flush_n));
        cudaEventRecord(a); for (int q = 0; q < 8; ++q) f(q & 3); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); v.push_back(ms * 1e3f / 8);
    }
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

template <int BC, int BR, int STAGES, bool ADD> static void run_tma(const float *x, int R, int C, float *out, const std::vector<float> &want, int promo = 0)
{
    CUtensorMap tm[4];
    for (int q = 0; q < 4; ++q) {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {BC, BR}; cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm[q], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)g_x[q], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    }
    const size_t smem = (size_t)STAGES * BC * BR * 4 + 2 * STAGES * 8;
    auto k = tma_kernel<BC, BR, STAGES, ADD>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    printf("promo=%3d ", promo);
    const float us = timeit([&](int q) { k<<<C / BC, 32 + (BC > 32 ? BC : 32), smem>>>(tm[q], R, C, out); });
    CK(cudaGetLastError());
    std::vector<float> got(C); CK(cudaMemcpy(got.data(), out, C * 4, cudaMemcpyDeviceToHost));
    bool ok = !ADD || std::equal(got.begin(), got.end(), want.begin());
    printf("TMA    BC=%2d BR=%3d stages=%2d (%3d KB/CTA) add=%d  %7.1f us  %6.0f GB/s  ok=%d\n", BC, BR, STAGES, (int)(smem >> 10), (int)ADD, us,
           (double)R * C * 4 / us / 1e3, (int)ok);
}

template <int BC, int BR, int STAGES, int PW, bool ADD> static void run_ldgsts(const float *x, int R, int C, float *out, const std::vector<float> &want)
{
    const size_t smem = (size_t)STAGES * BC * BR * 4 + 2 * STAGES * 8;
    auto k = ldgsts_kernel<BC, BR, STAGES, PW, ADD>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float us = timeit([&](int q) { k<<<C / BC, 32 + 32 * PW, smem>>>(g_x[q], R, C, out); });
    CK(cudaGetLastError());
    std::vector<float> got(C); CK(cudaMemcpy(got.data(), out, C * 4, cudaMemcpyDeviceToHost));
    bool ok = !ADD || std::equal(got.begin(), got.end(), want.begin());
    printf("LDGSTS BC=%2d BR=%3d stages=%2d pw=%d (%3d KB/CTA) add=%d  %7.1f us  %6.0f GB/s  ok=%d\n", BC, BR, STAGES, PW, (int)(smem >> 10), (int)ADD,
           us, (double)R * C * 4 / us / 1e3, (int)ok);
}

int main()
{
    const int R = 4096, C = 4096;
    CK(cudaFree(0));
    std::vector<float> h((size_t)R * C);
    srand(1);
    for (auto &v : h) v = (float)rand() / RAND_MAX;
    std::vector<float> want(C, 0.f);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) want[c] = want[c] + h[(size_t)r * C + c];
    float *x, *out;
    CK(cudaMalloc(&x, h.size() * 4)); CK(cudaMalloc(&out, C * 4)); CK(cudaMalloc(&g_flush, g_flush_n));
    CK(cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    for (int q = 0; q < 4; ++q) {
        float *y; CK(cudaMalloc(&y, h.size() * 4)); CK(cudaMemcpy(y, x, h.size() * 4, cudaMemcpyDeviceToDevice)); g_x[q] = y;
    }
    for (int promo : {0, 128, 256}) {
        run_tma<32, 32, 16, true>(x, R, C, out, want, promo);
        run_tma<32, 128, 8, true>(x, R, C, out, want, promo);
        run_tma<32, 128, 8, false>(x, R, C, out, want, promo);
        run_tma<16, 64, 16, true>(x, R, C, out, want, promo);
        run_tma<16, 256, 6, true>(x, R, C, out, want, promo);
        run_tma<64, 32, 16, true>(x, R, C, out, want, promo);
        run_tma<64, 64, 12, true>(x, R, C, out, want, promo);
        run_tma<64, 64, 12, false>(x, R, C, out, want, promo);
        run_tma<128, 32, 12, true>(x, R, C, out, want, promo);
    }
    run_ldgsts<32, 32, 16, 1, true>(x, R, C, out, want);
    run_ldgsts<32, 32, 48, 4, true>(x, R, C, out, want);
    run_ldgsts<16, 64, 16, 2, true>(x, R, C, out, want);
    return 0;
}
