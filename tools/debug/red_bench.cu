// Microbenchmark (not product code): how fast can 2^20 random leaves push one add into each of their K
// deepest ancestors of a 2^27-leaf heap?  Answers what the retire phase of the float update can cost at best.
//   A: lane = element, loop over levels, fire-and-forget RED.ADD.F32
//   B: lane = element, loop over levels, plain load + add + store (valid only for single-update nodes)
//   C: lane = level, elements broadcast by shuffle (the r1 retire_kernel mapping)
//   D: like A but only levels [lo, hi) -- to separate L2-resident levels from DRAM-resident ones
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <random>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void kA(float *H, const unsigned *leaf, const float *d, int m, int n_log, int lo, int hi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    unsigned h = leaf[i] + (1u << n_log);
    float v = d[i];
    for (int s = lo; s < hi; ++s) atomicAdd(H + (h >> s), v);
}
__global__ void kB(float *H, const unsigned *leaf, const float *d, int m, int n_log, int lo, int hi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    unsigned h = leaf[i] + (1u << n_log);
    float v = d[i];
    float t[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) if (lo + s < hi) t[s] = __ldcg(H + (h >> (lo + s)));
#pragma unroll
    for (int s = 0; s < 16; ++s) if (lo + s < hi) __stcg(H + (h >> (lo + s)), t[s] + v);
}
__global__ void kC(float *H, const unsigned *leaf, const float *d, int m, int n_log, int lo, int hi, int seg)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nw = (gridDim.x * blockDim.x) >> 5;
    int nseg = (m + seg - 1) / seg;
    for (int w = warp; w < nseg; w += nw) {
        int base0 = w * seg;
        int len = min(seg, m - base0);
        int s = lo + lane;
        for (int base = 0; base < len; base += 32) {
            unsigned l = 0; float v = 0.f;
            if (base + lane < len) { l = leaf[base0 + base + lane]; v = d[base0 + base + lane]; }
            int cnt = min(32, len - base);
            for (int x = 0; x < cnt; ++x) {
                unsigned lx = __shfl_sync(0xffffffffu, l, x);
                float vx = __shfl_sync(0xffffffffu, v, x);
                if (s < hi) atomicAdd(H + ((lx + (1u << n_log)) >> s), vx);
            }
        }
    }
}

__global__ void kPre(const float *H, const unsigned *leaf, int m, int n_log, int lo, int hi, float *sink)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    unsigned h = leaf[i] + (1u << n_log);
    for (int s = lo; s < hi; ++s) asm volatile("prefetch.global.L2 [%0];" ::"l"(H + (h >> s)));
}
// prefetch for the elements `ahead` positions later, then RED own elements
__global__ void kPipe(float *H, const unsigned *leaf, const float *d, int m, int n_log, int lo, int hi, int ahead)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + ahead < m) {
        unsigned hp = leaf[i + ahead] + (1u << n_log);
        for (int s = lo; s < hi; ++s) asm volatile("prefetch.global.L2 [%0];" ::"l"(H + (hp >> s)));
    }
    if (i >= m) return;
    unsigned h = leaf[i] + (1u << n_log);
    float v = d[i];
    for (int s = lo; s < hi; ++s) atomicAdd(H + (h >> s), v);
}

int main(int argc, char **argv)
{
    const int n_log = 27, m = 1 << 20;
    float *H; unsigned *leaf; float *d;
    CK(cudaMalloc(&H, sizeof(float) * (2ull << n_log)));
    CK(cudaMemset(H, 0, sizeof(float) * (2ull << n_log)));
    std::vector<unsigned> hl(m); std::vector<float> hd(m, 1.0f);
    std::mt19937 rng(1);
    for (auto &x : hl) x = rng() & ((1u << n_log) - 1);
    CK(cudaMalloc(&leaf, 4 * m)); CK(cudaMalloc(&d, 4 * m));
    // a big buffer to flush L2 between runs
    char *flush; CK(cudaMalloc(&flush, 256 << 20));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto fn) {
        float best = 1e9, sum = 0;
        for (int r = 0; r < 6; ++r) {
            for (auto &x : hl) x = rng() & ((1u << n_log) - 1);
            CK(cudaMemcpy(leaf, hl.data(), 4 * m, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d, hd.data(), 4 * m, cudaMemcpyHostToDevice));
            CK(cudaMemsetAsync(flush, r, 256 << 20));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a); fn(); cudaEventRecord(b); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (r) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-58s best %7.1f us  mean %7.1f us\n", name, best * 1e3, sum / 5 * 1e3);
    };
    const int g = (m + 255) / 256;
    run("A RED lane=element, levels 1..16 (depth 26..11)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 17); });
    run("A RED lane=element, levels 0..16 (incl. leaf)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 17); });
    run("A RED lane=element, deep only s=1..5 (DRAM levels 26..22)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 6); });
    run("A RED lane=element, s=6..16 (L2 levels 21..11)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 6, 17); });
    run("A RED lane=element, s=1 only", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 2); });
    run("A RED lane=element, s=12 only (level 15, 128 KiB)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 12, 13); });
    run("A RED lane=element, s=1..27 (all levels incl. root!)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 28); });
    run("B ld+st lane=element, s=1..16", [&] { kB<<<g, 256>>>(H, leaf, d, m, n_log, 1, 17); });
    run("B ld+st lane=element, deep only s=1..5", [&] { kB<<<g, 256>>>(H, leaf, d, m, n_log, 1, 6); });
    for (int seg : {512, 128, 32})
        for (int blocks : {1184, 2368}) {
            char nm[128]; snprintf(nm, 128, "C RED lane=level (retire mapping) seg=%d blocks=%d", seg, blocks);
            run(nm, [&] { kC<<<blocks, 256>>>(H, leaf, d, m, n_log, 1, 17, seg); });
        }
    // the same with the leaves SORTED (what a pass in leaf order sees): DRAM page locality, partial coalescing
    auto run_sorted = [&](const char *name, auto fn) {
        float best = 1e9, sum = 0;
        for (int r = 0; r < 6; ++r) {
            for (auto &x : hl) x = rng() & ((1u << n_log) - 1);
            std::sort(hl.begin(), hl.end());
            CK(cudaMemcpy(leaf, hl.data(), 4 * m, cudaMemcpyHostToDevice));
            CK(cudaMemsetAsync(flush, r, 256 << 20));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a); fn(); cudaEventRecord(b); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (r) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-58s best %7.1f us  mean %7.1f us\n", name, best * 1e3, sum / 5 * 1e3);
    };
    run_sorted("SORTED A RED s=0..0 (leaf only)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 1); });
    run_sorted("SORTED A RED s=1..5 (levels 26..22)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 6); });
    run_sorted("SORTED A RED s=0..6 (leaf + levels 26..21)", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7); });
    run_sorted("SORTED A RED s=0..8", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 9); });
    run_sorted("SORTED A RED s=1..16", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 1, 17); });
    run_sorted("SORTED A RED s=6..16", [&] { kA<<<g, 256>>>(H, leaf, d, m, n_log, 6, 17); });
    run_sorted("SORTED B ld+st s=0..6", [&] { kB<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7); });
    run_sorted("SORTED prefetch kernel alone s=0..6", [&] { kPre<<<g, 256>>>(H, leaf, m, n_log, 0, 7, nullptr); });
    run_sorted("SORTED prefetch kernel + RED kernel s=0..6", [&] { kPre<<<g, 256>>>(H, leaf, m, n_log, 0, 7, nullptr); kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7); });
    for (int ahead : {1 << 14, 1 << 16, 1 << 17, 1 << 18}) {
        char nm[128]; snprintf(nm, 128, "SORTED pipelined prefetch ahead=%d + RED s=0..6", ahead);
        run_sorted(nm, [&] { kPipe<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7, ahead); });
    }
    run("RANDOM prefetch kernel + RED kernel s=0..6", [&] { kPre<<<g, 256>>>(H, leaf, m, n_log, 0, 7, nullptr); kA<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7); });
    run("RANDOM pipelined prefetch ahead=65536 + RED s=0..6", [&] { kPipe<<<g, 256>>>(H, leaf, d, m, n_log, 0, 7, 1 << 16); });
    run("RANDOM pipelined prefetch ahead=65536 + RED s=0..16", [&] { kPipe<<<g, 256>>>(H, leaf, d, m, n_log, 0, 17, 1 << 16); });
    return 0;
}
