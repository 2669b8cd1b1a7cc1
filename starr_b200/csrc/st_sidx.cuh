// st_sidx.cuh -- sector-packed search index: the left-child values a prefix-sum descent needs,
// regrouped so that THREE levels of the descent (get_prefix_sum_idx, starr/src/_sumtree_array.pyx:96-113) cost
// one 32-byte sector instead of three.
//
// The descent only ever reads LEFT children (pyx:104-110: compare with the left child, subtract it when going
// right).  For a block rooted at heap node r of level L, the three steps below r need
//     H[2r] | H[4r], H[4r+2] | H[8r], H[8r+2], H[8r+4], H[8r+6]         (7 values)
// which the tree stores in 7 different sectors.  The index stores them as ONE 8-element record (element 0 unused)
// per node of the levels L0, L0+3, L0+6, ...: L0 = 13 is where the shared-memory staging of search_kernel ends.
// 4-byte element types only (7 x 4 B fit one sector); covered child levels L0+1 .. L0+3*nb <= depth-2 (all
// internal nodes).  The copies are kept bit-identical by every writer of tree nodes: each one applies the same
// operation, in the same order, to the node and to its copy (sidx_slot below).  Size: ~1/7 of the tree.
#pragma once
#include "st_common.cuh"

namespace st {

constexpr int SIDX_L0 = 13;
constexpr int SIDX_MAX_BLOCK_LEVELS = 7;

struct sidx_ref {
    void *base = nullptr;                      // nullptr: no index (plain descent, writers skip the copies)
    int nb = 0;                                // block levels L0, L0+3, ..., L0+3(nb-1)
    uint32_t off[SIDX_MAX_BLOCK_LEVELS] = {};  // first element of every block level
};

// copy of heap node v (level l) inside the index, or nullptr if v has none (right child / level not covered)
template <typename T> __device__ __forceinline__ T *sidx_slot(const sidx_ref &sx, uint32_t v, int l)
{
    if (!sx.base || (v & 1u)) return nullptr;
    const int rel = l - SIDX_L0 - 1;
    if (rel < 0 || rel >= 3 * sx.nb) return nullptr;
    const int i = rel / 3, k = rel - 3 * i + 1;          // block level index, depth below the block root (1..3)
    const uint32_t r = v >> k;
    const uint32_t p = (v & ((1u << k) - 1u)) >> 1;
    const uint32_t blk = r - (1u << (SIDX_L0 + 3 * i));
    return reinterpret_cast<T *>(sx.base) + sx.off[i] + blk * 8u + (1u << (k - 1)) + p;
}

// node store / ordered add that keep the index copy identical
template <typename T> __device__ __forceinline__ void sidx_store(T *H, const sidx_ref &sx, uint32_t v, int l, T x)
{
    H[v] = x;
    if (T *c = sidx_slot<T>(sx, v, l)) *c = x;
}

sidx_ref sidx_of(const st_tree *t);   // st_tree_kernels.cu

}  // namespace st
