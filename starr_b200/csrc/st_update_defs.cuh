// st_update_defs.cuh -- small shared definitions of the update pipeline.
#pragma once
#include "st_common.cuh"

namespace st {

constexpr int64_t UPDATE_MAX_CHUNK = 1 << 21;
constexpr int SHORT_SEG = 32;    // segments up to this length are folded by their head thread
#ifndef ST_RETIRE_SEG
#define ST_RETIRE_SEG 512
#endif
constexpr int RETIRE_SEG = ST_RETIRE_SEG;  // multi-CTA path: segments up to this length are retired with ordered atomics
constexpr int MEDIUM_SEG = 2048;  // up to this length by one warp, longer ones by one CTA
#ifndef ST_HUGE_SEG
#define ST_HUGE_SEG 16384
#endif
constexpr int HUGE_SEG = ST_HUGE_SEG;   // longer than this: all CTAs summarise chunks, one CTA stitches them

struct seg_entry {
    int level;
    int start;
    int len;
    uint32_t node;
};

// status-word layout (unsigned int[64])
enum { SW_STATUS = 0, SW_QCOUNT = 1, SW_QTAKE = 2, SW_QCOUNT_MED = 8, SW_QTAKE_MED = 9, SW_QCOUNT_HUGE = 10, SW_HUGE_CHUNKS = 11, SW_QTAKE_HUGE = 12, SW_REDO_ANY = 13, SW_ANYLONG = 14 /* and 15 */, SW_RETIRE_MASK = 16, SW_QCOUNT_RET = 19, SW_STAT_FAST = 4, SW_STAT_CHAINS = 5, SW_STAT_CHAIN_ELEMS = 6, SW_STAT_LONGEST = 7, SW_STAT_ZERO_DELTA = 20 /* updates whose difference was exactly 0 (cumulative) */,
       SW_NEED_COOP = 23 /* hybrid partition: a level-LT segment is too long for one CTA (21 = ST_SW_STICKY) */, SW_HY_TICKET = 24,
       SW_ANYBIG = 26 /* the leaf pass flagged a group in bigmap */,
       SW_HY_KEYLEVELS = 25 /* bit l: lev_key[l] of a top level l is read (a segment is retired there) */ };

// -----------------------------------------------------------------------------------------
// key helpers.  key = heap index of the leaf, shifted left so that its top bit sits at bit
// `depth`; sorting by key is the tree's left-to-right leaf order (the rotated order of
// non-power-of-two trees).  Ancestor at level l (root = 0) is key >> (depth - l),
// valid for l < leaf level.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_to_heap(uint32_t key, uint64_t two_n) { return key < two_n ? key : key >> 1; }
__device__ __forceinline__ int key_leaf_level(uint32_t key, uint64_t two_n, int depth) { return key < two_n ? depth : depth - 1; }


// One ordered add into a tree node: same-address atomics of ONE thread are performed in program
// order and each is a single IEEE round-to-nearest add at the L2, so a thread that issues its adds
// in batch order reproduces (((T + d0) + d1) + ...) exactly without ever waiting for T.
// float32 global atomic adds flush subnormals (PTX atom.add.f32 / SASS REDG...FTZ): they are used
// only for |d| >= 2^-100, where a subnormal T is absorbed identically and no subnormal result can
// arise (both operands are then multiples of >= 2^-124); smaller non-zero d take an atomic CAS
// (still ordered with the other atomics of this thread); d == 0 changes nothing and is skipped.
template <typename T> __device__ __forceinline__ void ordered_add(T *ptr, T d)
{
    if (d == T(0)) return;
    if constexpr (sizeof(T) == 4) {
        if (fabsf(d) < 7.888609052210118e-31f) {  // 2^-100
            unsigned int *u = reinterpret_cast<unsigned int *>(ptr);
            unsigned int old = atomicOr(u, 0u), assumed;
            do {
                assumed = old;
                old = atomicCAS(u, assumed, __float_as_uint(__fadd_rn(__uint_as_float(assumed), d)));
            } while (old != assumed);
            return;
        }
    }
    atomicAdd(ptr, d);
}

constexpr int SEG_DONE = 1 << 30;

// small-batch path: the single-CTA kernel queues its retirements for the follow-up kernel
template <typename T> struct retire_params {
    const seg_entry *qret;     // nullptr: nothing to retire in this launch
    const uint32_t *key0;      // level-0 keys (batch order)
    const uint32_t *lev_key;   // [level][lev_stride]
    int depth;
    uint32_t n;
};

}  // namespace st
