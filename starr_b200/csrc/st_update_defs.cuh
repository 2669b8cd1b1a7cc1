// st_update_defs.cuh -- small shared definitions of the update pipeline.
#pragma once
#include "st_common.cuh"

namespace st {

constexpr int64_t UPDATE_MAX_CHUNK = 1 << 20;
constexpr int SHORT_SEG = 32;    // segments up to this length are folded by their head thread
constexpr int MEDIUM_SEG = 2048;  // up to this length by one warp, longer ones by one CTA
constexpr int HUGE_SEG = 16384;   // longer than this: all CTAs summarise chunks, one CTA stitches them

struct seg_entry {
    int level;
    int start;
    int len;
    uint32_t node;
};

// status-word layout (unsigned int[64])
enum { SW_STATUS = 0, SW_QCOUNT = 1, SW_QTAKE = 2, SW_QCOUNT_MED = 8, SW_QTAKE_MED = 9, SW_QCOUNT_HUGE = 10, SW_HUGE_CHUNKS = 11, SW_QTAKE_HUGE = 12, SW_STAT_FAST = 4, SW_STAT_CHAINS = 5, SW_STAT_CHAIN_ELEMS = 6, SW_STAT_LONGEST = 7 };

// -----------------------------------------------------------------------------------------
// key helpers.  key = heap index of the leaf, shifted left so that its top bit sits at bit
// `depth`; sorting by key is the tree's left-to-right leaf order (the rotated order of
// non-power-of-two trees).  Ancestor at level l (root = 0) is key >> (depth - l),
// valid for l < leaf level.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t key_to_heap(uint32_t key, uint64_t two_n) { return key < two_n ? key : key >> 1; }
__device__ __forceinline__ int key_leaf_level(uint32_t key, uint64_t two_n, int depth) { return key < two_n ? depth : depth - 1; }

}  // namespace st
