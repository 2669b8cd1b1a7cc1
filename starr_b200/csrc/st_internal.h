// st_internal.h -- shared between the translation units of libstarr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/starr_b200.h"

// Device-side status word bits (written by validation kernels, read by every later kernel
// of the same batch: a failed batch is skipped as a whole, see DESIGN.md "error atomicity").
#define STF_NEG_VALS 1u
#define STF_NEG_ARRAY 2u
#define STF_INDEX 4u

struct st_workspace {
    void *base = nullptr;  // one cudaMalloc, carved up per batch
    size_t bytes = 0;
};

struct st_tree {
    int64_t n = 0;
    int dtype = 0;
    int device = 0;
    int esize = 0;
    int depth = 0;        // level of the deepest heap node: floor(log2(2n-1))
    void *heap = nullptr; // device, 2n elements
    bool owns_heap = false;
    unsigned int *d_status = nullptr;  // device status word (+ spare counters), 64 words
    unsigned int *h_status = nullptr;  // pinned host mirror, 64 words
    st_workspace ws;
    st_workspace scratch;
    // small host calls: inputs/outputs go through mapped pinned memory (no cudaMemcpy round trips)
    void *h_pin = nullptr;   // host address
    void *d_pin = nullptr;   // device alias of the same memory
    size_t pin_bytes = 0;   // grow-only staging area of the *_host entry points (no malloc/free per call)
    // sharded plans: 4 result slots (per-shard counts + flags, 1 KiB each) behind the staging area,
    // and the event that marks the end of the latest plan's kernels
    cudaEvent_t plan_ev = nullptr;
    int plan_pending = 0;
    int64_t reserved_batch = 0;
    int64_t cub_cap = -1;      // batch capacity for which cub_bytes was queried
    size_t cub_bytes = 0;
    int sm_count = 148;
    int coop_ok = 0;
    // optional per-phase timing of the update pipeline (st_set_profiling)
    int profile = 0;
    cudaEvent_t prof_ev[8] = {};
    double prof_ms[8] = {};   // sort, leaf pass, partition, retire, huge maps, stitch + folds, second pass, calls
};

// launchers implemented in the .cu files; every one only enqueues on `s`
namespace st {

// process-wide count of kernels of THIS library that were launched (bench.py's gpu_launches)
extern unsigned long long g_kernel_launches;

cudaError_t launch_check_leaves(st_tree *t, cudaStream_t s);  // sets STF_NEG_ARRAY
cudaError_t launch_build(st_tree *t, cudaStream_t s);         // skipped on device if status != 0
// m_dev (nullable): device-side query count, the launch covers min(m, *m_dev) queries
cudaError_t launch_search(const st_tree *t, const void *vals, int64_t m, int64_t *out, cudaStream_t s,
                          const unsigned int *m_dev = nullptr);
cudaError_t launch_sample(const st_tree *t, const double *u, int64_t m, int64_t *out, void *prio,
                          void *resid, cudaStream_t s);
cudaError_t launch_publish(const st_tree *t, void *dst_status, void *dst_root, cudaStream_t s);  // tiny: status word + root
cudaError_t launch_gather(const st_tree *t, const int64_t *idx, int64_t m, void *out, cudaStream_t s);
cudaError_t launch_sum_over(const st_tree *t, int64_t a, int64_t b, void *out_dev, cudaStream_t s);
cudaError_t launch_strided_sum(const st_tree *t, int64_t stride, void *out, int64_t nout, cudaStream_t s);
cudaError_t launch_axis_fold(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out, cudaStream_t s);

size_t update_workspace_bytes(const st_tree *t, int64_t batch);
// delta_out (nullable): receives d_j in batch order.  deltas_in (nullable): apply the given
// per-update differences to the proper ancestors only (vals ignored, leaves untouched).
cudaError_t launch_update(st_tree *t, const int64_t *idx, int64_t ni, const void *vals, int64_t nv,
                          void *delta_out, const void *deltas_in, cudaStream_t s);
int64_t update_max_chunk();
// tiny power-of-two float tree: ordered fold of all m differences into the ancestors of leaf[j]
cudaError_t launch_fold_by_leaf(st_tree *t, const int64_t *leaf, const void *deltas, int64_t m, cudaStream_t s);
// the same in two phases (chunk summaries for a chunk range, then the stitch); records [chunk][node-1]
size_t fold_rec_bytes(const st_tree *t);
// leaf ids are int64 (leaf_bytes 8) or uint8 (leaf_bytes 1)
cudaError_t launch_fold_by_leaf_maps(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                     int chunk_lo, int chunk_hi, void *recs, cudaStream_t s);
cudaError_t launch_fold_by_leaf_apply(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                      const void *recs, cudaStream_t s);

// subtree-sharded layout helpers (st_shard.cu)
size_t shard_plan_scratch_bytes(int64_t m, int shards);
// counts_out: [shards] element counts, then one word of STF_* validation flags
cudaError_t launch_shard_plan(st_tree *t, const int64_t *ids, int64_t m, int shift, int shards, int rank,
                              const void *payload, int64_t npayload, int check_negative, uint8_t *shard_out,
                              int32_t *slot_out, int64_t *own_ids_out, void *own_payload_out, void *scratch,
                              unsigned int *counts_out, cudaStream_t s);
cudaError_t launch_shard_expand_values(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                       const void *slabs, int64_t slab_stride, void *out, cudaStream_t s);
cudaError_t launch_shard_expand_indices(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                        const int32_t *slabs, int64_t slab_stride, int64_t leaves_per_shard,
                                        int64_t *out, cudaStream_t s);

}  // namespace st
