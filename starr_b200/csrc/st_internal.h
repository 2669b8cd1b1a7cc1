// st_internal.h -- shared between the translation units of libstarr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

#include "../../include/starr_b200.h"

// Device-side status word bits (written by validation kernels, read by every later kernel
// of the same batch: a failed batch is skipped as a whole, see DESIGN.md "error atomicity").
#define STF_NEG_VALS 1u
#define STF_NEG_ARRAY 2u
#define STF_INDEX 4u
// status word 0 holds the flags of the batch IN FLIGHT (every later kernel of that batch returns at once
// when it is set); the first kernel of the next batch moves them to the sticky word, which st_poll_status
// reports, so one failed batch does not silently drop the batches behind it.
#define ST_SW_STICKY 21
#define ST_SW_ZERO_DELTA 20

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
    int kind = 0;         // 0 = sum tree, 1 = min tree (st_min_create)
    void *aux = nullptr;  // min tree: int tag[n] (last-writer claims)
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
    int part_occ = 0;          // resident CTAs per SM of the cooperative partition kernel (0 = not queried yet)
    // The reference's functions hold the GIL (SURVEY 8b "Threading"); here every entry point that touches the
    // handle's host-side state or staging areas takes this lock, so concurrent callers on ONE tree serialise.
    std::recursive_mutex mu;
    int64_t pending_ni = 0;    // split update (st_update_begin_device): batch size waiting for st_update_finish_device
    int pending_split = 0;
    int pending_gb = 0;        // leaf-group bits of the latest float batch (deep_red_kernel / retire_kernel)
    int64_t stat_updates = 0;  // updates submitted since the last st_update_counters(reset)
    // pipelined host step (st_per_step_host_submit / _wait): ST_HOST_SLOTS slots of device staging + streams + events
    struct host_slot {
        void *dev = nullptr;       // idx | vals | u | out_idx | out_prio | out_prob
        size_t bytes = 0;
        cudaEvent_t ev_in = nullptr, ev_free = nullptr, ev_done = nullptr, ev_read = nullptr;
        unsigned int *h_flag = nullptr;   // pinned: validation flags of the slot's batch
        int busy = 0;
    } hs[4];   // ST_HOST_SLOTS
    cudaStream_t hs_in = nullptr, hs_main = nullptr, hs_out = nullptr;
    // sector-packed search index (st_sidx.cuh): copies of the left children below the staged levels, three levels
    // per 32-byte record; nullptr when the tree is too small / not a 4-byte sum tree / ST_SEARCH_INDEX=0
    void *sidx = nullptr;
    size_t sidx_bytes = 0;
    int sidx_nb = 0;
    // update pipeline: retirement (ordered atomics on the deep nodes) runs on this stream beside the folds of
    // the top nodes (disjoint node sets); forked / joined with the two events around every batch
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_hy = nullptr;   // histogram + scan of the hybrid partition, run ahead on the side stream beside sort + leaf pass
    int hy_ahead = 0;              // ... for the batch whose first half is enqueued
    // optional per-phase timing of the update pipeline (st_set_profiling)
    int profile = 0;
    cudaEvent_t prof_ev[8] = {};
    double prof_ms[8] = {};   // sort, leaf pass, partition, retire, huge maps, stitch + folds, second pass, calls
};

// launchers implemented in the .cu files; every one only enqueues on `s`
namespace st {

// epilogue of the sharded sample (st_peer.cu): instead of out[i], the GLOBAL leaf index of query i is
// stored at position pos[i] of an int64 array in the exchange buffer of EVERY rank (peer stores)
struct search_push {
    char *base[64];
    int n = 0;                     // 0: no push (plain search)
    const int32_t *pos = nullptr;
    int64_t index_offset = 0;
    int32_t *pairs = nullptr;      // != nullptr: result i goes to pairs[2i] = pos[i], pairs[2i+1] = local leaf (no push)
    int32_t *leaf32 = nullptr;     // != nullptr: result i goes to leaf32[i] as a 32-bit local leaf (slices layout)
    // segmented input (slices layout, sample): the queries are n segments of seg_cnt[r] elements, segment r at
    // in[r * seg_stride ..] / pos[r * seg_stride ..]; result k of segment r is stored as the pair (pos, local leaf) at
    // element k of base[r] (peer memory: consecutive queries of a segment are consecutive 8-byte stores)
    const unsigned int *seg_cnt = nullptr;
    int64_t seg_stride = 0;
};

// process-wide count of kernels of THIS library that were launched (bench.py's gpu_launches)
extern std::atomic<unsigned long long> g_kernel_launches;

// run `f` once per CUDA device (function attributes are per device); thread-safe
struct per_device_once {
    std::atomic<unsigned long long> mask{0};
    std::mutex mu;
    template <typename F> cudaError_t run(int dev, F f)
    {
        const unsigned long long bit = 1ull << (dev & 63);
        if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
        std::lock_guard<std::mutex> g(mu);
        if (mask.load(std::memory_order_relaxed) & bit) return cudaSuccess;
        cudaError_t e = f();
        if (e != cudaSuccess) return e;
        mask.fetch_or(bit, std::memory_order_release);
        return cudaSuccess;
    }
};

cudaError_t launch_batch_begin(st_tree *t, cudaStream_t s);   // flags of the previous batch -> sticky word
cudaError_t launch_check_leaves(st_tree *t, cudaStream_t s);  // sets STF_NEG_ARRAY
cudaError_t launch_build(st_tree *t, cudaStream_t s);         // skipped on device if status != 0; refreshes the search index
cudaError_t launch_sidx_build(st_tree *t, cudaStream_t s);    // search index from the current heap (no-op without index)
size_t sidx_bytes_for(int64_t n, int depth, int esize, int kind, int *nb_out);
// m_dev (nullable): device-side query count, the launch covers min(m, *m_dev) queries
cudaError_t launch_search(const st_tree *t, const void *vals, int64_t m, int64_t *out, cudaStream_t s,
                          const unsigned int *m_dev = nullptr, const search_push *push = nullptr);
cudaError_t launch_route_compact(const st_tree *top, const double *u, int64_t m, int rank, void *own_resid,
                                 int32_t *own_pos, unsigned int *own_count, cudaStream_t s);
// prob (nullable): prio / root in the tree dtype for f32 / f64, in float64 for integer trees
cudaError_t launch_sample(const st_tree *t, const double *u, int64_t m, int64_t *out, void *prio,
                          void *resid, cudaStream_t s, void *prob = nullptr);
cudaError_t launch_publish(const st_tree *t, void *dst_status, void *dst_root, cudaStream_t s);  // tiny: status word + root
cudaError_t launch_gather(const st_tree *t, const int64_t *idx, int64_t m, void *out, cudaStream_t s);
// leaves = leaves (op) scalar | operand[n] (st_inplace_op); the tree is NOT rebuilt
cudaError_t launch_inplace(st_tree *t, int op, const void *scalar_host, const void *operand, cudaStream_t s);
cudaError_t launch_sum_over(const st_tree *t, int64_t a, int64_t b, void *out_dev, cudaStream_t s);
cudaError_t launch_strided_sum(const st_tree *t, int64_t stride, void *out, int64_t nout, cudaStream_t s);
cudaError_t launch_axis_fold(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out, cudaStream_t s);

// min tree (st_mintree.cu)
cudaError_t launch_min_init(st_tree *t, cudaStream_t s);
cudaError_t launch_min_build(st_tree *t, cudaStream_t s);
cudaError_t launch_min_update(st_tree *t, const int64_t *idx, int64_t ni, const void *vals, int64_t nv, cudaStream_t s);

size_t update_workspace_bytes(const st_tree *t, int64_t batch);
// delta_out (nullable): receives d_j in batch order.  deltas_in (nullable): apply the given
// per-update differences to the proper ancestors only (vals ignored, leaves untouched).
// assign: the leaf becomes vals[j] itself instead of fl(leaf + fl(vals[j] - leaf)) (starr.experimental twin)
cudaError_t launch_update(st_tree *t, const int64_t *idx, int64_t ni, const void *vals, int64_t nv,
                          void *delta_out, const void *deltas_in, cudaStream_t s, bool assign = false, int phase = 0);
int64_t update_max_chunk();
// tiny power-of-two float tree: ordered fold of all m differences into the ancestors of leaf[j]
cudaError_t launch_fold_by_leaf(st_tree *t, const int64_t *leaf, const void *deltas, int64_t m, cudaStream_t s);
// the same in two phases (chunk summaries for a chunk range, then the stitch); records [chunk][node-1]
size_t fold_rec_bytes(const st_tree *t);
// leaf ids are int64 (leaf_bytes 8) or uint8 (leaf_bytes 1)
cudaError_t launch_fold_by_leaf_maps(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                     int chunk_lo, int chunk_hi, void *recs, cudaStream_t s);
// slice_d / slice_leaf (nullable, device arrays of per-rank pointers): the batch is distributed over the ranks, element i
// lives on rank i / slice_span at position i % slice_span (positions >= slice_m are padding); leaf / deltas unused then
cudaError_t launch_fold_by_leaf_apply(st_tree *t, const void *leaf, int leaf_bytes, const void *deltas, int64_t m,
                                      const void *recs, cudaStream_t s, const void *const *slice_d = nullptr,
                                      const void *const *slice_leaf = nullptr, int64_t slice_span = 0, int64_t slice_m = 0);

// subtree-sharded layout helpers (st_shard.cu)
size_t shard_plan_scratch_bytes(int64_t m, int shards);
// counts_out: [shards] element counts, then one word of STF_* validation flags
cudaError_t launch_shard_plan(st_tree *t, const int64_t *ids, int64_t m, int shift, int shards, int rank,
                              const void *payload, int64_t npayload, int check_negative, uint8_t *shard_out,
                              int32_t *slot_out, int64_t *own_ids_out, void *own_payload_out, void *scratch,
                              unsigned int *counts_out, cudaStream_t s, int32_t *own_pos_out = nullptr);
// slices layout: stable G-way partition of a slice by owner (shard-local ids + payload in destination order);
// row_out: [shards] counts, then one word of STF_* flags (device; nothing is read back)
cudaError_t launch_shard_partition(st_tree *t, const int64_t *ids, int64_t m, int shift, int shards, const void *payload,
                                   int check_negative, uint8_t *shard_out, int32_t *slot_out, int64_t *part_ids_out,
                                   void *part_payload_out, void *scratch, unsigned int *row_out, cudaStream_t s);
cudaError_t launch_shard_expand_values(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                       const void *slabs, int64_t slab_stride, void *out, cudaStream_t s);
cudaError_t launch_shard_expand_cat(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m, const void *cat,
                                    const int64_t *starts_host, int shards, void *out, cudaStream_t s);
cudaError_t launch_shard_expand_indices(const st_tree *t, const uint8_t *shard, const int32_t *slot, int64_t m,
                                        const int32_t *slabs, int64_t slab_stride, int64_t leaves_per_shard,
                                        int64_t *out, cudaStream_t s);

}  // namespace st
