/*
 * starr_b200.h -- C ABI of libstarr_b200.so: a device-resident sum tree for NVIDIA B200
 * (sm_100a) that replaces the reference kernel module `starr._cython`
 * (reference: starr/src/_sumtree_array.pyx, built by setup.py:50-55 and re-exported at
 * starr/__init__.py:1-5).
 *
 * The reference module is stateless: every call receives the caller's NumPy buffers
 * (array[n], sumtree[n]) and works on them in place.  The device replacement is stateful:
 * one opaque handle owns (or borrows) ONE device buffer `heap[2n]` of the element type,
 *
 *      heap[0]        = 0                       (sumtree[0], never read)
 *      heap[h], 1<=h<n = internal node h        (sumtree[h]);  children 2h, 2h+1
 *      heap[n + i]    = leaf i                  (array[i])
 *
 * so the reference's two-array child lookup (_sumtree_array.pyx:15-20) is one load.
 *
 * Conventions: every function returns an st_status (0 = ok).  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  *_device entry points
 * take device pointers and only enqueue work; *_host entry points take host pointers,
 * copy through the same stream and return after the stream has drained.  No torch
 * types, no C++ types, plain pointers and sizes only.  Nothing here falls back to the
 * CPU: if there is no usable CUDA device every call returns ST_ERR_CUDA.
 */
#ifndef STARR_B200_H
#define STARR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ST_API __attribute__((visibility("default")))
#else
#define ST_API
#endif

typedef struct st_tree st_tree; /* opaque */

/* element types = the members of the reference's fused ARRAY_TYPE
 * (_sumtree_array.pyx:4-11) that exist on the GPU (no 80-bit long double). */
typedef enum st_dtype {
    ST_I16 = 0,
    ST_I32 = 1,
    ST_I64 = 2,
    ST_U8 = 3,
    ST_U16 = 4,
    ST_U32 = 5,
    ST_U64 = 6,
    ST_F32 = 7,
    ST_F64 = 8
} st_dtype;

/* status codes; the Python layer maps them 1:1 onto the reference's exceptions */
typedef enum st_status {
    ST_OK = 0,
    ST_ERR_NEGATIVE_VALS = 1,  /* ValueError "vals must be non-negative"            pyx:37-39 */
    ST_ERR_NEGATIVE_ARRAY = 2, /* ValueError "array elements must be non-negative"  pyx:58-59 */
    ST_ERR_INDEX = 3,          /* IndexError (Cython boundscheck)                   pyx:41-49 */
    ST_ERR_EMPTY_VALS = 4,     /* ZeroDivisionError (i % nv with nv == 0)           pyx:44    */
    ST_ERR_SIZE = 5,           /* TypeError "... must have the same size" / "invalid output size" */
    ST_ERR_DTYPE = 6,          /* TypeError "No matching signature found"           pyx:4-11  */
    ST_ERR_ARG = 7,            /* ValueError: bad argument (n < 2, null pointer, ...)         */
    ST_ERR_CUDA = 8,           /* RuntimeError: CUDA runtime failure (see st_last_error)      */
    ST_ERR_TOO_LARGE = 9       /* n beyond the 32-bit heap-index limit (n <= 2^31 - 1)        */
} st_status;

/* flags for st_update_device */
#define ST_UPDATE_DEFAULT 0
#define ST_UPDATE_SYNC_CHECK 1 /* wait for the batch and return its validation status */
/* the leaf becomes vals[j] itself instead of fl(leaf + fl(vals[j] - leaf)): the semantics of the
 * reference's experimental C++ twin (starr/experimental/src/sumtree_array.cpp:19-23); ancestors still
 * receive the ordered differences */
#define ST_UPDATE_ASSIGN_LEAF 2

/* in-place leaf arithmetic, st_inplace_device */
typedef enum st_inplace_op {
    ST_OP_ADD = 0,  /* x += c   */
    ST_OP_SUB = 1,  /* x -= c   */
    ST_OP_MUL = 2,  /* x *= c   */
    ST_OP_DIV = 3,  /* x /= c   (float trees only, as in NumPy) */
    ST_OP_FILL = 4  /* x.fill(c) */
} st_inplace_op;

ST_API int st_abi_version(void); /* 4 */
ST_API const char *st_status_string(int status);
ST_API const char *st_last_error(void); /* detail of the last failure on this thread */
ST_API int st_dtype_size(int dtype);     /* bytes, or -1 */
/* number of kernels of this library launched by this process so far (profiling aid) */
ST_API unsigned long long st_kernel_launches(void);

/* ---- lifetime -------------------------------------------------------------------- */
/* Create a tree of n >= 2 zero leaves on CUDA device `device`.  If external_heap is not
 * NULL it must be a device buffer of 2n elements that outlives the handle (e.g. a torch
 * tensor); otherwise the library allocates it.  The tree starts all-zero (consistent). */
ST_API int st_create(int64_t n, int dtype, int device, void *external_heap, st_tree **out);
ST_API int st_destroy(st_tree *t);
ST_API int64_t st_size(const st_tree *t);
ST_API int st_dtype_of(const st_tree *t);
ST_API int st_device_of(const st_tree *t);
ST_API void *st_heap_ptr(const st_tree *t); /* device pointer to heap[2n] */
/* Pre-size the update workspace for batches of up to max_batch updates (so that no
 * allocation happens later, e.g. inside CUDA-graph capture). */
ST_API int st_reserve(st_tree *t, int64_t max_batch);

/* ---- raw state transfer (what the stateless reference functions need) ------------ */
/* leaves -> heap[n..2n), tree -> heap[0..n); either may be NULL.  No validation, no build:
 * replaces handing (array, sumtree) to the module, _sumtree_array.pyx:22-26. */
ST_API int st_load_host(st_tree *t, const void *leaves, const void *tree);
ST_API int st_store_host(const st_tree *t, void *leaves, void *tree);
ST_API int st_gather_leaves_host(const st_tree *t, const int64_t *idx_host, int64_t m, void *out_host);
ST_API int st_gather_leaves_device(const st_tree *t, const int64_t *idx_dev, int64_t m, void *out_dev,
                            void *stream);

/* ---- build: replaces build_sumtree_from_array, _sumtree_array.pyx:51-64 ---------- */
/* Reject any negative leaf BEFORE touching the tree (ST_ERR_NEGATIVE_ARRAY), then
 * node = left + right for h = n-1 .. 1 (pairwise, level by level). */
ST_API int st_build(st_tree *t, void *stream);                 /* from the resident leaves; waits */
ST_API int st_build_from_host(st_tree *t, const void *leaves); /* H2D leaves, then st_build */
ST_API int st_build_from_device(st_tree *t, const void *leaves_dev, void *stream);

/* ---- update: replaces update_sumtree, _sumtree_array.pyx:22-49 ------------------- */
/* For j = 0..ni-1 in batch order: d = vals[j % nv] - leaf; leaf += d; every ancestor += d.
 * Bit-exact with the sequential reference for every dtype (floats: ordered per-node folds).
 * Validation (negative vals over ALL nv entries, index range) happens before any write. */
ST_API int st_update_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev,
                     int64_t nv, void *stream, int flags);
ST_API int st_update_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host,
                   int64_t nv);
/* st_update_host with the ST_UPDATE_* flags of st_update_device (ST_UPDATE_ASSIGN_LEAF: the stateless
 * experimental twin, starr/experimental/src/_sumtree_array.pyx:22-32) */
ST_API int st_update_host_ex(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host,
                      int64_t nv, int flags);
/* st_update_host that also returns the resulting leaf values of idx (what SumTreeArray.put needs
 * for its host mirror: floats store fl(old + fl(val - old)), not val), in one staging round. */
ST_API int st_update_gather_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host,
                          int64_t nv, void *leaves_out_host);
/* the same, also returning the root after the update (the host layer's `sum()` then needs no call of its own) */
ST_API int st_update_gather_root_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv,
                               void *leaves_out_host, void *root_out_host);

/* ---- building blocks of the subtree-sharded tree (no reference counterpart: the reference is
 * single-process; these keep the sharded result bit-identical to ONE reference tree) -------- */
/* st_update_device that also returns every update's applied difference d_j (batch order,
 * tree dtype) -- the quantity the reference adds to each ancestor (pyx:43-48). */
ST_API int st_update_deltas_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev,
                            int64_t nv, void *delta_out_dev, void *stream);
/* st_update_deltas_device in two halves, so that the differences can travel to the other ranks while
 * the ancestors are still being updated: begin = validation, leaf order, leaf pass (delta_out is complete
 * in stream order behind it); finish = everything above the leaves, for the batch begin left pending.
 * Batches the library applies in one piece anyway (<= 4096 updates, or more than one 2^21 chunk) are
 * applied completely by begin; finish is then a no-op. */
ST_API int st_update_begin_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                           void *delta_out_dev, void *stream);
ST_API int st_update_finish_device(st_tree *t, void *stream);
/* For j in batch order: every PROPER ancestor of leaf idx[j] += deltas[j] (ordered fold for
 * floats); leaves are not touched.  Used on the replicated top tree whose leaves are the
 * shard roots. */
ST_API int st_apply_deltas_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *deltas_dev,
                           void *stream);
/* st_apply_deltas_device for a tiny power-of-two float tree (n <= 64) in two phases, so that the
 * ranks of a sharded tree can split phase A: chunk c covers batch positions [1024c, 1024c+1024);
 * phase A writes one record per (chunk, internal node) at recs[(c*(n-1) + node-1) * record_bytes]
 * for chunk_lo <= c < chunk_hi; phase B needs the records of ALL ceil(m/1024) chunks. */
ST_API int st_top_fold_record_bytes(const st_tree *t);
/* leaf ids: int64 (leaf_itemsize 8) or uint8 (leaf_itemsize 1, as written by st_shard_plan_device) */
ST_API int st_top_fold_maps_device(st_tree *t, const void *leaf_dev, int leaf_itemsize, const void *deltas_dev,
                            int64_t m, int64_t chunk_lo, int64_t chunk_hi, void *recs_dev, void *stream);
ST_API int st_top_fold_apply_device(st_tree *t, const void *leaf_dev, int leaf_itemsize, const void *deltas_dev,
                             int64_t m, const void *recs_dev, void *stream);
/* Plan of one replicated batch of the sharded tree (m > 0).  Element j belongs to shard
 * ids[j] >> shift.  Reads nothing of the tree, so it may run on a side stream ahead of the batch it
 * plans.  Writes
 *   shard_out[j]   the owning shard (255 for an id outside [0, shards << shift)),
 *   slot_out[j]    the stable position of j among the elements of its shard (-1 if invalid),
 *   own_ids_out[k], own_payload_out[k]   (both nullable) for the k-th element of shard `rank`:
 *                  ids[j] - (rank << shift) and payload[j % npayload] (tree dtype),
 * and, into result slot `result_slot` (0..3; a slot is overwritten by the next plan that names it),
 * the number of elements of every shard (shards <= 64).
 * counts_out_host != NULL: synchronous -- waits for the plan, fills counts_out_host[shards] and
 * returns ST_ERR_INDEX for an id outside every shard or, with check_negative, ST_ERR_NEGATIVE_VALS for
 * a negative entry anywhere in payload (pyx:37-39): the same verdict on every rank, before anything
 * was modified.  counts_out_host == NULL: only enqueues; st_shard_plan_wait() collects the same result
 * later and waits for the plan's kernels only, so work that needs nothing but the DEVICE-side counts
 * (st_shard_plan_counts_device, st_search_counted_device) can be enqueued in between.
 * shard_out must be 8-byte and slot_out 16-byte aligned (vector stores). */
ST_API int st_shard_plan_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards, int rank,
                         const void *payload_dev, int64_t npayload, int check_negative, uint8_t *shard_out_dev,
                         int32_t *slot_out_dev, int64_t *own_ids_out_dev, void *own_payload_out_dev, int result_slot,
                         int64_t *counts_out_host, void *stream);
/* the same, additionally writing own_pos_out[k] = batch position j of the k-th element of shard `rank`
 * (nullable; m < 2^31): where that element's results are stored on every rank by the peer exchange */
ST_API int st_shard_plan_pos_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards, int rank,
                             const void *payload_dev, int64_t npayload, int check_negative, uint8_t *shard_out_dev,
                             int32_t *slot_out_dev, int64_t *own_ids_out_dev, void *own_payload_out_dev,
                             int32_t *own_pos_out_dev, int result_slot, int64_t *counts_out_host, void *stream);
ST_API int st_shard_plan_wait(st_tree *t, int result_slot, int shards, int64_t *counts_out_host);
/* device address of result slot `result_slot`: uint32 counts[shards] (valid in stream order behind the plan) */
ST_API const uint32_t *st_shard_plan_counts_device(const st_tree *t, int result_slot);
/* st_search_device over min(m_max, *m_dev) queries: the count is read on the device at launch time */
ST_API int st_search_counted_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                             int64_t *out_dev, void *stream);
/* Per-shard results come back as slabs (shard s at slabs[s * slab_stride ...], compact, in the
 * shard's batch order); these put them back into batch order:
 *   values:  out[j] = slabs[shard[j] * slab_stride + slot[j]]               (tree dtype)
 *   indices: out[j] = shard[j] * leaves_per_shard + slabs[...]              (int32 slabs -> int64) */
ST_API int st_shard_expand_values_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                                  const void *slabs_dev, int64_t slab_stride, void *out_dev, void *stream);
/* out[j] = cat[starts[shard[j]] + slot[j]]: the shards' slabs stored back to back (peer exchange); starts_host: `shards`
 * element offsets, read during the call */
ST_API int st_shard_expand_cat_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                               const void *cat_dev, const int64_t *starts_host, int shards, void *out_dev, void *stream);
ST_API int st_shard_expand_indices_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev,
                                   int64_t m, const int32_t *slabs_dev, int64_t slab_stride,
                                   int64_t leaves_per_shard, int64_t *out_dev, void *stream);
/* sample() descent that also returns the prefix left over at the leaf that was reached:
 * on the top tree this is (owning shard, query to continue with inside that shard). */
ST_API int st_route_uniform_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *leaf_out_dev,
                            void *residual_out_dev, void *stream);

/* ---- peer-memory exchange of the sharded tree (one node, NVLink / NVSwitch peer access) ------------
 * Every rank owns one device buffer [flag page | payload]; its CUDA IPC handle (st_peer_handle_bytes()
 * bytes) is exchanged once by the host layer, after which kernels store results straight into the
 * buffers of all ranks.  All calls are stream-ordered and never synchronise with the host. */
typedef struct st_peer st_peer; /* opaque */
ST_API const char *st_peer_last_error(void);
ST_API int st_peer_handle_bytes(void);
ST_API int st_peer_create(int device, int rank, int world, int64_t payload_bytes, st_peer **out,
                   unsigned char *handle_out);
/* handles: world x st_peer_handle_bytes() bytes, rank-major (this rank's own entry is ignored) */
ST_API int st_peer_connect(st_peer *p, const unsigned char *handles);
/* a peer that lives in THIS process (ranks as threads; CUDA IPC cannot open a handle in the process that
 * exported it): base = st_peer_base() of that rank's area, on a device this one has peer access to.  Call
 * before st_peer_connect, which skips adopted ranks. */
ST_API int st_peer_adopt(st_peer *p, int q, void *base);
ST_API void *st_peer_base(const st_peer *p); /* start of the local buffer (flag page) */
/* closes the mappings of the other ranks' buffers (every rank calls this, then a host barrier, then destroy:
 * an exported buffer must not be freed while another process still has it open) */
ST_API int st_peer_disconnect(st_peer *p);
ST_API int st_peer_destroy(st_peer *p);
ST_API void *st_peer_payload(const st_peer *p); /* local device pointer of the payload */
ST_API int64_t st_peer_payload_bytes(const st_peer *p);
/* payload_q[offset .. offset+bytes) = this rank's payload[offset ..) for every other rank q (multiples of 4) */
ST_API int st_peer_broadcast(st_peer *p, int64_t offset, int64_t bytes, void *stream);
/* payload_q[offset + pos[k]*esize] = src[k] on EVERY rank q (this one included), k < min(count, *count_dev);
 * pos values must lie in [0, span_elems); count_dev may be NULL */
ST_API int st_peer_scatter(st_peer *p, const void *src_dev, const int32_t *pos_dev, int64_t count,
                    const uint32_t *count_dev, int esize, int64_t offset, int64_t span_elems, void *stream);
/* payload_q[dst_offset + k*esize] = src[k], k < min(count, *count_dev), on EVERY rank q: consecutive elements, so every
 * warp store is one full-width NVLink write (the scattered stores of st_peer_scatter cost a packet per element).
 * esize 4 or 8; count_dev nullable; count_offset >= 0: the number of elements also goes to payload_q[count_offset] (uint32) */
ST_API int st_peer_put(st_peer *p, const void *src_dev, int64_t count, const uint32_t *count_dev, int esize,
                int64_t dst_offset, int64_t count_offset, void *stream);
/* out[pos] = q * leaves_per_shard + idx for the int32 (pos, idx) pairs every rank q has put into THIS rank's buffer: rank q's
 * pairs start at payload[pairs_offset + q*slab_bytes], their number is the uint32 at payload[counts_offset + 4q] */
ST_API int st_peer_scatter_pairs(st_peer *p, int64_t pairs_offset, int64_t slab_bytes, int64_t counts_offset,
                          int64_t leaves_per_shard, int64_t *out_dev, int64_t m, void *stream);
/* st_search_counted_device whose result k is stored as the int32 pair (pos[k], local leaf) in pairs_out[2k], [2k+1] */
ST_API int st_search_pairs_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                           const int32_t *pos_dev, int32_t *pairs_out_dev, void *stream);
/* release-store `value` into flags[channel][this rank] of every rank / spin until flags[channel][q] has
 * reached `value` for every q (wrap-safe compare; gives up after 20 s and records it for st_peer_status) */
ST_API int st_peer_signal(st_peer *p, int channel, unsigned int value, void *stream);
ST_API int st_peer_wait(st_peer *p, int channel, unsigned int value, void *stream);
ST_API int st_peer_status(st_peer *p, void *stream); /* waits for stream; ST_ERR_CUDA after a timed-out wait */
/* sample() routing on the replicated top tree fused with the selection of this rank's queries: for the
 * queries whose top-tree descent ends in leaf `rank`, appends (left-over prefix, batch position) to
 * own_resid_out / own_pos_out (unordered) and counts them in *own_count_out_dev. */
ST_API int st_route_compact_device(const st_tree *top, const double *u_dev, int64_t m, int rank, void *own_resid_out_dev,
                            int32_t *own_pos_out_dev, uint32_t *own_count_out_dev, void *stream);
/* st_search_counted_device whose result for query k -- index_offset + local leaf -- is stored as int64 at
 * payload_q[payload_offset + 8*pos[k]] on every rank q */
ST_API int st_search_push_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                          const int32_t *pos_dev, int64_t index_offset, st_peer *p, int64_t payload_offset,
                          int64_t span_elems, void *stream);

/* ---- slices layout: the batch is DISTRIBUTED over the ranks (no reference counterpart) ----------
 * Rank r holds elements [r m, (r+1) m) of the global batch (data-parallel learners: every rank updates the
 * priorities of, and samples, its own 1/G of the batch), so nobody touches O(M_global) data.  One exchange:
 *   st_shard_partition_device   owner (uint8) + stable slot of every slice element, the slice partitioned by owner
 *                               (shard-local ids / payload, one contiguous run per destination) and a row of
 *                               world counts + 1 word of validation flags -- all on the device, no host round trip
 *   st_slices_rows              the row goes into the count matrix M[world][world + 1] of every rank, then the
 *                               release flag (channel, value); st_peer_wait completes the matrix
 *   st_slices_matrix_fetch/wait asynchronous host copy of the matrix (the local update needs its element count)
 *   st_slices_send              up to two arrays of the partitioned slice go to their owners' receive areas, each
 *                               owner receiving its elements in GLOBAL batch order (rank-major), contiguous;
 *                               *total_out_dev = elements this rank receives
 *   st_slices_return            the owner's results, in receive order, back to the source's return area at the
 *                               position the element had in the source's partitioned slice
 *   st_slices_unpermute         out[j] = ret[start[shard[j]] + slot[j]]  (leaves_per_shard > 0: ret holds int32 local
 *                               leaves and out[j] = shard * leaves_per_shard + leaf as int64)
 * Offsets are byte offsets into the payload; matrix_offset must be 4-byte, the others 8-byte aligned. */
ST_API int st_shard_partition_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards,
                              const void *payload_dev, int check_negative, uint8_t *shard_out_dev,
                              int32_t *slot_out_dev, int64_t *part_ids_out_dev, void *part_payload_out_dev,
                              uint32_t *row_out_dev, void *stream);
ST_API int st_slices_rows(st_peer *p, const uint32_t *row_dev, int64_t matrix_offset, int channel, unsigned int value,
                   void *stream);
ST_API int st_slices_matrix_fetch(st_peer *p, int64_t matrix_offset, void *stream);
ST_API int st_slices_matrix_wait(st_peer *p, uint32_t *matrix_out_host); /* world * (world + 1) words */
ST_API int st_slices_send(st_peer *p, int64_t matrix_offset, const void *src0_dev, int esize0, int64_t dst_offset0,
                   const void *src1_dev, int esize1, int64_t dst_offset1, int64_t m_max, uint32_t *total_out_dev,
                   void *stream);
ST_API int st_slices_return(st_peer *p, int64_t matrix_offset, const void *src_dev, int esize, int64_t dst_offset,
                     int64_t total_max, void *stream);
ST_API int st_slices_unpermute(st_peer *p, int64_t matrix_offset, const uint8_t *shard_dev, const int32_t *slot_dev,
                        int64_t m, int64_t ret_offset, int esize, int64_t leaves_per_shard, void *out_dev,
                        int64_t shard_copy_offset /* >= 0: shard ids are copied there too */, void *stream);
/* Top fold over a distributed batch.  maps_slice: chunk summaries of THIS rank's slice (m_slice elements, nchunks =
 * span / 1024 chunks, padding chunks are identities) into records [record_first, record_first + nchunks).
 * st_slices_table(parity, d_offset, leaf_offset): where every rank keeps its slice's differences and shard ids (uint8)
 * inside its exchange buffer; apply_slices stitches the records of all ranks and reads those arrays through peer
 * memory only for chunks that must be folded one add at a time. */
ST_API int st_top_fold_maps_slice_device(st_tree *t, const uint8_t *leaf_dev, const void *deltas_dev, int64_t m_slice,
                                  int64_t record_first, int64_t nchunks, void *recs_dev, void *stream);
ST_API int st_slices_table(st_peer *p, int parity, int64_t d_offset, int64_t leaf_offset);
ST_API int st_top_fold_apply_slices_device(st_tree *t, st_peer *p, int parity, int64_t span, int64_t m_slice,
                                    const void *recs_dev, void *stream);
/* sample() of the slices layout without a count matrix (one flag round trip each way):
 *   st_route_push_device            routes this rank's own uniforms on the top tree and stores every left-over prefix,
 *                                   with its slice position, straight into the owner's receive region of this source
 *                                   (element [rank * cap + k], k from a local counter per owner; cnt_dev: world counters)
 *   st_slices_counts                the counters go to the owners (cnt_offset + 4 * rank) and into this rank's own buffer
 *                                   (sent_offset + 4 * owner), then the release flag (channel, value)
 *   st_search_segments_push_device  the local descent over the world segments that arrived; result k of source r's
 *                                   segment is stored as the pair (slice position, local leaf) at element
 *                                   [rank * cap + k] of r's back region -- consecutive queries, consecutive stores
 *   st_peer_scatter_pairs           (above) turns the pairs that came back into global indices in slice order */
ST_API int st_route_push_device(const st_tree *top, const double *u_dev, int64_t m, st_peer *p, int64_t resid_offset,
                         int64_t pos_offset, int64_t cap, uint32_t *cnt_dev, void *stream);
ST_API int st_slices_counts(st_peer *p, const uint32_t *cnt_dev, int64_t cnt_offset, int64_t sent_offset, int channel,
                     unsigned int value, void *stream);
ST_API int st_search_segments_push_device(const st_tree *t, st_peer *p, int64_t resid_offset, int64_t pos_offset,
                                   int64_t cnt_offset, int64_t cap, int64_t back_offset, void *stream);
/* st_search_counted_device with 32-bit local leaves as the result (what travels back to the querying rank) */
ST_API int st_search_leaf32_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                            int32_t *out_dev, void *stream);

/* ---- prefix-sum search: replaces get_prefix_sum_idx, _sumtree_array.pyx:66-122 --- */
ST_API int st_search_device(const st_tree *t, const void *vals_dev, int64_t m, int64_t *out_dev,
                     void *stream);
ST_API int st_search_host(const st_tree *t, const void *vals_host, int64_t m, int64_t *out_host);
/* sample(): vals = (T)(double(root) * u) then search; replaces sumtree_array.py:417-421.
 * prio_out (nullable) receives the sampled leaves' values (same dtype as the tree). */
ST_API int st_sample_uniform_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *out_dev,
                             void *prio_out_dev, void *stream);
ST_API int st_sample_uniform_host(const st_tree *t, const double *u_host, int64_t m, int64_t *out_host);
/* same plus the root (sumtree[1]) the uniforms were scaled by: sample()'s zero-total check
 * (sumtree_array.py:417-418) without a second device round trip */
ST_API int st_sample_root_host(const st_tree *t, const double *u_host, int64_t m, int64_t *out_host,
                        void *root_out);

/* ---- sums: replace sum_over / strided_sum, _sumtree_array.pyx:131-194 ------------ */
ST_API int st_root_host(const st_tree *t, void *out_scalar); /* sumtree[1], sumtree_array.py:503 */
ST_API int st_sum_over_host(const st_tree *t, int64_t index_from, int64_t index_to, void *out_scalar);
/* nout must be ceil(n/stride) (ST_ERR_SIZE otherwise, pyx:178-179) */
ST_API int st_strided_sum_device(const st_tree *t, int64_t stride, void *out_dev, int64_t nout,
                          void *stream);
ST_API int st_strided_sum_host(const st_tree *t, int64_t stride, void *out_host, int64_t nout);
/* Non-contiguous axis sum (replaces the ndarray.sum fallback, sumtree_array.py:514):
 * leaves viewed as [A][R][C]; out[a][c] = ((x[a][0][c] + x[a][1][c]) + ...) in row order.
 * Output element type: the tree dtype for f32/f64, int64 for signed and uint64 for unsigned
 * integers (NumPy's default accumulator).  Requires A*R*C == n. */
ST_API int st_axis_fold_device(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out_dev,
                        void *stream);
ST_API int st_axis_fold_host(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out_host);

/* ---- PER step (reference README.md:7, users do this on the host through
 * sumtree_array.py:229-233, :365-421 and x[idx] / x.sum()) ----------------------------- */
/* st_sample_uniform_device that also returns, per sample, the leaf's priority (tree dtype) and the
 * sampling probability priority / total (tree dtype for f32 / f64 -- one IEEE division, as NumPy's
 * x[idx] / x.sum() -- float64 for integer trees).  Either output may be NULL. */
ST_API int st_sample_weights_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *out_dev,
                             void *prio_out_dev, void *prob_out_dev, void *stream);
/* update -> sample -> priorities -> probabilities as ONE call: enqueues only (capturable in a CUDA
 * graph after st_reserve); validation is deferred to st_poll_status. */
ST_API int st_per_step_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                       const double *u_dev, int64_t m, int64_t *out_idx_dev, void *out_prio_dev,
                       void *out_prob_dev, void *stream);
/* The same step for HOST buffers, pipelined over ST_HOST_SLOTS slots: submit(slot) enqueues H2D of
 * (idx, vals, u) on a copy stream, the step on the compute stream and D2H of the results on a third
 * stream, and returns; wait(slot) blocks until that slot's results are in the caller's buffers and
 * returns the batch's validation status.  Steps run in submission order.  With three or more slots in
 * flight (submit step k before waiting for step k - 2) the upload of step k+1 and the read-back of step
 * k-1 overlap the kernels of step k and the host never holds up the copy stream; with two, the next upload
 * starts only after the host has collected a result.  Buffers must stay valid until wait(); use
 * st_host_alloc (pinned memory) for truly asynchronous copies.  out_prio_host / out_prob_host may be NULL. */
#define ST_HOST_SLOTS 4
ST_API int st_per_step_host_submit(st_tree *t, int slot, const int64_t *idx_host, int64_t ni, const void *vals_host,
                            int64_t nv, const double *u_host, int64_t m, int64_t *out_idx_host,
                            void *out_prio_host, void *out_prob_host);
ST_API int st_per_step_host_wait(st_tree *t, int slot);
ST_API int st_host_alloc(int64_t bytes, void **out); /* page-locked host memory */
ST_API int st_host_free(void *p);

/* ---- in-place leaf arithmetic (replaces host ufunc + full rebuild upload,
 * reference sumtree_array.py:150-173, :202-205) ---------------------------------------- */
/* leaves = leaves (op) c, c = *scalar_host (one element of the tree dtype) or, if operand_dev is not
 * NULL, operand_dev[i] (operand_len must be n).  Bit-identical to the NumPy ufunc in the tree dtype.
 * Does NOT rebuild the tree: follow with st_build (which validates the new leaves, pyx:58-59). */
ST_API int st_inplace_device(st_tree *t, int op, const void *scalar_host, const void *operand_dev,
                      int64_t operand_len, void *stream);

/* ---- min tree + ring append (reference starr/experimental/slow.py:53-98) - */
/* A tree whose internal nodes are min(left, right); unset leaves are +inf (integers: type max).
 * st_root_host = MinTree.min(); st_store_host / st_gather_leaves_* / st_destroy work as usual; the
 * sum-tree entry points reject a min tree with ST_ERR_ARG. */
ST_API int st_min_create(int64_t n, int dtype, int device, void *external_heap, st_tree **out); /* as st_create */
/* heap[n + idx[j]] = vals[j % nv] for j in batch order (the last write to a leaf wins), then every
 * ancestor = min(left, right) (slow.py:62-68).  Index range is validated before any write. */
ST_API int st_min_update_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                         void *stream);
ST_API int st_min_update_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv);
ST_API int st_min_build(st_tree *t, void *stream); /* all internal nodes from the resident leaves */

/* ---- bookkeeping -------------------------------------------------------------------- */
/* out[0] = updates submitted through this handle, out[1] = how many of them had a difference of
 * exactly 0 (nothing to propagate); cumulative, reset != 0 clears both after reading.  Waits for stream. */
ST_API int st_update_counters(st_tree *t, void *stream, int64_t out[2], int reset);
/* st_reserve for the sharded helpers too: sizes the scratch area for st_shard_plan_device and the
 * top-fold records of batches up to max_batch over `shards` shards, so that none of the stream-ordered
 * entry points reallocates (or synchronises) later. */
ST_API int st_reserve_sharded(st_tree *t, int64_t max_batch, int shards);

/* ---- deferred validation status of *_device calls -------------------------------- */
/* Waits for `stream`, returns the first validation error recorded since the last clear
 * (ST_OK if none) and optionally clears it.  A failed batch is skipped as a whole; batches
 * enqueued behind it are applied normally (its flags move to a sticky word that this call reports). */
ST_API int st_poll_status(st_tree *t, void *stream, int clear);

/* Optional: pin heap[0, bytes) (the upper tree levels) in the persisting part of L2 for work
 * submitted to `stream` (cudaAccessPolicyWindow).  bytes <= 0 removes the window. */
ST_API int st_set_l2_persistence(st_tree *t, void *stream, int64_t bytes, int64_t *granted);

/* Per-phase timing of the float update pipeline, measured with CUDA events on the launching
 * stream.  While enabled every update call waits for its own completion (profiling only).
 * out_ms[0..6] = accumulated milliseconds of: key build + leaf-order sort, leaf pass, top-down
 * partition kernel, retirement of short segments, huge-segment summaries, huge-segment stitch,
 * ordered-fold kernel;
 * out_ms[7] = number of update chunks accumulated. */
ST_API int st_set_profiling(st_tree *t, int enable);
ST_API int st_update_breakdown(const st_tree *t, double out_ms[8]);

/* statistics of the most recent float update (for profiling; all zeros for int trees):
 * out[0]=segments resolved by the exact order-independent fast path,
 * out[1]=segments folded sequentially, out[2]=elements in those, out[3]=longest chain */
ST_API int st_update_stats(st_tree *t, void *stream, int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* STARR_B200_H */
