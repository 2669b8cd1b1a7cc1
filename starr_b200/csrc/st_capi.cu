// st_capi.cu -- the extern "C" surface declared in include/starr_b200.h: handle lifetime,
// host<->device staging for the *_host entry points, and status plumbing.  No kernels here.
#include <cstdio>
#include <cstring>
#include <new>

#include "st_common.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int status, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return status;
}

int cuda_fail(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return ST_ERR_CUDA;
}

#define CU(call)                                                \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
    } while (0)

const int kSizes[9] = {2, 4, 8, 1, 2, 4, 8, 4, 8};
constexpr int kPlanSlots = 4;            // result slots of st_shard_plan_device behind the pinned staging area
constexpr size_t kPlanSlotBytes = 1024;  // 64 per-shard counts + one flags word

int status_from_flags(unsigned int f)
{
    if (f & STF_NEG_VALS) return ST_ERR_NEGATIVE_VALS;
    if (f & STF_NEG_ARRAY) return ST_ERR_NEGATIVE_ARRAY;
    if (f & STF_INDEX) return ST_ERR_INDEX;
    return ST_OK;
}

struct device_guard {
    int prev = -1;
    bool ok = true;
    explicit device_guard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~device_guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Serialises the callers of one tree (the reference's functions hold the GIL end to end):
// staging areas, workspace, status queue and the legacy-stream host paths belong to one call at a time.
struct tree_lock {
    std::lock_guard<std::recursive_mutex> g;
    explicit tree_lock(const st_tree *t) : g(const_cast<st_tree *>(t)->mu) {}
};

cudaError_t clear_status_words(st_tree *t, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(t->d_status, 0, sizeof(unsigned int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->d_status + ST_SW_STICKY, 0, sizeof(unsigned int), s);
    return e;
}

// Wait for the stream and fetch the validation flags, optionally clearing what was reported.
// own_only: just the batch that was enqueued last (synchronous calls report their OWN verdict; an
// earlier asynchronous failure stays in the sticky word for st_poll_status).
int fetch_status(st_tree *t, cudaStream_t s, bool clear, bool own_only = false)
{
    CU(cudaMemcpyAsync(t->h_status, t->d_status, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const unsigned int f = own_only ? t->h_status[0] : (t->h_status[0] | t->h_status[ST_SW_STICKY]);
    if (f && clear) {
        if (own_only) CU(cudaMemsetAsync(t->d_status, 0, sizeof(unsigned int), s));
        else CU(clear_status_words(t, s));
        CU(cudaStreamSynchronize(s));
    }
    return status_from_flags(f);
}

// Staging area of the *_host entry points: carved out of one grow-only device buffer per tree so
// that a call costs copies + launches, not cudaMalloc/cudaFree.
struct dev_buf {
    void *p = nullptr;
};
struct scratch_carver {
    st_tree *t;
    size_t need = 0;
    explicit scratch_carver(const st_tree *tree) : t(const_cast<st_tree *>(tree)) {}
    static size_t up(size_t b) { return (b + 255) & ~(size_t)255; }
    // first pass: add(bytes) for every buffer; then commit(); then take(bytes) in the same order
    void add(size_t bytes) { need += up(bytes ? bytes : 1); }
    cudaError_t commit()
    {
        off = 0;
        if (need <= t->scratch.bytes) return cudaSuccess;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return e;
        if (t->scratch.base) cudaFree(t->scratch.base);
        t->scratch.base = nullptr;
        t->scratch.bytes = 0;
        const size_t want = need < (1u << 20) ? (1u << 20) : need;
        e = cudaMalloc(&t->scratch.base, want);
        if (e != cudaSuccess) return e;
        t->scratch.bytes = want;
        return cudaSuccess;
    }
    void *take(size_t bytes)
    {
        void *p = (char *)t->scratch.base + off;
        off += up(bytes ? bytes : 1);
        return p;
    }
    size_t off = 0;
};

}  // namespace

namespace st { std::atomic<unsigned long long> g_kernel_launches{0}; }

extern "C" {

int st_abi_version(void) { return 4; }   // 4: slices layout, sample push, ST_HOST_SLOTS = 4

unsigned long long st_kernel_launches(void) { return st::g_kernel_launches.load(); }

const char *st_status_string(int status)
{
    switch (status) {
        case ST_OK: return "ok";
        case ST_ERR_NEGATIVE_VALS: return "vals must be non-negative";
        case ST_ERR_NEGATIVE_ARRAY: return "array elements must be non-negative";
        case ST_ERR_INDEX: return "index out of bounds";
        case ST_ERR_EMPTY_VALS: return "integer division or modulo by zero";
        case ST_ERR_SIZE: return "size mismatch";
        case ST_ERR_DTYPE: return "No matching signature found";
        case ST_ERR_ARG: return "invalid argument";
        case ST_ERR_CUDA: return "CUDA error";
        case ST_ERR_TOO_LARGE: return "tree too large";
        default: return "unknown status";
    }
}

const char *st_last_error(void) { return g_err; }

int st_dtype_size(int dtype) { return (dtype >= 0 && dtype < 9) ? kSizes[dtype] : -1; }

int st_create(int64_t n, int dtype, int device, void *external_heap, st_tree **out)
{
    if (!out) return fail(ST_ERR_ARG, "st_create: out is NULL");
    *out = nullptr;
    if (dtype < 0 || dtype >= 9) return fail(ST_ERR_DTYPE, "st_create: unsupported dtype");
    if (n < 2) return fail(ST_ERR_ARG, "input to SumTreeArray must have shape with at least 2 elements");
    if (n > 0x7fffffffLL) return fail(ST_ERR_TOO_LARGE, "st_create: n exceeds 2^31-1 (32-bit heap index)");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ST_ERR_CUDA, "st_create: no CUDA device available (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(ST_ERR_ARG, "st_create: bad device ordinal");
    device_guard g(device);
    if (!g.ok) return fail(ST_ERR_CUDA, "st_create: cudaSetDevice failed");

    st_tree *t = new (std::nothrow) st_tree();
    if (!t) return fail(ST_ERR_ARG, "st_create: out of host memory");
    t->n = n;
    t->dtype = dtype;
    t->device = device;
    t->esize = kSizes[dtype];
    t->depth = st::heap_level((uint32_t)(2 * n - 1));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        t->sm_count = prop.multiProcessorCount;
        t->coop_ok = prop.cooperativeLaunch;
    }
    const size_t heap_bytes = (size_t)(2 * n) * t->esize;
    if (external_heap) {
        t->heap = external_heap;
        t->owns_heap = false;
    } else {
        e = cudaMalloc(&t->heap, heap_bytes);
        if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(heap)"); }
        t->owns_heap = true;
    }
    e = cudaMalloc((void **)&t->d_status, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&t->h_status, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) {
        t->pin_bytes = 64 * 1024;
        e = cudaHostAlloc(&t->h_pin, t->pin_bytes + kPlanSlots * kPlanSlotBytes, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&t->d_pin, t->h_pin, 0);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->plan_ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_hy, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) {
        // search index (st_sidx.cuh): all zeros matches the all-zero tree
        t->sidx_bytes = st::sidx_bytes_for(t->n, t->depth, t->esize, t->kind, &t->sidx_nb);
        if (t->sidx_bytes) {
            e = cudaMalloc(&t->sidx, t->sidx_bytes);
            if (e == cudaSuccess) e = cudaMemset(t->sidx, 0, t->sidx_bytes);
        }
    }
    if (e == cudaSuccess) e = cudaMemset(t->d_status, 0, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(t->heap, 0, heap_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        st_destroy(t);
        return cuda_fail(e, "st_create");
    }
    *out = t;
    return ST_OK;
}

int st_destroy(st_tree *t)
{
    if (!t) return ST_OK;
    device_guard g(t->device);
    if (t->owns_heap && t->heap) cudaFree(t->heap);
    if (t->d_status) cudaFree(t->d_status);
    if (t->h_status) cudaFreeHost(t->h_status);
    if (t->h_pin) cudaFreeHost(t->h_pin);
    if (t->plan_ev) cudaEventDestroy(t->plan_ev);
    if (t->sidx) cudaFree(t->sidx);
    if (t->ev_fork) cudaEventDestroy(t->ev_fork);
    if (t->ev_join) cudaEventDestroy(t->ev_join);
    if (t->ev_hy) cudaEventDestroy(t->ev_hy);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->ws.base) cudaFree(t->ws.base);
    if (t->scratch.base) cudaFree(t->scratch.base);
    if (t->aux) cudaFree(t->aux);
    for (auto &h : t->hs) {
        if (h.dev) cudaFree(h.dev);
        if (h.h_flag) cudaFreeHost(h.h_flag);
        for (cudaEvent_t ev : {h.ev_in, h.ev_free, h.ev_done, h.ev_read})
            if (ev) cudaEventDestroy(ev);
    }
    for (cudaStream_t st : {t->hs_in, t->hs_main, t->hs_out})
        if (st) cudaStreamDestroy(st);
    for (int k = 0; k < 8; ++k)
        if (t->prof_ev[k]) cudaEventDestroy(t->prof_ev[k]);
    delete t;
    return ST_OK;
}

int64_t st_size(const st_tree *t) { return t ? t->n : -1; }
int st_dtype_of(const st_tree *t) { return t ? t->dtype : -1; }
int st_device_of(const st_tree *t) { return t ? t->device : -1; }
void *st_heap_ptr(const st_tree *t) { return t ? t->heap : nullptr; }

int st_reserve(st_tree *t, int64_t max_batch)
{
    if (!t || max_batch < 1) return fail(ST_ERR_ARG, "st_reserve: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    if (max_batch > st::update_max_chunk()) max_batch = st::update_max_chunk();
    if (max_batch <= t->reserved_batch && t->ws.base) return ST_OK;
    const size_t need = st::update_workspace_bytes(t, max_batch);
    CU(cudaDeviceSynchronize());
    if (t->ws.base) cudaFree(t->ws.base);
    t->ws.base = nullptr;
    t->ws.bytes = 0;
    CU(cudaMalloc(&t->ws.base, need));
    t->ws.bytes = need;
    t->reserved_batch = max_batch;
    return ST_OK;
}

// ---- raw state transfer ------------------------------------------------------------------
int st_load_host(st_tree *t, const void *leaves, const void *tree)
{
    if (!t) return fail(ST_ERR_ARG, "st_load_host: NULL tree");
    tree_lock lk(t);
    device_guard g(t->device);
    const size_t half = (size_t)t->n * t->esize;
    if (tree) CU(cudaMemcpy(t->heap, tree, half, cudaMemcpyHostToDevice));
    if (leaves) CU(cudaMemcpy((char *)t->heap + half, leaves, half, cudaMemcpyHostToDevice));
    if (tree && t->sidx) {   // the search index follows the tree that was just loaded
        CU(cudaMemsetAsync(t->d_status, 0, sizeof(unsigned int), 0));
        CU(st::launch_sidx_build(t, 0));
        CU(cudaStreamSynchronize(0));
    }
    return ST_OK;
}

int st_store_host(const st_tree *t, void *leaves, void *tree)
{
    if (!t) return fail(ST_ERR_ARG, "st_store_host: NULL tree");
    tree_lock lk(t);
    device_guard g(t->device);
    const size_t half = (size_t)t->n * t->esize;
    if (tree) CU(cudaMemcpy(tree, t->heap, half, cudaMemcpyDeviceToHost));
    if (leaves) CU(cudaMemcpy(leaves, (const char *)t->heap + half, half, cudaMemcpyDeviceToHost));
    return ST_OK;
}

int st_gather_leaves_device(const st_tree *t, const int64_t *idx_dev, int64_t m, void *out_dev, void *stream)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_gather_leaves_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_gather(t, idx_dev, m, out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_gather_leaves_host(const st_tree *t, const int64_t *idx_host, int64_t m, void *out_host)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_gather_leaves_host: bad argument");
    if (m == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf di, dout;
    scratch_carver sc(t);
    sc.add(m * sizeof(int64_t)); sc.add(m * t->esize);
    CU(sc.commit());
    di.p = sc.take(m * sizeof(int64_t)); dout.p = sc.take(m * t->esize);
    CU(cudaMemcpy(di.p, idx_host, m * sizeof(int64_t), cudaMemcpyHostToDevice));
    CU(st::launch_gather(t, (const int64_t *)di.p, m, dout.p, 0));
    CU(cudaMemcpy(out_host, dout.p, m * t->esize, cudaMemcpyDeviceToHost));
    return ST_OK;
}

// ---- build ---------------------------------------------------------------------------------
int st_build(st_tree *t, void *stream)
{
    if (!t || t->kind != 0) return fail(ST_ERR_ARG, "st_build: NULL tree (or a min tree: use st_min_build)");
    tree_lock lk(t);
    device_guard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    CU(st::launch_batch_begin(t, s));   // an earlier failed batch must not cancel this build
    CU(st::launch_check_leaves(t, s));
    CU(st::launch_build(t, s));
    const int st = fetch_status(t, s, true, true);
    if (st != ST_OK) return fail(st, "%s", st_status_string(st));
    return ST_OK;
}

int st_build_from_host(st_tree *t, const void *leaves)
{
    if (!t || !leaves) return fail(ST_ERR_ARG, "st_build_from_host: NULL argument");
    tree_lock lk(t);
    device_guard g(t->device);
    const size_t half = (size_t)t->n * t->esize;
    CU(cudaMemcpy((char *)t->heap + half, leaves, half, cudaMemcpyHostToDevice));
    return st_build(t, nullptr);
}

int st_build_from_device(st_tree *t, const void *leaves_dev, void *stream)
{
    if (!t || !leaves_dev) return fail(ST_ERR_ARG, "st_build_from_device: NULL argument");
    tree_lock lk(t);
    device_guard g(t->device);
    const size_t half = (size_t)t->n * t->esize;
    void *dst = (char *)t->heap + half;
    if (dst != leaves_dev)
        CU(cudaMemcpyAsync(dst, leaves_dev, half, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return st_build(t, stream);
}

// ---- update --------------------------------------------------------------------------------
int st_update_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                     void *stream, int flags)
{
    if (!t || ni < 0 || nv < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_update_device: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    tree_lock lk(t);
    device_guard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (ni == 0 && nv == 0) return ST_OK;
    CU(st::launch_update(t, idx_dev, ni, vals_dev, nv, nullptr, nullptr, s, (flags & ST_UPDATE_ASSIGN_LEAF) != 0));
    if (flags & ST_UPDATE_SYNC_CHECK) {
        const int st = fetch_status(t, s, true, true);
        if (st != ST_OK) return fail(st, "%s", st_status_string(st));
    }
    return ST_OK;
}

int st_update_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv)
{
    return st_update_host_ex(t, idx_host, ni, vals_host, nv, ST_UPDATE_DEFAULT);
}

int st_update_host_ex(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv, int flags)
{
    if (!t || ni < 0 || nv < 0) return fail(ST_ERR_ARG, "st_update_host: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    if (ni == 0 && nv == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf di, dv;
    scratch_carver sc(t);
    sc.add(ni * sizeof(int64_t)); sc.add(nv * t->esize);
    CU(sc.commit());
    di.p = sc.take(ni * sizeof(int64_t)); dv.p = sc.take(nv * t->esize);
    CU(cudaMemcpy(di.p, idx_host, ni * sizeof(int64_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dv.p, vals_host, nv * t->esize, cudaMemcpyHostToDevice));
    return st_update_device(t, (const int64_t *)di.p, ni, dv.p, nv, nullptr, flags | ST_UPDATE_SYNC_CHECK);
}

// ---- sharded-tree helpers ------------------------------------------------------
int st_update_deltas_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                            void *delta_out_dev, void *stream)
{
    if (!t || ni < 0 || nv < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_update_deltas_device: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    if (ni == 0 && nv == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_update(t, idx_dev, ni, vals_dev, nv, delta_out_dev, nullptr, (cudaStream_t)stream));
    return ST_OK;
}

int st_update_begin_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                           void *delta_out_dev, void *stream)
{
    if (!t || ni < 0 || nv < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_update_begin_device: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    tree_lock lk(t);
    device_guard g(t->device);
    t->pending_split = 0;
    if (ni == 0 && nv == 0) return ST_OK;
    CU(st::launch_update(t, idx_dev, ni, vals_dev, nv, delta_out_dev, nullptr, (cudaStream_t)stream, false, 1));
    return ST_OK;
}

int st_update_finish_device(st_tree *t, void *stream)
{
    if (!t || t->kind != 0) return fail(ST_ERR_ARG, "st_update_finish_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_update(t, nullptr, 0, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream, false, 2));
    return ST_OK;
}

int st_apply_deltas_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *deltas_dev, void *stream)
{
    if (!t || ni < 0 || (ni > 0 && !deltas_dev)) return fail(ST_ERR_ARG, "st_apply_deltas_device: bad argument");
    if (ni == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    const bool is_float = (t->dtype == ST_F32 || t->dtype == ST_F64);
    if (is_float && t->n <= 64 && (t->n & (t->n - 1)) == 0 && ni > 4096) {
        // tiny tree, big batch (the sharded top tree): masked ordered fold per node, no sort / partition.
        // Indices are trusted to be in range here (they are shard ids computed by the caller).
        CU(st::launch_fold_by_leaf(t, idx_dev, deltas_dev, ni, (cudaStream_t)stream));
        return ST_OK;
    }
    CU(st::launch_update(t, idx_dev, ni, nullptr, 0, nullptr, deltas_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_top_fold_record_bytes(const st_tree *t) { return t ? (int)st::fold_rec_bytes(t) : -1; }

int st_top_fold_maps_device(st_tree *t, const void *leaf_dev, int leaf_itemsize, const void *deltas_dev, int64_t m,
                            int64_t chunk_lo, int64_t chunk_hi, void *recs_dev, void *stream)
{
    if (!t || m < 0 || !recs_dev || (leaf_itemsize != 1 && leaf_itemsize != 8))
        return fail(ST_ERR_ARG, "st_top_fold_maps_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_fold_by_leaf_maps(t, leaf_dev, leaf_itemsize, deltas_dev, m, (int)chunk_lo, (int)chunk_hi, recs_dev,
                                    (cudaStream_t)stream));
    return ST_OK;
}

int st_top_fold_apply_device(st_tree *t, const void *leaf_dev, int leaf_itemsize, const void *deltas_dev, int64_t m,
                             const void *recs_dev, void *stream)
{
    if (!t || m < 0 || !recs_dev || (leaf_itemsize != 1 && leaf_itemsize != 8))
        return fail(ST_ERR_ARG, "st_top_fold_apply_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_fold_by_leaf_apply(t, leaf_dev, leaf_itemsize, deltas_dev, m, recs_dev, (cudaStream_t)stream));
    return ST_OK;
}

namespace {
int plan_result(st_tree *t, int slot, int shards, int64_t *counts_out_host)
{
    const unsigned int *pin = (const unsigned int *)((const char *)t->h_pin + t->pin_bytes + slot * kPlanSlotBytes);
    for (int i = 0; i < shards; ++i) counts_out_host[i] = (int64_t)pin[i];
    const int status = status_from_flags(pin[shards]);
    return status == ST_OK ? ST_OK : fail(status, "%s", st_status_string(status));
}
}  // namespace

int st_shard_plan_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards, int rank,
                         const void *payload_dev, int64_t npayload, int check_negative, uint8_t *shard_out_dev,
                         int32_t *slot_out_dev, int64_t *own_ids_out_dev, void *own_payload_out_dev, int result_slot,
                         int64_t *counts_out_host, void *stream)
{
    return st_shard_plan_pos_device(t, ids_dev, m, shift, shards, rank, payload_dev, npayload, check_negative, shard_out_dev,
                                    slot_out_dev, own_ids_out_dev, own_payload_out_dev, nullptr, result_slot, counts_out_host,
                                    stream);
}

int st_shard_plan_pos_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards, int rank,
                             const void *payload_dev, int64_t npayload, int check_negative, uint8_t *shard_out_dev,
                             int32_t *slot_out_dev, int64_t *own_ids_out_dev, void *own_payload_out_dev,
                             int32_t *own_pos_out_dev, int result_slot, int64_t *counts_out_host, void *stream)
{
    if (!t || m <= 0 || shards < 1 || shards > 64 || rank < 0 || rank >= shards || shift < 0 || shift > 62 ||
        result_slot < 0 || result_slot >= kPlanSlots || !ids_dev || !shard_out_dev || !slot_out_dev ||
        ((own_payload_out_dev || check_negative) && (!payload_dev || npayload <= 0)))
        return fail(ST_ERR_ARG, "st_shard_plan_device: bad argument");
    if (((uintptr_t)shard_out_dev & 7) || ((uintptr_t)slot_out_dev & 15))
        return fail(ST_ERR_ARG, "st_shard_plan_device: shard_out must be 8-byte and slot_out 16-byte aligned");
    tree_lock lk(t);
    device_guard g(t->device);
    if (t->plan_pending) {   // the scratch area belongs to one plan at a time
        CU(cudaEventSynchronize(t->plan_ev));
        t->plan_pending = 0;
    }
    scratch_carver c(t);
    const size_t bytes = st::shard_plan_scratch_bytes(m, shards);
    c.add(bytes);
    CU(c.commit());
    void *scratch = c.take(bytes);
    unsigned int *counts_dev = (unsigned int *)((char *)t->d_pin + t->pin_bytes + result_slot * kPlanSlotBytes);
    if (m > 0x7fffffffLL && own_pos_out_dev) return fail(ST_ERR_TOO_LARGE, "st_shard_plan_pos_device: batch positions are 32-bit");
    CU(st::launch_shard_plan(t, ids_dev, m, shift, shards, rank, payload_dev, npayload, check_negative, shard_out_dev,
                             slot_out_dev, own_ids_out_dev, own_payload_out_dev, scratch, counts_dev,
                             (cudaStream_t)stream, own_pos_out_dev));
    CU(cudaEventRecord(t->plan_ev, (cudaStream_t)stream));
    t->plan_pending = 1;
    if (!counts_out_host) return ST_OK;   // asynchronous: st_shard_plan_wait collects the result
    return st_shard_plan_wait(t, result_slot, shards, counts_out_host);
}

int st_shard_partition_device(st_tree *t, const int64_t *ids_dev, int64_t m, int shift, int shards, const void *payload_dev,
                              int check_negative, uint8_t *shard_out_dev, int32_t *slot_out_dev, int64_t *part_ids_out_dev,
                              void *part_payload_out_dev, uint32_t *row_out_dev, void *stream)
{
    if (!t || m <= 0 || m > 0x7fffffffLL || shards < 1 || shards > 64 || shift < 0 || shift > 62 || !ids_dev || !shard_out_dev ||
        !slot_out_dev || !row_out_dev || ((part_payload_out_dev || check_negative) && !payload_dev))
        return fail(ST_ERR_ARG, "st_shard_partition_device: bad argument");
    if (((uintptr_t)shard_out_dev & 7) || ((uintptr_t)slot_out_dev & 15))
        return fail(ST_ERR_ARG, "st_shard_partition_device: shard_out must be 8-byte and slot_out 16-byte aligned");
    tree_lock lk(t);
    device_guard g(t->device);
    if (t->plan_pending) {   // the scratch area belongs to one plan at a time
        CU(cudaEventSynchronize(t->plan_ev));
        t->plan_pending = 0;
    }
    scratch_carver c(t);
    const size_t bytes = st::shard_plan_scratch_bytes(m, shards);
    c.add(bytes);
    CU(c.commit());
    void *scratch = c.take(bytes);
    CU(st::launch_shard_partition(t, ids_dev, m, shift, shards, payload_dev, check_negative, shard_out_dev, slot_out_dev,
                                  part_ids_out_dev, part_payload_out_dev, scratch, row_out_dev, (cudaStream_t)stream));
    CU(cudaEventRecord(t->plan_ev, (cudaStream_t)stream));
    t->plan_pending = 1;
    return ST_OK;
}

int st_shard_plan_wait(st_tree *t, int result_slot, int shards, int64_t *counts_out_host)
{
    if (!t || !counts_out_host || shards < 1 || shards > 64 || result_slot < 0 || result_slot >= kPlanSlots)
        return fail(ST_ERR_ARG, "st_shard_plan_wait: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    if (t->plan_pending) {
        CU(cudaEventSynchronize(t->plan_ev));   // the plan's kernels only, not what was enqueued behind them
        t->plan_pending = 0;
    }
    return plan_result(t, result_slot, shards, counts_out_host);
}

const uint32_t *st_shard_plan_counts_device(const st_tree *t, int result_slot)
{
    if (!t || result_slot < 0 || result_slot >= kPlanSlots) return nullptr;
    return (const uint32_t *)((const char *)t->d_pin + t->pin_bytes + result_slot * kPlanSlotBytes);
}

int st_search_counted_device(const st_tree *t, const void *vals_dev, int64_t m_max, const uint32_t *m_dev,
                             int64_t *out_dev, void *stream)
{
    if (!t || m_max < 0 || !m_dev) return fail(ST_ERR_ARG, "st_search_counted_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_search(t, vals_dev, m_max, out_dev, (cudaStream_t)stream, m_dev));
    return ST_OK;
}

int st_shard_expand_values_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                                  const void *slabs_dev, int64_t slab_stride, void *out_dev, void *stream)
{
    if (!t || m < 0 || slab_stride < 0 || (m > 0 && (!shard_dev || !slot_dev || !slabs_dev || !out_dev)))
        return fail(ST_ERR_ARG, "st_shard_expand_values_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_shard_expand_values(t, shard_dev, slot_dev, m, slabs_dev, slab_stride, out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_shard_expand_cat_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                               const void *cat_dev, const int64_t *starts_host, int shards, void *out_dev, void *stream)
{
    if (!t || m < 0 || shards < 1 || shards > 64 || !starts_host || (m > 0 && (!shard_dev || !slot_dev || !cat_dev || !out_dev)))
        return fail(ST_ERR_ARG, "st_shard_expand_cat_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_shard_expand_cat(t, shard_dev, slot_dev, m, cat_dev, starts_host, shards, out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_shard_expand_indices_device(const st_tree *t, const uint8_t *shard_dev, const int32_t *slot_dev, int64_t m,
                                   const int32_t *slabs_dev, int64_t slab_stride, int64_t leaves_per_shard,
                                   int64_t *out_dev, void *stream)
{
    if (!t || m < 0 || slab_stride < 0 || (m > 0 && (!shard_dev || !slot_dev || !slabs_dev || !out_dev)))
        return fail(ST_ERR_ARG, "st_shard_expand_indices_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_shard_expand_indices(t, shard_dev, slot_dev, m, slabs_dev, slab_stride, leaves_per_shard, out_dev,
                                       (cudaStream_t)stream));
    return ST_OK;
}

int st_route_uniform_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *leaf_out_dev,
                            void *residual_out_dev, void *stream)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_route_uniform_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_sample(t, u_dev, m, leaf_out_dev, nullptr, residual_out_dev, (cudaStream_t)stream));
    return ST_OK;
}

// put(): update, then read back the leaves that were written, one staging round (sumtree_array.py:229-233)
int st_update_gather_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv,
                          void *leaves_out_host)
{
    return st_update_gather_root_host(t, idx_host, ni, vals_host, nv, leaves_out_host, nullptr);
}

int st_update_gather_root_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv,
                               void *leaves_out_host, void *root_out_host)
{
    if (!t || ni < 0 || nv < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_update_gather_host: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    if (ni == 0 && nv == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    {
        // small call: everything through mapped pinned memory, one stream sync, no cudaMemcpy
        const size_t up = 256;
        auto rnd = [&](size_t b) { return (b + up - 1) / up * up; };
        const size_t o_idx = 0, o_val = o_idx + rnd(ni * sizeof(int64_t)), o_leaf = o_val + rnd(nv * t->esize);
        const size_t o_stat = o_leaf + rnd(ni * t->esize), o_root = o_stat + up, total = o_root + up;
        if (total <= t->pin_bytes) {
            char *hp = (char *)t->h_pin, *dp = (char *)t->d_pin;
            memcpy(hp + o_idx, idx_host, ni * sizeof(int64_t));
            memcpy(hp + o_val, vals_host, nv * t->esize);
            CU(st::launch_update(t, (const int64_t *)(dp + o_idx), ni, dp + o_val, nv, nullptr, nullptr, 0));
            if (ni && leaves_out_host) CU(st::launch_gather(t, (const int64_t *)(dp + o_idx), ni, dp + o_leaf, 0));
            CU(st::launch_publish(t, dp + o_stat, root_out_host ? dp + o_root : nullptr, 0));   // flags (+ the new root)
            CU(cudaStreamSynchronize(0));
            if (root_out_host) memcpy(root_out_host, hp + o_root, t->esize);
            const unsigned int f = *(volatile unsigned int *)(hp + o_stat);
            if (f) {
                CU(clear_status_words(t, 0));
                CU(cudaStreamSynchronize(0));
                const int st = status_from_flags(f);
                return fail(st, "%s", st_status_string(st));
            }
            if (ni && leaves_out_host) memcpy(leaves_out_host, hp + o_leaf, ni * t->esize);
            return ST_OK;
        }
    }
    scratch_carver sc(t);
    sc.add(ni * sizeof(int64_t)); sc.add(nv * t->esize); sc.add(ni * t->esize);
    CU(sc.commit());
    int64_t *di = (int64_t *)sc.take(ni * sizeof(int64_t));
    void *dv = sc.take(nv * t->esize);
    void *dl = sc.take(ni * t->esize);
    if (ni) CU(cudaMemcpyAsync(di, idx_host, ni * sizeof(int64_t), cudaMemcpyHostToDevice, 0));
    CU(cudaMemcpyAsync(dv, vals_host, nv * t->esize, cudaMemcpyHostToDevice, 0));
    CU(st::launch_update(t, di, ni, dv, nv, nullptr, nullptr, 0));
    if (ni && leaves_out_host) CU(st::launch_gather(t, di, ni, dl, 0));
    CU(cudaMemcpyAsync(t->h_status, t->d_status, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, 0));
    if (ni && leaves_out_host) CU(cudaMemcpyAsync(leaves_out_host, dl, ni * t->esize, cudaMemcpyDeviceToHost, 0));
    if (root_out_host) CU(cudaMemcpyAsync(root_out_host, (const char *)t->heap + t->esize, t->esize, cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    const unsigned int f = t->h_status[0] | t->h_status[ST_SW_STICKY];
    if (f) {
        CU(clear_status_words(t, 0));
        CU(cudaStreamSynchronize(0));
        const int st = status_from_flags(f);
        return fail(st, "%s", st_status_string(st));
    }
    return ST_OK;
}

// sample(): indices for the uniforms plus the root they were scaled by, one staging round
int st_sample_root_host(const st_tree *t, const double *u_host, int64_t m, int64_t *out_host, void *root_out)
{
    if (!t || m < 0 || !root_out) return fail(ST_ERR_ARG, "st_sample_root_host: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    {
        const size_t up = 256;
        auto rnd = [&](size_t b) { return (b + up - 1) / up * up; };
        const size_t o_u = 0, o_out = o_u + rnd(m * sizeof(double)), o_root = o_out + rnd(m * sizeof(int64_t));
        if (o_root + up <= t->pin_bytes) {
            char *hp = (char *)t->h_pin, *dp = (char *)t->d_pin;
            memcpy(hp + o_u, u_host, m * sizeof(double));
            if (m) CU(st::launch_sample(t, (const double *)(dp + o_u), m, (int64_t *)(dp + o_out), nullptr, nullptr, 0));
            CU(st::launch_publish(t, nullptr, dp + o_root, 0));
            CU(cudaStreamSynchronize(0));
            memcpy(out_host, hp + o_out, m * sizeof(int64_t));
            memcpy(root_out, hp + o_root, t->esize);
            return ST_OK;
        }
    }
    scratch_carver sc(t);
    sc.add(m * sizeof(double)); sc.add(m * sizeof(int64_t));
    CU(sc.commit());
    double *du = (double *)sc.take(m * sizeof(double));
    int64_t *dout = (int64_t *)sc.take(m * sizeof(int64_t));
    if (m) {
        CU(cudaMemcpyAsync(du, u_host, m * sizeof(double), cudaMemcpyHostToDevice, 0));
        CU(st::launch_sample(t, du, m, dout, nullptr, nullptr, 0));
        CU(cudaMemcpyAsync(out_host, dout, m * sizeof(int64_t), cudaMemcpyDeviceToHost, 0));
    }
    CU(cudaMemcpyAsync(root_out, (const char *)t->heap + t->esize, t->esize, cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    return ST_OK;
}

// ---- search / sample -----------------------------------------------------------------------
int st_search_device(const st_tree *t, const void *vals_dev, int64_t m, int64_t *out_dev, void *stream)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_search_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_search(t, vals_dev, m, out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_search_host(const st_tree *t, const void *vals_host, int64_t m, int64_t *out_host)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_search_host: bad argument");
    if (m == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    {
        const size_t o_out = ((size_t)m * t->esize + 255) / 256 * 256;
        if (o_out + (size_t)m * sizeof(int64_t) <= t->pin_bytes) {
            char *hp = (char *)t->h_pin, *dp = (char *)t->d_pin;
            memcpy(hp, vals_host, (size_t)m * t->esize);
            CU(st::launch_search(t, dp, m, (int64_t *)(dp + o_out), 0));
            CU(cudaStreamSynchronize(0));
            memcpy(out_host, hp + o_out, (size_t)m * sizeof(int64_t));
            return ST_OK;
        }
    }
    dev_buf dv, dout;
    scratch_carver sc(t);
    sc.add(m * t->esize); sc.add(m * sizeof(int64_t));
    CU(sc.commit());
    dv.p = sc.take(m * t->esize); dout.p = sc.take(m * sizeof(int64_t));
    CU(cudaMemcpy(dv.p, vals_host, m * t->esize, cudaMemcpyHostToDevice));
    CU(st::launch_search(t, dv.p, m, (int64_t *)dout.p, 0));
    CU(cudaMemcpy(out_host, dout.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return ST_OK;
}

int st_sample_uniform_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *out_dev,
                             void *prio_out_dev, void *stream)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_sample_uniform_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_sample(t, u_dev, m, out_dev, prio_out_dev, nullptr, (cudaStream_t)stream));
    return ST_OK;
}

int st_sample_uniform_host(const st_tree *t, const double *u_host, int64_t m, int64_t *out_host)
{
    if (!t || m < 0) return fail(ST_ERR_ARG, "st_sample_uniform_host: bad argument");
    if (m == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf du, dout;
    scratch_carver sc(t);
    sc.add(m * sizeof(double)); sc.add(m * sizeof(int64_t));
    CU(sc.commit());
    du.p = sc.take(m * sizeof(double)); dout.p = sc.take(m * sizeof(int64_t));
    CU(cudaMemcpy(du.p, u_host, m * sizeof(double), cudaMemcpyHostToDevice));
    CU(st::launch_sample(t, (const double *)du.p, m, (int64_t *)dout.p, nullptr, nullptr, 0));
    CU(cudaMemcpy(out_host, dout.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return ST_OK;
}

// ---- sums ----------------------------------------------------------------------------------
int st_root_host(const st_tree *t, void *out_scalar)
{
    if (!t || !out_scalar) return fail(ST_ERR_ARG, "st_root_host: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_publish(t, nullptr, t->d_pin, 0));
    CU(cudaStreamSynchronize(0));
    memcpy(out_scalar, t->h_pin, t->esize);
    return ST_OK;
}

int st_sum_over_host(const st_tree *t, int64_t index_from, int64_t index_to, void *out_scalar)
{
    if (!t || !out_scalar) return fail(ST_ERR_ARG, "st_sum_over_host: bad argument");
    // Cython boundscheck: the reference raises IndexError for positions outside the array
    if (index_from < 0 || index_from > t->n || index_to < 0 || index_to > t->n)
        return fail(ST_ERR_INDEX, "%s", st_status_string(ST_ERR_INDEX));
    if (index_to - 1 == index_from && index_from >= t->n) return fail(ST_ERR_INDEX, "%s", st_status_string(ST_ERR_INDEX));
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf d;
    scratch_carver sc(t);
    sc.add(8);
    CU(sc.commit());
    d.p = sc.take(8);
    CU(st::launch_sum_over(t, index_from, index_to, d.p, 0));
    CU(cudaMemcpy(out_scalar, d.p, t->esize, cudaMemcpyDeviceToHost));
    return ST_OK;
}

static int64_t strided_len(int64_t n, int64_t stride) { return n / stride + (n % stride ? 1 : 0); }

int st_strided_sum_device(const st_tree *t, int64_t stride, void *out_dev, int64_t nout, void *stream)
{
    if (!t || stride < 1) return fail(ST_ERR_ARG, "st_strided_sum_device: bad argument");
    if (nout != strided_len(t->n, stride)) return fail(ST_ERR_SIZE, "invalid output size");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_strided_sum(t, stride, out_dev, nout, (cudaStream_t)stream));
    return ST_OK;
}

int st_strided_sum_host(const st_tree *t, int64_t stride, void *out_host, int64_t nout)
{
    if (!t || stride < 1) return fail(ST_ERR_ARG, "st_strided_sum_host: bad argument");
    if (nout != strided_len(t->n, stride)) return fail(ST_ERR_SIZE, "invalid output size");
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf d;
    scratch_carver sc(t);
    sc.add(nout * t->esize);
    CU(sc.commit());
    d.p = sc.take(nout * t->esize);
    CU(st::launch_strided_sum(t, stride, d.p, nout, 0));
    CU(cudaMemcpy(out_host, d.p, nout * t->esize, cudaMemcpyDeviceToHost));
    return ST_OK;
}

static int fold_out_size(const st_tree *t) { return (t->dtype == ST_F32) ? 4 : 8; }

int st_axis_fold_device(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out_dev, void *stream)
{
    if (!t || A < 1 || R < 1 || C < 1) return fail(ST_ERR_ARG, "st_axis_fold_device: bad argument");
    if (A * R * C != t->n) return fail(ST_ERR_SIZE, "axis fold shape does not match the tree size");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_axis_fold(t, A, R, C, out_dev, (cudaStream_t)stream));
    return ST_OK;
}

int st_axis_fold_host(const st_tree *t, int64_t A, int64_t R, int64_t C, void *out_host)
{
    if (!t || A < 1 || R < 1 || C < 1) return fail(ST_ERR_ARG, "st_axis_fold_host: bad argument");
    if (A * R * C != t->n) return fail(ST_ERR_SIZE, "axis fold shape does not match the tree size");
    tree_lock lk(t);
    device_guard g(t->device);
    dev_buf d;
    const size_t bytes = (size_t)A * C * fold_out_size(t);
    scratch_carver sc(t);
    sc.add(bytes);
    CU(sc.commit());
    d.p = sc.take(bytes);
    CU(st::launch_axis_fold(t, A, R, C, d.p, 0));
    CU(cudaMemcpy(out_host, d.p, bytes, cudaMemcpyDeviceToHost));
    return ST_OK;
}

// ---- deferred status -----------------------------------------------------------------------
int st_poll_status(st_tree *t, void *stream, int clear)
{
    if (!t) return fail(ST_ERR_ARG, "st_poll_status: NULL tree");
    tree_lock lk(t);
    device_guard g(t->device);
    const int st = fetch_status(t, (cudaStream_t)stream, clear != 0);
    if (st != ST_OK) return fail(st, "%s", st_status_string(st));
    return ST_OK;
}

// Keep the top of the heap (the levels every query and every update walk through) resident in L2:
// reserve persisting L2 and attach an access-policy window over heap[0, bytes) to `stream`.
// bytes <= 0 removes the window.  Returns the window size actually set through *granted.
int st_set_l2_persistence(st_tree *t, void *stream, int64_t bytes, int64_t *granted)
{
    if (!t) return fail(ST_ERR_ARG, "st_set_l2_persistence: NULL tree");
    tree_lock lk(t);
    device_guard g(t->device);
    if (granted) *granted = 0;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, t->device));
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (bytes <= 0) {
        attr.accessPolicyWindow.num_bytes = 0;
        CU(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        CU(cudaCtxResetPersistingL2Cache());
        return ST_OK;
    }
    size_t want = (size_t)bytes;
    const size_t heap_bytes = (size_t)(2 * t->n) * t->esize;
    if (want > heap_bytes) want = heap_bytes;
    if (want > (size_t)prop.accessPolicyMaxWindowSize) want = (size_t)prop.accessPolicyMaxWindowSize;
    size_t carve = want;
    if (carve > (size_t)prop.persistingL2CacheMaxSize) carve = (size_t)prop.persistingL2CacheMaxSize;
    if (carve == 0) return fail(ST_ERR_CUDA, "st_set_l2_persistence: device has no persisting L2");
    CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    attr.accessPolicyWindow.base_ptr = t->heap;
    attr.accessPolicyWindow.num_bytes = want;
    attr.accessPolicyWindow.hitRatio = (float)((double)carve / (double)want);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (granted) *granted = (int64_t)want;
    return ST_OK;
}

int st_set_profiling(st_tree *t, int enable)
{
    if (!t) return fail(ST_ERR_ARG, "st_set_profiling: NULL tree");
    tree_lock lk(t);
    device_guard g(t->device);
    if (enable && !t->prof_ev[0])
        for (int k = 0; k < 8; ++k) CU(cudaEventCreate(&t->prof_ev[k]));
    t->profile = enable ? 1 : 0;
    for (int k = 0; k < 8; ++k) t->prof_ms[k] = 0.0;
    return ST_OK;
}

int st_update_breakdown(const st_tree *t, double out_ms[8])
{
    if (!t || !out_ms) return fail(ST_ERR_ARG, "st_update_breakdown: bad argument");
    for (int k = 0; k < 8; ++k) out_ms[k] = t->prof_ms[k];
    return ST_OK;
}

int st_update_stats(st_tree *t, void *stream, int64_t out[4])
{
    if (!t || !out) return fail(ST_ERR_ARG, "st_update_stats: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(t->h_status, t->d_status, 64 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int i = 0; i < 4; ++i) out[i] = t->h_status[4 + i];
    return ST_OK;
}

// ---- PER step ----------------------------------------------------------------------------------
int st_sample_weights_device(const st_tree *t, const double *u_dev, int64_t m, int64_t *out_dev,
                             void *prio_out_dev, void *prob_out_dev, void *stream)
{
    if (!t || m < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_sample_weights_device: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_sample(t, u_dev, m, out_dev, prio_out_dev, nullptr, (cudaStream_t)stream, prob_out_dev));
    return ST_OK;
}

int st_per_step_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv,
                       const double *u_dev, int64_t m, int64_t *out_idx_dev, void *out_prio_dev,
                       void *out_prob_dev, void *stream)
{
    if (!t || ni < 0 || nv < 0 || m < 0 || t->kind != 0) return fail(ST_ERR_ARG, "st_per_step_device: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    tree_lock lk(t);
    device_guard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (ni > 0 || nv > 0) CU(st::launch_update(t, idx_dev, ni, vals_dev, nv, nullptr, nullptr, s));
    CU(st::launch_sample(t, u_dev, m, out_idx_dev, out_prio_dev, nullptr, s, out_prob_dev));
    return ST_OK;
}

int st_host_alloc(int64_t bytes, void **out)
{
    if (!out || bytes <= 0) return fail(ST_ERR_ARG, "st_host_alloc: bad argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return ST_OK;
}

int st_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return ST_OK;
}

namespace {
int prob_size(const st_tree *t) { return t->dtype == ST_F32 ? 4 : 8; }

int host_pipeline_init(st_tree *t)
{
    if (t->hs_main) return ST_OK;
    CU(cudaStreamCreateWithFlags(&t->hs_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&t->hs_main, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&t->hs_out, cudaStreamNonBlocking));
    for (auto &h : t->hs) {
        CU(cudaEventCreateWithFlags(&h.ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h.ev_free, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h.ev_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h.ev_read, cudaEventDisableTiming));
        CU(cudaMallocHost((void **)&h.h_flag, 32 * sizeof(unsigned int)));
    }
    return ST_OK;
}
}  // namespace

int st_per_step_host_submit(st_tree *t, int slot, const int64_t *idx_host, int64_t ni, const void *vals_host,
                            int64_t nv, const double *u_host, int64_t m, int64_t *out_idx_host,
                            void *out_prio_host, void *out_prob_host)
{
    if (!t || slot < 0 || slot >= ST_HOST_SLOTS || ni < 0 || nv < 0 || m < 0 || t->kind != 0 || (m > 0 && (!u_host || !out_idx_host)))
        return fail(ST_ERR_ARG, "st_per_step_host_submit: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    tree_lock lk(t);
    device_guard g(t->device);
    { const int st = host_pipeline_init(t); if (st != ST_OK) return st; }
    st_tree::host_slot &h = t->hs[slot];
    if (h.busy) {   // the slot's previous step has not been collected: finish it first (its status is dropped)
        CU(cudaEventSynchronize(h.ev_read));
        h.busy = 0;
    }
    auto rnd = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t es = (size_t)t->esize, ps = (size_t)prob_size(t);
    const size_t o_idx = 0, o_val = o_idx + rnd(ni * 8), o_u = o_val + rnd(nv * es), o_out = o_u + rnd(m * 8),
                 o_prio = o_out + rnd(m * 8), o_prob = o_prio + rnd(m * es), total = o_prob + rnd(m * ps);
    if (total > h.bytes) {
        CU(cudaDeviceSynchronize());
        if (h.dev) cudaFree(h.dev);
        h.dev = nullptr; h.bytes = 0;
        CU(cudaMalloc(&h.dev, total));
        h.bytes = total;
    }
    // make sure the update workspace never reallocates while the other slot's kernels are in flight
    if (ni > t->reserved_batch || !t->ws.base) {
        const int st = st_reserve(t, ni > 0 ? ni : 1);
        if (st != ST_OK) return st;
    }
    char *d = (char *)h.dev;
    CU(cudaStreamWaitEvent(t->hs_in, h.ev_free, 0));    // the slot's previous inputs have been consumed
    if (ni) CU(cudaMemcpyAsync(d + o_idx, idx_host, ni * 8, cudaMemcpyHostToDevice, t->hs_in));
    if (nv) CU(cudaMemcpyAsync(d + o_val, vals_host, nv * es, cudaMemcpyHostToDevice, t->hs_in));
    if (m) CU(cudaMemcpyAsync(d + o_u, u_host, m * 8, cudaMemcpyHostToDevice, t->hs_in));
    CU(cudaEventRecord(h.ev_in, t->hs_in));
    CU(cudaStreamWaitEvent(t->hs_main, h.ev_in, 0));
    CU(cudaStreamWaitEvent(t->hs_main, h.ev_read, 0));  // the slot's previous results have left the device
    if (ni > 0 || nv > 0)
        CU(st::launch_update(t, (const int64_t *)(d + o_idx), ni, d + o_val, nv, nullptr, nullptr, t->hs_main));
    if (m)
        CU(st::launch_sample(t, (const double *)(d + o_u), m, (int64_t *)(d + o_out), out_prio_host ? d + o_prio : nullptr,
                             nullptr, t->hs_main, out_prob_host ? d + o_prob : nullptr));
    CU(cudaEventRecord(h.ev_free, t->hs_main));
    CU(cudaEventRecord(h.ev_done, t->hs_main));
    CU(cudaStreamWaitEvent(t->hs_out, h.ev_done, 0));
    if (m) {
        CU(cudaMemcpyAsync(out_idx_host, d + o_out, m * 8, cudaMemcpyDeviceToHost, t->hs_out));
        if (out_prio_host) CU(cudaMemcpyAsync(out_prio_host, d + o_prio, m * es, cudaMemcpyDeviceToHost, t->hs_out));
        if (out_prob_host) CU(cudaMemcpyAsync(out_prob_host, d + o_prob, m * ps, cudaMemcpyDeviceToHost, t->hs_out));
    }
    CU(cudaMemcpyAsync(h.h_flag, t->d_status, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, t->hs_out));
    CU(cudaEventRecord(h.ev_read, t->hs_out));
    h.busy = 1;
    return ST_OK;
}

int st_per_step_host_wait(st_tree *t, int slot)
{
    if (!t || slot < 0 || slot >= ST_HOST_SLOTS) return fail(ST_ERR_ARG, "st_per_step_host_wait: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    st_tree::host_slot &h = t->hs[slot];
    if (!h.busy) return ST_OK;
    CU(cudaEventSynchronize(h.ev_read));
    h.busy = 0;
    const unsigned int f = h.h_flag[0] | h.h_flag[ST_SW_STICKY];
    if (f) {
        CU(cudaStreamSynchronize(t->hs_main));
        CU(clear_status_words(t, t->hs_main));
        CU(cudaStreamSynchronize(t->hs_main));
        const int st = status_from_flags(f);
        return fail(st, "%s", st_status_string(st));
    }
    return ST_OK;
}

// ---- in-place leaf arithmetic ---------------------------------------------------------------------
int st_inplace_device(st_tree *t, int op, const void *scalar_host, const void *operand_dev, int64_t operand_len,
                      void *stream)
{
    if (!t || t->kind != 0 || op < ST_OP_ADD || op > ST_OP_FILL || (!scalar_host && !operand_dev))
        return fail(ST_ERR_ARG, "st_inplace_device: bad argument");
    if (operand_dev && operand_len != t->n) return fail(ST_ERR_SIZE, "operand must have the size of the array");
    const bool is_f = (t->dtype == ST_F32 || t->dtype == ST_F64);
    if (op == ST_OP_DIV && !is_f) return fail(ST_ERR_DTYPE, "in-place true division needs a floating-point tree");
    if (op == ST_OP_FILL && operand_dev) return fail(ST_ERR_ARG, "st_inplace_device: fill takes a scalar");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_inplace(t, op, scalar_host, operand_dev, (cudaStream_t)stream));
    return ST_OK;
}

// ---- min tree -------------------------------------------------------------------------------------
int st_min_create(int64_t n, int dtype, int device, void *external_heap, st_tree **out)
{
    const int st = st_create(n, dtype, device, external_heap, out);
    if (st != ST_OK) return st;
    st_tree *t = *out;
    device_guard g(device);
    t->kind = 1;
    if (t->sidx) {   // the search index belongs to sum trees
        cudaFree(t->sidx);
        t->sidx = nullptr;
        t->sidx_bytes = 0;
        t->sidx_nb = 0;
    }
    cudaError_t e = cudaMalloc(&t->aux, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = st::launch_min_init(t, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        st_destroy(t);
        *out = nullptr;
        return cuda_fail(e, "st_min_create");
    }
    return ST_OK;
}

int st_min_update_device(st_tree *t, const int64_t *idx_dev, int64_t ni, const void *vals_dev, int64_t nv, void *stream)
{
    if (!t || t->kind != 1 || ni < 0 || nv < 0) return fail(ST_ERR_ARG, "st_min_update_device: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    if (ni > 0x7fffffffLL) return fail(ST_ERR_TOO_LARGE, "st_min_update_device: batch too large");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_min_update(t, idx_dev, ni, vals_dev, nv, (cudaStream_t)stream));
    return ST_OK;
}

int st_min_update_host(st_tree *t, const int64_t *idx_host, int64_t ni, const void *vals_host, int64_t nv)
{
    if (!t || t->kind != 1 || ni < 0 || nv < 0) return fail(ST_ERR_ARG, "st_min_update_host: bad argument");
    if (ni > 0 && nv == 0) return fail(ST_ERR_EMPTY_VALS, "%s", st_status_string(ST_ERR_EMPTY_VALS));
    if (ni == 0) return ST_OK;
    tree_lock lk(t);
    device_guard g(t->device);
    scratch_carver sc(t);
    sc.add(ni * sizeof(int64_t)); sc.add(nv * t->esize);
    CU(sc.commit());
    int64_t *di = (int64_t *)sc.take(ni * sizeof(int64_t));
    void *dv = sc.take(nv * t->esize);
    CU(cudaMemcpyAsync(di, idx_host, ni * sizeof(int64_t), cudaMemcpyHostToDevice, 0));
    CU(cudaMemcpyAsync(dv, vals_host, nv * t->esize, cudaMemcpyHostToDevice, 0));
    CU(st::launch_min_update(t, di, ni, dv, nv, 0));
    const int st = fetch_status(t, 0, true, true);
    if (st != ST_OK) return fail(st, "%s", st_status_string(st));
    return ST_OK;
}

int st_min_build(st_tree *t, void *stream)
{
    if (!t || t->kind != 1) return fail(ST_ERR_ARG, "st_min_build: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    CU(st::launch_min_build(t, (cudaStream_t)stream));
    return ST_OK;
}

// ---- bookkeeping ------------------------------------------------------------------------------------
int st_update_counters(st_tree *t, void *stream, int64_t out[2], int reset)
{
    if (!t || !out) return fail(ST_ERR_ARG, "st_update_counters: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(t->h_status, t->d_status, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    out[0] = t->stat_updates;
    out[1] = (int64_t)t->h_status[ST_SW_ZERO_DELTA];
    if (reset) {
        t->stat_updates = 0;
        CU(cudaMemsetAsync(t->d_status + ST_SW_ZERO_DELTA, 0, sizeof(unsigned int), s));
        CU(cudaStreamSynchronize(s));
    }
    return ST_OK;
}

int st_reserve_sharded(st_tree *t, int64_t max_batch, int shards)
{
    if (!t || max_batch < 1 || shards < 1 || shards > 64) return fail(ST_ERR_ARG, "st_reserve_sharded: bad argument");
    tree_lock lk(t);
    device_guard g(t->device);
    { const int st = st_reserve(t, max_batch); if (st != ST_OK) return st; }
    size_t need = st::shard_plan_scratch_bytes(max_batch, shards);
    const size_t nchunks = (size_t)((max_batch + 1023) / 1024);
    const size_t recs = nchunks * (size_t)(shards > 1 ? shards - 1 : 1) * 64;   // top-fold records (<= 64 B each)
    if (recs > need) need = recs;
    scratch_carver c(t);
    c.add(need);
    CU(c.commit());
    return ST_OK;
}

}  // extern "C"
