// vrod_capi.cu -- the C ABI (include/vrod_knn.h): contexts, collections, SEARCH.
//
// Stands behind the reference's command bodies (all empty today): CreateCollectionCommand
// (src/command/types.rs:14-19), InsertCommand / BulkInsertCommand (:62-67, :75-80), SearchCommand
// (:114-119), and plays the part of the collections the reference's Database never got
// (src/database/mod.rs:6-10).  No CPU fallback: every entry point needs the CUDA device.
#include "../../include/vrod_knn.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "knn_batched.cuh"
#include "knn_scan.cuh"

using namespace vrod;

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static vrod_status fail(vrod_status st, const std::string &msg) {
    g_last_error = msg;
    return st;
}

// No C++ exception crosses the C ABI: every entry point that can allocate runs inside this guard.
template <typename F>
static vrod_status guarded(F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        try { g_last_error = "host allocation failed"; } catch (...) {}
        return VROD_ENOMEM;
    } catch (const std::exception &e) {
        try { g_last_error = std::string("internal error: ") + e.what(); } catch (...) {}
        return VROD_EINVAL;
    } catch (...) {
        return VROD_EINVAL;
    }
}
#define VROD_CUDA(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(e__ == cudaErrorMemoryAllocation ? VROD_ENOMEM : VROD_ECUDA,                      \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                             \
    } while (0)

// -------------------------------------------------------------------------------------------------
// NCCL, loaded on first use (a single-GPU context never touches it)
// -------------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static vrod_status nccl_load() {
    if (g_nccl.lib) return VROD_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(VROD_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
        return fail(VROD_ENCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return VROD_OK;
}
#define VROD_NCCL(expr)                                                                                   \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess) return fail(VROD_ENCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

// -------------------------------------------------------------------------------------------------
// handles
// -------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

struct vrod_ctx {
    int device = 0, rank = 0, world = 1, sms = 0;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    std::map<std::string, vrod_collection *> colls;
    // scratch shared by every collection of the context
    DevBuf blk_cand, small, q_pad, q_dev, hits_local, hits_all, out_ids, out_dist, batched;
    PinBuf q_host, ids_host, dist_host, status_host;
    unsigned long long *dev_counters = nullptr;  // [0] exact rescans (counted on the device)
    vrod_stats stats{};
    // fused NVLink exchange (sharded contexts): every rank's window mapped through CUDA IPC
    unsigned char *xchg = nullptr;              // this rank's window
    std::vector<unsigned char *> xchg_peers;    // [world] mapped windows (own pointer at [rank])
    unsigned char **d_windows = nullptr;        // device copy of the table
    int *d_xchg_err = nullptr;
    uint32_t xchg_seq = 0;
    bool fused_exchange = false;
    // optional kernel timing (vrod_ctx_profile)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
    size_t ev_used = 0;
    cudaEvent_t prof_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
};

struct vrod_collection {
    vrod_ctx *ctx = nullptr;
    std::string name;
    uint32_t dim = 0, ld = 0;
    vrod_metric metric = VROD_EUCLIDEAN;
    uint64_t capacity = 0;    // global
    uint64_t shard_rows = 0;  // rows per rank (capacity split)
    uint64_t id_base = 0;     // first global id of this rank's range
    uint64_t count = 0;       // global rows inserted
    uint64_t local = 0;       // rows held by this rank
    float *rows = nullptr, *inv_norm = nullptr, *sq_norm = nullptr;
    int *flags = nullptr;     // device: bit0 non-finite value seen, bit1 value outside the f32 scan's safe range
    bool fast_ok = true;
    int path = 0;
    // bf16 operand mirror for the batched path (knn_batched.cu), built lazily by the first batched search and
    // extended when rows were appended since; rows [0, mirror_rows) are valid
    unsigned short *rows_h = nullptr;
    uint64_t mirror_rows = 0;
    bool mirror_failed = false;   // the mirror did not fit: the batched path feeds the f32 rows as tf32 instead
    // share of the queries of recent batched searches whose guard failed (they were rescanned one by one): the
    // automatic path choice discounts the batched pass by it, so that data the tensor-core pass cannot resolve
    // (tight clusters under the Euclidean metric) stops paying for a pass that proves nothing
    double rescan_share = 0.0;
};

static int *ctx_ticket(vrod_ctx *c) { return reinterpret_cast<int *>(c->small.p); }
static int *ctx_status(vrod_ctx *c) { return reinterpret_cast<int *>(c->small.p) + 64; }
constexpr uint32_t kMaxBatch = 1u << 16;

static vrod_status setup_fused_exchange(vrod_ctx *c);

static vrod_status ctx_init(vrod_ctx *c, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(VROD_ENOGPU, "no CUDA device: vrod_knn has no CPU path");
    if (device < 0 || device >= ndev) return fail(VROD_EINVAL, "device ordinal out of range");
    VROD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VROD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(VROD_ENOGPU, "vrod_knn is built for sm_100a (B200) only");
    c->device = device;
    c->sms = prop.multiProcessorCount;
    VROD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    VROD_CUDA(c->blk_cand.ensure(scan_cand_bytes(c->sms)));
    VROD_CUDA(c->small.ensure(256 + sizeof(int) * kMaxBatch + 64));
    VROD_CUDA(cudaMemsetAsync(c->small.p, 0, c->small.bytes, c->stream));
    VROD_CUDA(cudaMalloc(&c->dev_counters, 64));
    VROD_CUDA(cudaMemsetAsync(c->dev_counters, 0, 64, c->stream));
    VROD_CUDA(cudaStreamSynchronize(c->stream));
    return VROD_OK;
}

// -------------------------------------------------------------------------------------------------
// context API
// -------------------------------------------------------------------------------------------------
static vrod_status vrod_ctx_create_impl(int device, vrod_ctx **out);
extern "C" vrod_status vrod_ctx_create(int device, vrod_ctx **out) {
    return guarded([&]() -> vrod_status { return vrod_ctx_create_impl(device, out); });
}
static vrod_status vrod_ctx_create_impl(int device, vrod_ctx **out) {
    if (!out) return fail(VROD_EINVAL, "out is NULL");
    *out = nullptr;
    std::unique_ptr<vrod_ctx> c(new (std::nothrow) vrod_ctx());
    if (!c) return fail(VROD_ENOMEM, "host allocation failed");
    vrod_status st = ctx_init(c.get(), device);
    if (st != VROD_OK) return st;
    *out = c.release();
    return VROD_OK;
}

static vrod_status vrod_comm_unique_id_impl(void *out);
extern "C" vrod_status vrod_comm_unique_id(void *out) {
    return guarded([&]() -> vrod_status { return vrod_comm_unique_id_impl(out); });
}
static vrod_status vrod_comm_unique_id_impl(void *out) {
    if (!out) return fail(VROD_EINVAL, "out is NULL");
    vrod_status st = nccl_load();
    if (st != VROD_OK) return st;
    static_assert(sizeof(ncclUniqueId) == VROD_COMM_ID_BYTES, "communicator id size");
    ncclUniqueId id;
    VROD_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return VROD_OK;
}

static vrod_status vrod_ctx_create_sharded_impl(int device, int rank, int world, const void *comm_id, vrod_ctx **out);
extern "C" vrod_status vrod_ctx_create_sharded(int device, int rank, int world, const void *comm_id, vrod_ctx **out) {
    return guarded([&]() -> vrod_status { return vrod_ctx_create_sharded_impl(device, rank, world, comm_id, out); });
}
static vrod_status vrod_ctx_create_sharded_impl(int device, int rank, int world, const void *comm_id, vrod_ctx **out) {
    if (!out) return fail(VROD_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > 256 || rank < 0 || rank >= world) return fail(VROD_EINVAL, "bad rank/world");
    if (world > 1 && !comm_id) return fail(VROD_EINVAL, "comm_id is NULL");
    std::unique_ptr<vrod_ctx> c(new (std::nothrow) vrod_ctx());
    if (!c) return fail(VROD_ENOMEM, "host allocation failed");
    vrod_status st = ctx_init(c.get(), device);
    if (st != VROD_OK) return st;
    c->rank = rank;
    c->world = world;
    if (world > 1) {
        st = nccl_load();
        if (st != VROD_OK) return st;
        ncclUniqueId id;
        memcpy(&id, comm_id, sizeof(id));
        VROD_NCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
        st = setup_fused_exchange(c.get());
        if (st != VROD_OK) return st;
    }
    *out = c.release();
    return VROD_OK;
}

// Map every rank's exchange window into this process (CUDA IPC over NVLink P2P).  Collective.  On any failure
// on any rank all ranks fall back to ncclAllGather + merge.
static vrod_status setup_fused_exchange(vrod_ctx *c) {
    if (c->world > (int)kXchgMaxWorld || getenv("VROD_NO_P2P_EXCHANGE")) return VROD_OK;
    const int W = c->world;
    bool ok = true;
    cudaIpcMemHandle_t mine{};
    unsigned char *d_tmp = nullptr;   // [W+1] handles, then [W+1] ok bytes
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    ok = ok && cudaMalloc(&c->xchg, xchg_window_bytes()) == cudaSuccess;
    ok = ok && cudaMemsetAsync(c->xchg, 0, xchg_window_bytes(), c->stream) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_xchg_err, 64) == cudaSuccess && cudaMemsetAsync(c->d_xchg_err, 0, 64, c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine, c->xchg) == cudaSuccess;
    VROD_CUDA(cudaMalloc(&d_tmp, (size_t)(W + 1) * (hb + 64)));
    std::vector<cudaIpcMemHandle_t> all(W);
    VROD_CUDA(cudaMemcpyAsync(d_tmp + (size_t)W * hb, &mine, hb, cudaMemcpyHostToDevice, c->stream));
    VROD_NCCL(g_nccl.AllGather(d_tmp + (size_t)W * hb, d_tmp, hb, ncclChar, c->comm, c->stream));
    VROD_CUDA(cudaMemcpyAsync(all.data(), d_tmp, (size_t)W * hb, cudaMemcpyDeviceToHost, c->stream));
    VROD_CUDA(cudaStreamSynchronize(c->stream));
    c->xchg_peers.assign(W, nullptr);
    for (int r = 0; r < W && ok; ++r) {
        if (r == c->rank) {
            c->xchg_peers[r] = c->xchg;
        } else {
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
            }
            c->xchg_peers[r] = reinterpret_cast<unsigned char *>(p);
        }
    }
    if (ok) {
        ok = cudaMalloc(&c->d_windows, sizeof(unsigned char *) * W) == cudaSuccess &&
             cudaMemcpyAsync(c->d_windows, c->xchg_peers.data(), sizeof(unsigned char *) * W, cudaMemcpyHostToDevice, c->stream) ==
                 cudaSuccess;
    }
    // agree: fused only if every rank mapped every window
    unsigned char *okbytes = d_tmp + (size_t)(W + 1) * hb;
    const unsigned char my_ok = ok ? 1 : 0;
    std::vector<unsigned char> oks(W);
    VROD_CUDA(cudaMemcpyAsync(okbytes + W, &my_ok, 1, cudaMemcpyHostToDevice, c->stream));
    VROD_NCCL(g_nccl.AllGather(okbytes + W, okbytes, 1, ncclChar, c->comm, c->stream));
    VROD_CUDA(cudaMemcpyAsync(oks.data(), okbytes, W, cudaMemcpyDeviceToHost, c->stream));
    VROD_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_tmp);
    bool all_ok = true;
    for (unsigned char v : oks) all_ok = all_ok && v;
    c->fused_exchange = all_ok;
    if (getenv("VROD_VERBOSE"))
        fprintf(stderr, "[vrod] rank %d/%d: fused NVLink exchange %s\n", c->rank, c->world, all_ok ? "enabled" : "unavailable (NCCL all-gather)");
    return VROD_OK;
}

static void collection_free(vrod_collection *c) {
    if (!c) return;
    cudaFree(c->rows);
    cudaFree(c->inv_norm);
    cudaFree(c->sq_norm);
    cudaFree(c->flags);
    cudaFree(c->rows_h);
    delete c;
}

extern "C" void vrod_ctx_destroy(vrod_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->colls) collection_free(kv.second);
    for (size_t r = 0; r < ctx->xchg_peers.size(); ++r)
        if ((int)r != ctx->rank && ctx->xchg_peers[r]) cudaIpcCloseMemHandle(ctx->xchg_peers[r]);
    cudaFree(ctx->xchg);
    cudaFree(ctx->d_windows);
    cudaFree(ctx->d_xchg_err);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    DevBuf *dev[] = {&ctx->blk_cand, &ctx->small, &ctx->q_pad, &ctx->q_dev, &ctx->hits_local,
                     &ctx->hits_all, &ctx->out_ids, &ctx->out_dist, &ctx->batched};
    for (DevBuf *b : dev) b->release();
    PinBuf *pin[] = {&ctx->q_host, &ctx->ids_host, &ctx->dist_host, &ctx->status_host};
    for (PinBuf *b : pin) b->release();
    cudaFree(ctx->dev_counters);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static vrod_status vrod_ctx_synchronize_impl(vrod_ctx *ctx);
extern "C" vrod_status vrod_ctx_synchronize(vrod_ctx *ctx) {
    return guarded([&]() -> vrod_status { return vrod_ctx_synchronize_impl(ctx); });
}
static vrod_status vrod_ctx_synchronize_impl(vrod_ctx *ctx) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    return VROD_OK;
}
extern "C" void *vrod_ctx_stream(vrod_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int vrod_ctx_rank(vrod_ctx *ctx) { return ctx ? ctx->rank : -1; }
extern "C" int vrod_ctx_world(vrod_ctx *ctx) { return ctx ? ctx->world : -1; }

static vrod_status vrod_ctx_profile_impl(vrod_ctx *ctx, int enable);
extern "C" vrod_status vrod_ctx_profile(vrod_ctx *ctx, int enable) {
    return guarded([&]() -> vrod_status { return vrod_ctx_profile_impl(ctx, enable); });
}
static vrod_status vrod_ctx_profile_impl(vrod_ctx *ctx, int enable) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    ctx->profiling = enable != 0;
    return VROD_OK;
}

static vrod_status vrod_ctx_profile_read_impl(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches);
extern "C" vrod_status vrod_ctx_profile_read(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches) {
    return guarded([&]() -> vrod_status { return vrod_ctx_profile_read_impl(ctx, kernel_ms, launches); });
}
static vrod_status vrod_ctx_profile_read_impl(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    VROD_CUDA(cudaSetDevice(ctx->device));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0.f;
        VROD_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
        total += ms;
    }
    if (kernel_ms) *kernel_ms = total;
    if (launches) *launches = ctx->ev_used / 2;
    ctx->ev_used = 0;
    return VROD_OK;
}

static vrod_status vrod_ctx_stats_impl(vrod_ctx *ctx, vrod_stats *out);
extern "C" vrod_status vrod_ctx_stats(vrod_ctx *ctx, vrod_stats *out) {
    return guarded([&]() -> vrod_status { return vrod_ctx_stats_impl(ctx, out); });
}
static vrod_status vrod_ctx_stats_impl(vrod_ctx *ctx, vrod_stats *out) {
    if (!ctx || !out) return fail(VROD_EINVAL, "NULL argument");
    VROD_CUDA(cudaSetDevice(ctx->device));
    unsigned long long dc[2] = {0, 0};
    VROD_CUDA(cudaMemcpyAsync(dc, ctx->dev_counters, sizeof(dc), cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.exact_rescans = dc[0];
    *out = ctx->stats;
    return VROD_OK;
}

// -------------------------------------------------------------------------------------------------
// collections
// -------------------------------------------------------------------------------------------------
static vrod_status vrod_collection_create_impl(vrod_ctx *ctx, const char *name, uint32_t dim, vrod_metric metric,
                                              uint64_t capacity_rows, vrod_collection **out);
extern "C" vrod_status vrod_collection_create(vrod_ctx *ctx, const char *name, uint32_t dim, vrod_metric metric,
                                              uint64_t capacity_rows, vrod_collection **out) {
    return guarded([&]() -> vrod_status { return vrod_collection_create_impl(ctx, name, dim, metric, capacity_rows, out); });
}
static vrod_status vrod_collection_create_impl(vrod_ctx *ctx, const char *name, uint32_t dim, vrod_metric metric,
                                              uint64_t capacity_rows, vrod_collection **out) {
    if (out) *out = nullptr;
    if (!ctx || !name || !*name) return fail(VROD_EINVAL, "ctx/name is NULL or empty");
    if (dim == 0 || dim > (1u << 20)) return fail(VROD_EINVAL, "dim must be in [1, 2^20]");
    if (metric != VROD_EUCLIDEAN && metric != VROD_COSINE) return fail(VROD_EINVAL, "unknown metric");
    if (capacity_rows == 0) return fail(VROD_EINVAL, "capacity_rows must be > 0");
    if (ctx->colls.count(name)) return fail(VROD_EEXISTS, std::string("collection '") + name + "' already exists");
    VROD_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<vrod_collection> c(new (std::nothrow) vrod_collection());
    if (!c) return fail(VROD_ENOMEM, "host allocation failed");
    c->ctx = ctx;
    c->name = name;
    c->dim = dim;
    c->ld = (dim + 3u) & ~3u;
    c->metric = metric;
    c->capacity = capacity_rows;
    c->shard_rows = (capacity_rows + ctx->world - 1) / ctx->world;
    if (c->shard_rows >= 0xFFFFFFFFull) return fail(VROD_EINVAL, "more than 2^32-1 rows per GPU");
    c->id_base = c->shard_rows * (uint64_t)ctx->rank;
    const size_t row_bytes = (size_t)c->shard_rows * c->ld * sizeof(float);
    cudaError_t e = cudaMalloc(&c->rows, row_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->inv_norm, (size_t)c->shard_rows * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->sq_norm, (size_t)c->shard_rows * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->flags, 64);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->flags, 0, 64, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->rows); cudaFree(c->inv_norm); cudaFree(c->sq_norm); cudaFree(c->flags);
        return fail(e == cudaErrorMemoryAllocation ? VROD_ENOMEM : VROD_ECUDA,
                    std::string("allocating the collection: ") + cudaGetErrorString(e));
    }
    vrod_collection *raw = c.release();
    ctx->colls[name] = raw;
    if (out) *out = raw;
    return VROD_OK;
}

static vrod_status vrod_collection_get_impl(vrod_ctx *ctx, const char *name, vrod_collection **out);
extern "C" vrod_status vrod_collection_get(vrod_ctx *ctx, const char *name, vrod_collection **out) {
    return guarded([&]() -> vrod_status { return vrod_collection_get_impl(ctx, name, out); });
}
static vrod_status vrod_collection_get_impl(vrod_ctx *ctx, const char *name, vrod_collection **out) {
    if (!ctx || !name || !out) return fail(VROD_EINVAL, "NULL argument");
    auto it = ctx->colls.find(name);
    if (it == ctx->colls.end()) {
        *out = nullptr;
        return fail(VROD_ENOTFOUND, std::string("no collection '") + name + "'");
    }
    *out = it->second;
    return VROD_OK;
}

static vrod_status vrod_collection_drop_impl(vrod_ctx *ctx, const char *name);
extern "C" vrod_status vrod_collection_drop(vrod_ctx *ctx, const char *name) {
    return guarded([&]() -> vrod_status { return vrod_collection_drop_impl(ctx, name); });
}
static vrod_status vrod_collection_drop_impl(vrod_ctx *ctx, const char *name) {
    if (!ctx || !name) return fail(VROD_EINVAL, "NULL argument");
    auto it = ctx->colls.find(name);
    if (it == ctx->colls.end()) return fail(VROD_ENOTFOUND, std::string("no collection '") + name + "'");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    collection_free(it->second);
    ctx->colls.erase(it);
    return VROD_OK;
}

static vrod_status vrod_collection_list_impl(vrod_ctx *ctx, char *buf, size_t cap, size_t *needed);
extern "C" vrod_status vrod_collection_list(vrod_ctx *ctx, char *buf, size_t cap, size_t *needed) {
    return guarded([&]() -> vrod_status { return vrod_collection_list_impl(ctx, buf, cap, needed); });
}
static vrod_status vrod_collection_list_impl(vrod_ctx *ctx, char *buf, size_t cap, size_t *needed) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    std::string s;
    for (auto &kv : ctx->colls) {
        if (!s.empty()) s += '\n';
        s += kv.first;
    }
    if (needed) *needed = s.size() + 1;
    if (buf && cap > 0) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return VROD_OK;
}

static vrod_status vrod_collection_info_impl(vrod_collection *c, uint32_t *dim, vrod_metric *metric, uint64_t *count,
                                            uint64_t *capacity);
extern "C" vrod_status vrod_collection_info(vrod_collection *c, uint32_t *dim, vrod_metric *metric, uint64_t *count,
                                            uint64_t *capacity) {
    return guarded([&]() -> vrod_status { return vrod_collection_info_impl(c, dim, metric, count, capacity); });
}
static vrod_status vrod_collection_info_impl(vrod_collection *c, uint32_t *dim, vrod_metric *metric, uint64_t *count,
                                            uint64_t *capacity) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (dim) *dim = c->dim;
    if (metric) *metric = c->metric;
    if (count) *count = c->count;
    if (capacity) *capacity = c->capacity;
    return VROD_OK;
}

static vrod_status vrod_collection_shard_impl(vrod_collection *c, uint64_t *id_base, uint64_t *local_rows);
extern "C" vrod_status vrod_collection_shard(vrod_collection *c, uint64_t *id_base, uint64_t *local_rows) {
    return guarded([&]() -> vrod_status { return vrod_collection_shard_impl(c, id_base, local_rows); });
}
static vrod_status vrod_collection_shard_impl(vrod_collection *c, uint64_t *id_base, uint64_t *local_rows) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (id_base) *id_base = c->id_base;
    if (local_rows) *local_rows = c->local;
    return VROD_OK;
}

static vrod_status vrod_collection_set_path_impl(vrod_collection *c, int path);
extern "C" vrod_status vrod_collection_set_path(vrod_collection *c, int path) {
    return guarded([&]() -> vrod_status { return vrod_collection_set_path_impl(c, path); });
}
static vrod_status vrod_collection_set_path_impl(vrod_collection *c, int path) {
    if (!c || path < 0 || path > 4) return fail(VROD_EINVAL, "bad path");
    c->path = path;
    return VROD_OK;
}

// rows [g0, g0+n) of the global sequence: which part lands on this rank?  Returns local start / count / source offset.
static void shard_overlap(const vrod_collection *c, uint64_t g0, uint64_t n, uint64_t *loc0, uint64_t *cnt, uint64_t *src_off) {
    const uint64_t lo = c->id_base, hi = c->id_base + c->shard_rows;
    const uint64_t a = g0 > lo ? g0 : lo;
    const uint64_t b = (g0 + n) < hi ? (g0 + n) : hi;
    if (b <= a) {
        *loc0 = 0; *cnt = 0; *src_off = 0;
        return;
    }
    *loc0 = a - lo;
    *cnt = b - a;
    *src_off = a - g0;
}

// after new rows landed in [loc0, loc0+cnt): norms + validation
static vrod_status finish_append(vrod_collection *c, uint64_t loc0, uint64_t cnt, bool check_flags) {
    vrod_ctx *ctx = c->ctx;
    if (c->mirror_rows > loc0) c->mirror_rows = loc0;
    VROD_CUDA(launch_row_norms(c->rows, (uint32_t)loc0, (uint32_t)cnt, c->ld, c->inv_norm, c->sq_norm, c->flags, ctx->stream));
    if (cnt) ctx->stats.kernel_launches++;
    if (check_flags) {
        int fl = 0;
        VROD_CUDA(cudaMemcpyAsync(&fl, c->flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        VROD_CUDA(cudaStreamSynchronize(ctx->stream));
        if (fl & 1) {
            VROD_CUDA(cudaMemsetAsync(c->flags, 0, sizeof(int), ctx->stream));
            return fail(VROD_EINVAL, "rows contain NaN or infinity");
        }
        if (fl & 2) c->fast_ok = false;
    }
    return VROD_OK;
}

// Sharded contexts: every rank sees the whole batch but validates (on the device) only the rows of its own
// range, so a rank that owns none of a batch's bad rows would accept what the owning rank rejects and the ranks'
// counts would drift apart.  All ranks therefore check the whole host buffer first (exponent all ones = inf/NaN).
static bool host_rows_finite(const float *rows, size_t n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rows);
    uint32_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad |= ((w[i] & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
    return bad == 0;
}

// single-GPU collections: make room for at least `need` rows (capacity at least doubles)
static vrod_status collection_grow(vrod_collection *c, uint64_t need) {
    vrod_ctx *ctx = c->ctx;
    uint64_t cap = c->capacity * 2 > need ? c->capacity * 2 : need;
    if (cap >= 0xFFFFFFFFull) cap = 0xFFFFFFFEull;
    if (cap < need) return fail(VROD_ENOMEM, "more than 2^32-2 rows on one GPU");
    float *rows = nullptr, *inv = nullptr, *sq = nullptr;
    cudaError_t e = cudaSuccess;
    for (int attempt = 0; attempt < 2; ++attempt) {   // doubled capacity first, exactly `need` rows if that does not fit
        e = cudaMalloc(&rows, (size_t)cap * c->ld * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&inv, (size_t)cap * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&sq, (size_t)cap * sizeof(float));
        if (e != cudaErrorMemoryAllocation || cap == need) break;
        cudaGetLastError();
        cudaFree(rows); cudaFree(inv); cudaFree(sq);
        rows = inv = sq = nullptr;
        cap = need;
    }
    if (e == cudaSuccess && c->local) {
        e = cudaMemcpyAsync(rows, c->rows, (size_t)c->local * c->ld * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(inv, c->inv_norm, (size_t)c->local * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sq, c->sq_norm, (size_t)c->local * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(rows); cudaFree(inv); cudaFree(sq);
        return fail(e == cudaErrorMemoryAllocation ? VROD_ENOMEM : VROD_ECUDA, std::string("growing the collection: ") + cudaGetErrorString(e));
    }
    cudaFree(c->rows); cudaFree(c->inv_norm); cudaFree(c->sq_norm);
    cudaFree(c->rows_h);   // sized by the old capacity: rebuilt by the next batched search
    c->rows_h = nullptr;
    c->mirror_rows = 0;
    c->mirror_failed = false;
    c->rows = rows; c->inv_norm = inv; c->sq_norm = sq;
    c->capacity = cap;
    c->shard_rows = cap;
    return VROD_OK;
}

static vrod_status vrod_collection_insert_impl(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id);
extern "C" vrod_status vrod_collection_insert(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id) {
    return guarded([&]() -> vrod_status { return vrod_collection_insert_impl(c, rows, n, first_id); });
}
static vrod_status vrod_collection_insert_impl(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (n == 0) {
        if (first_id) *first_id = c->count;
        return VROD_OK;
    }
    if (!rows) return fail(VROD_EINVAL, "rows is NULL");
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaSetDevice(ctx->device));
    if (c->count + n > c->capacity) {
        if (ctx->world > 1) return fail(VROD_ENOMEM, "collection capacity exceeded (sharded collections do not grow)");
        vrod_status gs = collection_grow(c, c->count + n);
        if (gs != VROD_OK) return gs;
    }
    if (ctx->world > 1 && !host_rows_finite(rows, (size_t)n * c->dim))
        return fail(VROD_EINVAL, "rows contain NaN or infinity");   // decided identically on every rank (same rows)
    uint64_t loc0, cnt, off;
    shard_overlap(c, c->count, n, &loc0, &cnt, &off);
    if (cnt) {
        const float *src = rows + (size_t)off * c->dim;
        float *dst = c->rows + (size_t)loc0 * c->ld;
        if (c->ld != c->dim) VROD_CUDA(cudaMemsetAsync(dst, 0, (size_t)cnt * c->ld * sizeof(float), ctx->stream));
        VROD_CUDA(cudaMemcpy2DAsync(dst, (size_t)c->ld * sizeof(float), src, (size_t)c->dim * sizeof(float),
                                    (size_t)c->dim * sizeof(float), (size_t)cnt, cudaMemcpyHostToDevice, ctx->stream));
    }
    vrod_status st = finish_append(c, loc0, cnt, true);
    if (st != VROD_OK) return st;  // rejected rows are not counted: the next insert overwrites them
    if (first_id) *first_id = c->count;
    c->count += n;
    c->local += cnt;
    return VROD_OK;
}

static vrod_status vrod_collection_fill_synthetic_impl(vrod_collection *c, uint64_t n, uint64_t seed);
extern "C" vrod_status vrod_collection_fill_synthetic(vrod_collection *c, uint64_t n, uint64_t seed) {
    return guarded([&]() -> vrod_status { return vrod_collection_fill_synthetic_impl(c, n, seed); });
}
static vrod_status vrod_collection_fill_synthetic_impl(vrod_collection *c, uint64_t n, uint64_t seed) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (c->count + n > c->capacity) return fail(VROD_ENOMEM, "collection capacity exceeded");
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaSetDevice(ctx->device));
    uint64_t loc0, cnt, off;
    shard_overlap(c, c->count, n, &loc0, &cnt, &off);
    if (cnt) {
        VROD_CUDA(launch_fill_synthetic(c->rows, (uint32_t)loc0, (uint32_t)cnt, c->dim, c->ld, c->id_base + loc0, seed,
                                        ctx->stream));
        ctx->stats.kernel_launches++;
    }
    vrod_status st = finish_append(c, loc0, cnt, false);
    if (st != VROD_OK) return st;
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    c->count += n;
    c->local += cnt;
    return VROD_OK;
}

static vrod_status vrod_collection_read_rows_impl(vrod_collection *c, uint64_t row0, uint64_t n, float *out);
extern "C" vrod_status vrod_collection_read_rows(vrod_collection *c, uint64_t row0, uint64_t n, float *out) {
    return guarded([&]() -> vrod_status { return vrod_collection_read_rows_impl(c, row0, n, out); });
}
static vrod_status vrod_collection_read_rows_impl(vrod_collection *c, uint64_t row0, uint64_t n, float *out) {
    if (!c || (!out && n)) return fail(VROD_EINVAL, "NULL argument");
    if (row0 + n > c->local) return fail(VROD_EINVAL, "row range outside this rank's shard");
    if (n == 0) return VROD_OK;
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaSetDevice(ctx->device));
    VROD_CUDA(cudaMemcpy2DAsync(out, (size_t)c->dim * sizeof(float), c->rows + (size_t)row0 * c->ld,
                                (size_t)c->ld * sizeof(float), (size_t)c->dim * sizeof(float), (size_t)n,
                                cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    return VROD_OK;
}

// -------------------------------------------------------------------------------------------------
// persistence
// -------------------------------------------------------------------------------------------------
struct ColFileHeader {
    char magic[8];        // "VRODCOL1"
    uint32_t dim, metric;
    uint64_t count;
    uint8_t reserved[40];
};
static_assert(sizeof(ColFileHeader) == 64, "header is 64 bytes");
constexpr uint64_t kIoChunkRows = 1u << 16;

static vrod_status vrod_collection_save_impl(vrod_collection *c, const char *path);
extern "C" vrod_status vrod_collection_save(vrod_collection *c, const char *path) {
    return guarded([&]() -> vrod_status { return vrod_collection_save_impl(c, path); });
}
static vrod_status vrod_collection_save_impl(vrod_collection *c, const char *path) {
    if (!c || !path) return fail(VROD_EINVAL, "NULL argument");
    vrod_ctx *ctx = c->ctx;
    if (ctx->world > 1) return fail(VROD_EINVAL, "vrod_collection_save: single-GPU contexts only");
    VROD_CUDA(cudaSetDevice(ctx->device));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(VROD_EINVAL, std::string("cannot open '") + path + "' for writing");
    ColFileHeader h{};
    memcpy(h.magic, "VRODCOL1", 8);
    h.dim = c->dim;
    h.metric = (uint32_t)c->metric;
    h.count = c->count;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    std::vector<float> buf((size_t)kIoChunkRows * c->dim);
    for (uint64_t r0 = 0; ok && r0 < c->local; r0 += kIoChunkRows) {
        const uint64_t n = c->local - r0 < kIoChunkRows ? c->local - r0 : kIoChunkRows;
        vrod_status st = vrod_collection_read_rows(c, r0, n, buf.data());
        if (st != VROD_OK) {
            fclose(f);
            return st;
        }
        ok = fwrite(buf.data(), sizeof(float) * c->dim, n, f) == n;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? VROD_OK : fail(VROD_EINVAL, std::string("short write to '") + path + "'");
}

static vrod_status vrod_collection_load_impl(vrod_ctx *ctx, const char *name, const char *path, uint64_t capacity_rows,
                                            vrod_collection **out);
extern "C" vrod_status vrod_collection_load(vrod_ctx *ctx, const char *name, const char *path, uint64_t capacity_rows,
                                            vrod_collection **out) {
    return guarded([&]() -> vrod_status { return vrod_collection_load_impl(ctx, name, path, capacity_rows, out); });
}
static vrod_status vrod_collection_load_impl(vrod_ctx *ctx, const char *name, const char *path, uint64_t capacity_rows,
                                            vrod_collection **out) {
    if (out) *out = nullptr;
    if (!ctx || !name || !path) return fail(VROD_EINVAL, "NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(VROD_ENOTFOUND, std::string("cannot open '") + path + "'");
    ColFileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "VRODCOL1", 8) != 0 || h.dim == 0 || h.metric > 1) {
        fclose(f);
        return fail(VROD_EINVAL, std::string("'") + path + "' is not a vrod collection file");
    }
    const uint64_t cap = capacity_rows > h.count ? capacity_rows : (h.count ? h.count : 1);
    vrod_collection *c = nullptr;
    vrod_status st = vrod_collection_create(ctx, name, h.dim, (vrod_metric)h.metric, cap, &c);
    if (st != VROD_OK) {
        fclose(f);
        return st;
    }
    // every rank appends the whole file in chunks; vrod_collection_insert keeps the rows of its own range
    // (a sharded rank still reads the whole file: simple, and the file is read once)
    std::vector<float> buf((size_t)kIoChunkRows * h.dim);
    for (uint64_t r0 = 0; r0 < h.count; r0 += kIoChunkRows) {
        const uint64_t n = h.count - r0 < kIoChunkRows ? h.count - r0 : kIoChunkRows;
        if (fread(buf.data(), sizeof(float) * h.dim, n, f) != n) {
            fclose(f);
            vrod_collection_drop(ctx, name);
            return fail(VROD_EINVAL, std::string("'") + path + "' is truncated");
        }
        st = vrod_collection_insert(c, buf.data(), n, nullptr);
        if (st != VROD_OK) {
            fclose(f);
            const std::string msg = g_last_error;
            vrod_collection_drop(ctx, name);
            return fail(st, msg);
        }
    }
    fclose(f);
    if (out) *out = c;
    return VROD_OK;
}

// -------------------------------------------------------------------------------------------------
// SEARCH
// -------------------------------------------------------------------------------------------------
static ShardView shard_view(const vrod_collection *c) {
    ShardView s{};
    s.rows = c->rows;
    s.inv_norm = c->inv_norm;
    s.sq_norm = c->sq_norm;
    s.maxnorm_bits = reinterpret_cast<const unsigned int *>(c->flags) + 1;
    s.n = (uint32_t)c->local;
    s.dim = c->dim;
    s.ld = c->ld;
    s.metric = (int)c->metric;
    s.id_base = c->id_base;
    return s;
}

// queries already on the device as [b x ld]; enqueue everything, no synchronisation
// d_status: per-query guard flags (device).  host_checks: the caller reads d_status after synchronising and
// rescans the flagged queries itself, so the conditional exact-scan launches are left out (single GPU only).
static vrod_status search_enqueue(vrod_collection *c, const float *d_q, uint32_t b, uint32_t k, uint64_t *d_ids,
                                  float *d_dist, bool force_exact, int *d_status = nullptr, bool host_checks = false,
                                  bool *used_scan = nullptr, int *d_xerr = nullptr) {
    vrod_ctx *ctx = c->ctx;
    if (!d_status) d_status = ctx_status(ctx);
    if (ctx->world > 1) host_checks = false;
    if (used_scan) *used_scan = false;
    const ShardView s = shard_view(c);
    const size_t nhits = (size_t)b * k;
    VROD_CUDA(ctx->hits_local.ensure(nhits * sizeof(Hit)));
    Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
    ScanScratch scr{reinterpret_cast<unsigned long long *>(ctx->blk_cand.p),
                    reinterpret_cast<unsigned int *>(ctx_ticket(ctx)), d_status, ctx->dev_counters};
    const bool exact_only = force_exact || c->path == 2 || !c->fast_ok;
    // single GPU: the scan kernels write the final arrays themselves, no merge launch
    const bool direct = ctx->world == 1;
    auto oid = [&](uint32_t qi) { return direct ? reinterpret_cast<unsigned long long *>(d_ids) + (size_t)qi * k : nullptr; };
    auto odd = [&](uint32_t qi) { return direct ? d_dist + (size_t)qi * k : nullptr; };
    // automatic choice between b single-query scans and one batched (tensor-core) pass: a crude cost model from the
    // round-1 measurements -- a scan costs ~25 us + its HBM bytes at 6.5 TB/s, a batched pass ~200 us of phase and
    // launch overhead + 1.45 us per (128-row tile x 144 bf16 columns x query group of 256) spread over the SMs
    // (configs[2]: 2111 tiles per SM in 3.2 ms)
    bool prefer_batched = false;
    if (b >= 2) {
        const double bytes = (double)s.n * s.ld * 4.0;
        const double t_scan = 25e-6 + bytes / 6.5e12;
        const double groups = (double)((b + 255) / 256);
        const double t_batched = 200e-6 + groups * ((double)s.n / 128.0) * 1.45e-6 * ((double)mirror_ld(s.dim) / 144.0) / (double)ctx->sms;
        prefer_batched = (1.0 - c->rescan_share) * (double)b * t_scan > t_batched;
        if (!prefer_batched) c->rescan_share *= 0.995;   // forget slowly: the batched pass is probed again later
    }
    const bool batched = !exact_only && batched_supported(s, b, k) && (c->path == 3 || c->path == 4 || (c->path == 0 && prefer_batched));
    if (batched) {
        ShardView sb = s;
        if (c->path != 4 && !c->mirror_failed) {
            // bf16 operand mirror: allocate once for the shard's capacity, convert the rows appended since the last time
            if (!c->rows_h) {
                // VROD_NO_MIRROR=1 behaves like a failed allocation (tests of the fallback; memory-tight deployments)
                static const bool no_mirror = getenv("VROD_NO_MIRROR") != nullptr;
                const cudaError_t me = no_mirror ? cudaErrorMemoryAllocation
                                                 : cudaMalloc(&c->rows_h, (size_t)c->shard_rows * mirror_ld(c->dim) * sizeof(unsigned short));
                if (me != cudaSuccess) {
                    cudaGetLastError();
                    c->rows_h = nullptr;
                    c->mirror_failed = true;
                }
                c->mirror_rows = 0;
            }
            if (c->rows_h) {
                if (c->mirror_rows < c->local) {
                    VROD_CUDA(launch_build_mirror(s, c->rows_h, (uint32_t)c->mirror_rows, (uint32_t)(c->local - c->mirror_rows), ctx->stream));
                    ctx->stats.kernel_launches++;
                    c->mirror_rows = c->local;
                }
                sb.rows_h = c->rows_h;
            }
        }
        BatchedStats bs{};
        cudaEvent_t e0 = ctx->profiling ? ctx->prof_event() : nullptr, e1 = ctx->profiling ? ctx->prof_event() : nullptr;
        cudaError_t e = launch_batched_search(sb, d_q, b, k, ctx->sms, &ctx->batched.p, &ctx->batched.bytes, d_status,
                                              local, oid(0), odd(0), ctx->stream, &bs, e0, e1);
        if (e != cudaSuccess) return fail(VROD_ECUDA, std::string("batched search: ") + cudaGetErrorString(e));
        ctx->stats.kernel_launches += bs.launches;
        ctx->stats.batched_tiles += bs.tiles;
        // queries whose guard failed are answered by the single-query scan.  The count is only known on
        // the device, so this path synchronises once (the shard-local rescans involve no collective).
        VROD_CUDA(ctx->status_host.ensure(sizeof(int) * b));
        int *hs = reinterpret_cast<int *>(ctx->status_host.p);
        VROD_CUDA(cudaMemcpyAsync(hs, d_status, sizeof(int) * b, cudaMemcpyDeviceToHost, ctx->stream));
        VROD_CUDA(cudaStreamSynchronize(ctx->stream));
        const ScanPlan fp = make_scan_plan(s, k, ctx->sms, false);
        const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
        uint32_t flagged = 0;
        for (uint32_t qi = 0; qi < b; ++qi) flagged += hs[qi] ? 1u : 0u;
        {   // rises at once, falls slowly
            const double share = (double)flagged / (double)b, ema = 0.5 * c->rescan_share + 0.5 * share;
            c->rescan_share = share > ema ? share : ema;
        }
        for (uint32_t qi = 0; qi < b; ++qi) {
            if (!hs[qi]) continue;
            const float *q = d_q + (size_t)qi * s.ld;
            VROD_CUDA(launch_fast_scan(s, q, k, fp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi), ctx->stream));
            VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi), ctx->stream));
            ctx->stats.kernel_launches += 2;
            ctx->stats.fast_scans++;
        }
    } else if (exact_only) {
        const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
        for (uint32_t qi = 0; qi < b; ++qi) {
            VROD_CUDA(launch_exact_scan(s, d_q + (size_t)qi * s.ld, k, xp, scr, nullptr, local + (size_t)qi * k, oid(qi), odd(qi),
                                        ctx->stream));
            ctx->stats.kernel_launches++;
        }
        ctx->stats.exact_rescans += 0;
    } else {
        const ScanPlan fp = make_scan_plan(s, k, ctx->sms, false);
        const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
        for (uint32_t qi = 0; qi < b; ++qi) {
            const float *q = d_q + (size_t)qi * s.ld;
            static const bool scan_dbg = getenv("VROD_SCAN_DEBUG") != nullptr;
            unsigned long long *dbgbuf = nullptr;
            if (scan_dbg) {
                dbgbuf = scan_debug_enable();
                const unsigned long long init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
                cudaMemcpyAsync(dbgbuf, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
            }
            if (ctx->profiling) VROD_CUDA(cudaEventRecord(ctx->prof_event(), ctx->stream));
            VROD_CUDA(launch_fast_scan(s, q, k, fp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi),
                                       ctx->stream));
            if (scan_dbg) {
                unsigned long long h[8];
                cudaStreamSynchronize(ctx->stream);
                cudaMemcpy(h, dbgbuf, sizeof(h), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[scan dbg] scan loops %.1f us | wait+ticket %.1f | merge %.1f | rerank %.1f | sort+write %.1f | total %.1f us\n",
                        (h[1] - h[0]) / 1e3, (h[2] - h[1]) / 1e3, (h[3] - h[2]) / 1e3, (h[4] - h[3]) / 1e3, (h[5] - h[4]) / 1e3,
                        (h[5] - h[0]) / 1e3);
            }
            if (ctx->profiling) VROD_CUDA(cudaEventRecord(ctx->prof_event(), ctx->stream));
            if (!host_checks) {
                VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi),
                                            ctx->stream));
                ctx->stats.kernel_launches++;
            }
            ctx->stats.kernel_launches++;
        }
        ctx->stats.fast_scans += b;
        if (used_scan) *used_scan = true;
    }
    const Hit *lists = local;
    uint32_t g = 1;
    if (ctx->world > 1 && ctx->fused_exchange && b <= kXchgMaxB && nhits <= kXchgMaxHits) {
        // one kernel: push the local lists into every rank's window over NVLink, wait for the peers', merge
        VROD_CUDA(launch_exchange_merge(ctx->d_windows, (uint32_t)ctx->rank, (uint32_t)ctx->world, ++ctx->xchg_seq, local, b, k,
                                        reinterpret_cast<unsigned long long *>(d_ids), d_dist, d_xerr ? d_xerr : ctx->d_xchg_err,
                                        ctx->stream));
        ctx->stats.kernel_launches++;
        ctx->stats.searches += b;
        return VROD_OK;
    }
    if (ctx->world > 1) {
        VROD_CUDA(ctx->hits_all.ensure(nhits * sizeof(Hit) * ctx->world));
        VROD_NCCL(g_nccl.AllGather(local, ctx->hits_all.p, nhits * sizeof(Hit), ncclChar, ctx->comm, ctx->stream));
        lists = reinterpret_cast<const Hit *>(ctx->hits_all.p);
        g = (uint32_t)ctx->world;
    }
    if (!direct) {
        VROD_CUDA(launch_merge_hits(lists, g, b, k, reinterpret_cast<unsigned long long *>(d_ids), d_dist, ctx->stream));
        ctx->stats.kernel_launches++;
    }
    ctx->stats.searches += b;
    return VROD_OK;
}

static vrod_status check_search_args(vrod_collection *c, const void *q, uint32_t b, uint32_t k, const void *ids,
                                     const void *dist) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (b == 0) return VROD_OK;
    if (!q || !ids || !dist) return fail(VROD_EINVAL, "NULL buffer");
    if (k == 0 || k > VROD_MAX_K) return fail(VROD_EINVAL, "k must be in [1, VROD_MAX_K]");
    if (b > kMaxBatch) return fail(VROD_EINVAL, "more than 65536 queries in one call");
    return VROD_OK;
}

static vrod_status vrod_collection_search_device_impl(vrod_collection *c, const float *d_queries, uint32_t b, uint32_t k,
                                                     uint64_t *d_out_ids, float *d_out_dist);
extern "C" vrod_status vrod_collection_search_device(vrod_collection *c, const float *d_queries, uint32_t b, uint32_t k,
                                                     uint64_t *d_out_ids, float *d_out_dist) {
    return guarded([&]() -> vrod_status { return vrod_collection_search_device_impl(c, d_queries, b, k, d_out_ids, d_out_dist); });
}
static vrod_status vrod_collection_search_device_impl(vrod_collection *c, const float *d_queries, uint32_t b, uint32_t k,
                                                     uint64_t *d_out_ids, float *d_out_dist) {
    vrod_status st = check_search_args(c, d_queries, b, k, d_out_ids, d_out_dist);
    if (st != VROD_OK || b == 0) return st;
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaSetDevice(ctx->device));
    const float *q = d_queries;
    if (c->ld != c->dim) {
        VROD_CUDA(ctx->q_pad.ensure((size_t)b * c->ld * sizeof(float)));
        VROD_CUDA(launch_pad_queries(d_queries, reinterpret_cast<float *>(ctx->q_pad.p), b, c->dim, c->ld, ctx->stream));
        ctx->stats.kernel_launches++;
        q = reinterpret_cast<const float *>(ctx->q_pad.p);
    }
    return search_enqueue(c, q, b, k, d_out_ids, d_out_dist, false);
}

static vrod_status vrod_collection_search_impl(vrod_collection *c, const float *queries, uint32_t b, uint32_t k,
                                              uint64_t *out_ids, float *out_dist);
extern "C" vrod_status vrod_collection_search(vrod_collection *c, const float *queries, uint32_t b, uint32_t k,
                                              uint64_t *out_ids, float *out_dist) {
    return guarded([&]() -> vrod_status { return vrod_collection_search_impl(c, queries, b, k, out_ids, out_dist); });
}
static vrod_status vrod_collection_search_impl(vrod_collection *c, const float *queries, uint32_t b, uint32_t k,
                                              uint64_t *out_ids, float *out_dist) {
    vrod_status st = check_search_args(c, queries, b, k, out_ids, out_dist);
    if (st != VROD_OK || b == 0) return st;
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaSetDevice(ctx->device));
    // validate + pack (zero padded to ld) into pinned staging
    const size_t qbytes = (size_t)b * c->ld * sizeof(float);
    VROD_CUDA(ctx->q_host.ensure(qbytes));
    float *qh = reinterpret_cast<float *>(ctx->q_host.p);
    // queries outside the range the f32 / tensor-core passes' error bounds cover are answered by the exact f64 scan:
    // each on its own on a single GPU (the rest of the batch keeps its fast path), the whole call on sharded contexts
    // (every rank must enqueue the same collectives)
    std::vector<unsigned char> unsafe_q(b, 0);
    uint32_t n_unsafe = 0;
    for (uint32_t i = 0; i < b; ++i) {
        const float *src = queries + (size_t)i * c->dim;
        float *dst = qh + (size_t)i * c->ld;
        double nq = 0.0;
        bool u = false;
        for (uint32_t j = 0; j < c->dim; ++j) {
            const float v = src[j];
            if (!isfinite(v)) return fail(VROD_EINVAL, "query contains NaN or infinity");
            if (fabsf(v) > 0x1p40f) u = true;
            nq += (double)v * (double)v;
            dst[j] = v;
        }
        for (uint32_t j = c->dim; j < c->ld; ++j) dst[j] = 0.f;
        if (nq > 0.0 && (nq < 0x1p-80 || nq > 0x1p100)) u = true;
        if (c->metric == VROD_COSINE && nq == 0.0) u = true;  // all distances are 1: answered by the exact scan
        unsafe_q[i] = u ? 1 : 0;
        n_unsafe += u ? 1 : 0;
    }
    const bool unsafe = n_unsafe != 0 && (ctx->world > 1 || n_unsafe == b);   // the whole call goes the exact way
    // one device buffer [ids | dist | status] so that the results and the guard flags come back in ONE copy
    const size_t nres = (size_t)b * k;
    const size_t off_dist = nres * sizeof(uint64_t);
    const size_t off_stat = off_dist + ((nres * sizeof(float) + 15) & ~(size_t)15);
    const size_t pack = off_stat + ((size_t)b + 1) * sizeof(int);   // + 1: peer-exchange error flag (sharded contexts)
    VROD_CUDA(ctx->q_dev.ensure(qbytes));
    VROD_CUDA(ctx->out_ids.ensure(pack));
    VROD_CUDA(ctx->ids_host.ensure(pack));
    unsigned char *dpack = reinterpret_cast<unsigned char *>(ctx->out_ids.p);
    unsigned char *hpack = reinterpret_cast<unsigned char *>(ctx->ids_host.p);
    uint64_t *d_ids = reinterpret_cast<uint64_t *>(dpack);
    float *d_dist = reinterpret_cast<float *>(dpack + off_dist);
    int *d_stat = reinterpret_cast<int *>(dpack + off_stat);
    VROD_CUDA(cudaMemcpyAsync(ctx->q_dev.p, qh, qbytes, cudaMemcpyHostToDevice, ctx->stream));
    bool used_scan = false;
    st = search_enqueue(c, reinterpret_cast<const float *>(ctx->q_dev.p), b, k, d_ids, d_dist, unsafe, d_stat, true, &used_scan,
                        d_stat + b);
    if (st != VROD_OK) return st;
    VROD_CUDA(cudaMemcpyAsync(hpack, dpack, pack, cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->fused_exchange && reinterpret_cast<const int *>(hpack + off_stat)[b] == 1)
        return fail(VROD_ENCCL, "peer exchange timed out: a rank did not take part in this search");
    if (ctx->world == 1 && !unsafe && (used_scan || n_unsafe)) {
        // the scans ran without their device-side conditional rescans: answer the flagged queries exactly now,
        // and with them the queries that were out of range for the fast paths
        int *hs = reinterpret_cast<int *>(hpack + off_stat);
        bool any = false;
        for (uint32_t qi = 0; qi < b; ++qi) {
            hs[qi] = (used_scan && hs[qi] != 0) || unsafe_q[qi] ? 1 : 0;
            any = any || hs[qi] != 0;
        }
        if (any) {
            const ShardView s = shard_view(c);
            const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
            ScanScratch scr{reinterpret_cast<unsigned long long *>(ctx->blk_cand.p),
                            reinterpret_cast<unsigned int *>(ctx_ticket(ctx)), d_stat, ctx->dev_counters};
            Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
            for (uint32_t qi = 0; qi < b; ++qi) {
                if (!hs[qi]) continue;
                VROD_CUDA(launch_exact_scan(s, reinterpret_cast<const float *>(ctx->q_dev.p) + (size_t)qi * s.ld, k, xp, scr, nullptr,
                                            local + (size_t)qi * k, reinterpret_cast<unsigned long long *>(d_ids) + (size_t)qi * k,
                                            d_dist + (size_t)qi * k, ctx->stream));
                ctx->stats.kernel_launches++;
            }
            VROD_CUDA(cudaMemcpyAsync(hpack, dpack, off_stat, cudaMemcpyDeviceToHost, ctx->stream));
            VROD_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    memcpy(out_ids, hpack, nres * sizeof(uint64_t));
    memcpy(out_dist, hpack + off_dist, nres * sizeof(float));
    ctx->stats.h2d_bytes += qbytes;
    ctx->stats.d2h_bytes += pack;
    return VROD_OK;
}

extern "C" const char *vrod_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *vrod_version(void) { return "vrod_knn_b200 0.1.0 (sm_100a)"; }
