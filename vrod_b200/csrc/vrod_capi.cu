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
#include "knn_device.cuh"
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

// Scan-versus-batched cost model of the automatic path choice.  Seeded with figures measured on one box in round 1
// (a scan = ~25 us + its bytes at 6.5 TB/s; a batched pass = ~200 us of phase and launch overhead + 1.45 us per
// 128-row tile x 144 bf16 columns x 256-query group, spread over the SMs), then corrected by what THIS context
// measures: every 16th search brackets its scan launch / batched pass with two events of its own; once they have
// completed (polled at a later search, never waited for) the elapsed time is folded into the matching term.  The
// boxes of this pool differed by 1537..1965 MHz under load, which the seeds alone would not follow.
struct CostModel {
    double scan_fixed = 20e-6, hbm_bps = 6.5e12, batched_fixed = 150e-6, tile_seconds = 1.2e-6;   // seeds: round-2 measurements
    uint32_t calls = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int pending_kind = 0;      // 0 none, 1 scan, 2 batched
    double pending_work = 0.0;
    double scan_seconds(double bytes) const { return scan_fixed + bytes / hbm_bps; }
    double batched_seconds(double units) const { return batched_fixed + units * tile_seconds; }
    bool want_sample() const { return pending_kind == 0 && (calls & 15u) == 1u; }
    bool events(cudaEvent_t *e0, cudaEvent_t *e1) {
        for (cudaEvent_t &e : ev)
            if (!e && cudaEventCreate(&e) != cudaSuccess) {
                cudaGetLastError();
                e = nullptr;
                return false;
            }
        *e0 = ev[0];
        *e1 = ev[1];
        return true;
    }
    void pending(int kind, double work) {
        pending_kind = kind;
        pending_work = work;
    }
    static double blend(double old_v, double new_v, double seed) {
        if (!(new_v > 0.25 * seed)) new_v = 0.25 * seed;   // one odd sample (a clock ramp, a page fault) must not swing the choice
        if (new_v > 4.0 * seed) new_v = 4.0 * seed;
        return 0.75 * old_v + 0.25 * new_v;
    }
    void collect() {
        ++calls;
        if (pending_kind == 0 || cudaEventQuery(ev[1]) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess && ms > 0.f) {
            const double t = ms * 1e-3;
            if (pending_kind == 1) {
                const double stream_t = pending_work / hbm_bps;
                if (stream_t > 4.0 * scan_fixed) hbm_bps = blend(hbm_bps, pending_work / (t > scan_fixed ? t - scan_fixed : t), 6.5e12);
                else if (stream_t < scan_fixed) scan_fixed = blend(scan_fixed, t - stream_t, 20e-6);
            } else {
                const double tiles_t = pending_work * tile_seconds;
                if (tiles_t > 4.0 * batched_fixed) tile_seconds = blend(tile_seconds, (t > batched_fixed ? t - batched_fixed : t) / pending_work, 1.2e-6);
                else if (tiles_t < batched_fixed) batched_fixed = blend(batched_fixed, t - tiles_t, 150e-6);
            }
        } else {
            cudaGetLastError();
        }
        pending_kind = 0;
    }
    void release() {
        for (cudaEvent_t &e : ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
    }
};

struct vrod_ctx {
    int device = 0, rank = 0, world = 1, sms = 0;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    std::map<std::string, vrod_collection *> colls;
    // scratch shared by every collection of the context
    DevBuf blk_cand, small, q_pad, q_dev, hits_local, hits_all, out_ids, out_dist, batched, agree;
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
    CostModel cost;
    cudaEvent_t ev_status = nullptr;            // marks the guard flags' arrival in status_host (batched path)
    // single-process multi-GPU context (vrod_ctx_create_multi): the parent owns one sub-context per device (sub r is
    // "rank r of subs.size()": same shard rule, same kernels) and drives them all from the caller's one thread
    std::vector<vrod_ctx *> subs;               // empty for single-GPU and per-rank contexts
    vrod_ctx *parent = nullptr;
    cudaEvent_t ev_done = nullptr;              // sub: this device's hits have been copied to device 0 (gather path)
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
    // the batched pass keeps every candidate inside the approximate surrogate's error band (knn_batched.cuh: band_mode)
    // instead of a fixed k' of them: switched on, for good, by the first batch whose fixed-k' proofs fail for more than
    // 5 % of the queries (tight clusters under the Euclidean metric)
    bool band_mode = false;
    // the batched pass filters its phases at guessed thresholds (knn_batched.cuh: guess) until a batch shows that this
    // collection's row order defeats the guesses (rows inserted cluster by cluster, sorted rows)
    bool guess_off = false;
    // the batched pass keeps k' = 2k + 16 (pow2) candidates instead of 1.5k + 16: switched on by the first batch in which
    // more than 5 % of the proofs fail at the tight margin; band mode is the step after that
    bool wide_margin = false;
    std::vector<vrod_collection *> parts;   // collection of a multi-GPU parent context: one part per device, in id order
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

// Single-process multi-GPU context (the reference's caller is ONE single-threaded
// process -- Rc<RefCell<Database>>, src/command/types.rs:10; fn main, src/main.rs:42).  One sub-context per device
// ("rank r of n": same shard rule, same kernels as the process-per-GPU mode), all driven from the caller's thread.
// The devices' exchange windows are plain cudaMalloc memory reached through direct peer access (no IPC needed inside
// one process); without peer access the per-device lists are gathered on device 0 with cudaMemcpyPeerAsync instead.
static vrod_status vrod_ctx_create_multi_impl(const int *device_ids, int n_devices, vrod_ctx **out);
extern "C" vrod_status vrod_ctx_create_multi(const int *device_ids, int n_devices, vrod_ctx **out) {
    return guarded([&]() -> vrod_status { return vrod_ctx_create_multi_impl(device_ids, n_devices, out); });
}
static vrod_status vrod_ctx_create_multi_impl(const int *device_ids, int n_devices, vrod_ctx **out) {
    if (!out) return fail(VROD_EINVAL, "out is NULL");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > (int)kXchgMaxWorld) return fail(VROD_EINVAL, "need 1..16 device ordinals");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return fail(VROD_EINVAL, "a device ordinal is listed twice");
    if (n_devices == 1) return vrod_ctx_create_impl(device_ids[0], out);
    vrod_ctx *parent = new (std::nothrow) vrod_ctx();
    if (!parent) return fail(VROD_ENOMEM, "host allocation failed");
    struct Undo {   // any early return tears down what was built (the error message is kept)
        vrod_ctx *p;
        ~Undo() {
            if (!p) return;
            const std::string msg = g_last_error;
            vrod_ctx_destroy(p);
            g_last_error = msg;
        }
    } undo{parent};
    parent->device = device_ids[0];
    const int W = n_devices;
    for (int r = 0; r < W; ++r) {
        vrod_ctx *sub = new (std::nothrow) vrod_ctx();
        if (!sub) return fail(VROD_ENOMEM, "host allocation failed");
        parent->subs.push_back(sub);
        sub->parent = parent;
        vrod_status st = ctx_init(sub, device_ids[r]);
        if (st != VROD_OK) return st;
        sub->rank = r;
        sub->world = W;
        VROD_CUDA(cudaEventCreateWithFlags(&sub->ev_done, cudaEventDisableTiming));
    }
    parent->sms = parent->subs[0]->sms;
    // direct peer access between every pair of devices (NVLink through the NVSwitch on an HGX box)
    bool peer_ok = getenv("VROD_NO_P2P_EXCHANGE") == nullptr;
    for (int i = 0; i < W && peer_ok; ++i) {
        VROD_CUDA(cudaSetDevice(device_ids[i]));
        for (int j = 0; j < W && peer_ok; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]) != cudaSuccess || !can) {
                peer_ok = false;
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
            cudaGetLastError();
        }
    }
    if (peer_ok) {
        std::vector<unsigned char *> wins(W, nullptr);
        for (int r = 0; r < W; ++r) {
            vrod_ctx *sub = parent->subs[r];
            VROD_CUDA(cudaSetDevice(sub->device));
            VROD_CUDA(cudaMalloc(&sub->xchg, xchg_window_bytes()));
            VROD_CUDA(cudaMemsetAsync(sub->xchg, 0, xchg_window_bytes(), sub->stream));
            VROD_CUDA(cudaMalloc(&sub->d_xchg_err, 64));
            VROD_CUDA(cudaMemsetAsync(sub->d_xchg_err, 0, 64, sub->stream));
            wins[r] = sub->xchg;
        }
        for (int r = 0; r < W; ++r) {
            vrod_ctx *sub = parent->subs[r];
            VROD_CUDA(cudaSetDevice(sub->device));
            VROD_CUDA(cudaMalloc(&sub->d_windows, sizeof(unsigned char *) * W));
            VROD_CUDA(cudaMemcpyAsync(sub->d_windows, wins.data(), sizeof(unsigned char *) * W, cudaMemcpyHostToDevice, sub->stream));
            VROD_CUDA(cudaStreamSynchronize(sub->stream));
            sub->fused_exchange = true;
        }
    }
    if (getenv("VROD_VERBOSE"))
        fprintf(stderr, "[vrod] single-process context over %d devices: %s\n", W,
                peer_ok ? "fused NVLink exchange through direct peer access" : "no peer access, lists gathered with cudaMemcpyPeerAsync");
    VROD_CUDA(cudaSetDevice(parent->device));
    undo.p = nullptr;
    *out = parent;
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
    if (!ctx->subs.empty()) {
        // parent of a single-process multi-GPU context: its collections are thin handles over the subs' parts
        for (auto &kv : ctx->colls) delete kv.second;
        ctx->colls.clear();
        for (vrod_ctx *sub : ctx->subs) vrod_ctx_destroy(sub);
        ctx->subs.clear();
        cudaSetDevice(ctx->device);
        PinBuf *pin[] = {&ctx->q_host, &ctx->ids_host, &ctx->dist_host, &ctx->status_host};
        for (PinBuf *b : pin) b->release();
        delete ctx;
        return;
    }
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
                     &ctx->hits_all, &ctx->out_ids, &ctx->out_dist, &ctx->batched, &ctx->agree};
    for (DevBuf *b : dev) b->release();
    PinBuf *pin[] = {&ctx->q_host, &ctx->ids_host, &ctx->dist_host, &ctx->status_host};
    for (PinBuf *b : pin) b->release();
    cudaFree(ctx->dev_counters);
    ctx->cost.release();
    if (ctx->ev_status) cudaEventDestroy(ctx->ev_status);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
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
    if (!ctx->subs.empty()) {
        for (vrod_ctx *sub : ctx->subs) {
            vrod_status st = vrod_ctx_synchronize_impl(sub);
            if (st != VROD_OK) return st;
        }
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
    VROD_CUDA(cudaSetDevice(ctx->device));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->fused_exchange && !ctx->parent) {
        // searches enqueued through the device-pointer API report a peer that never arrived only here
        int err = 0;
        VROD_CUDA(cudaMemcpy(&err, ctx->d_xchg_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (err == 1) {
            VROD_CUDA(cudaMemset(ctx->d_xchg_err, 0, sizeof(int)));
            return fail(VROD_ENCCL, "peer exchange timed out: a rank did not take part in a search");
        }
    }
    return VROD_OK;
}
extern "C" void *vrod_ctx_stream(vrod_ctx *ctx) {
    if (!ctx) return nullptr;
    return ctx->subs.empty() ? (void *)ctx->stream : (void *)ctx->subs[0]->stream;
}
extern "C" int vrod_ctx_rank(vrod_ctx *ctx) { return ctx ? ctx->rank : -1; }
extern "C" int vrod_ctx_world(vrod_ctx *ctx) { return ctx ? ctx->world : -1; }
extern "C" int vrod_ctx_devices(vrod_ctx *ctx) { return ctx ? (ctx->subs.empty() ? 1 : (int)ctx->subs.size()) : -1; }

static vrod_status vrod_ctx_profile_impl(vrod_ctx *ctx, int enable);
extern "C" vrod_status vrod_ctx_profile(vrod_ctx *ctx, int enable) {
    return guarded([&]() -> vrod_status { return vrod_ctx_profile_impl(ctx, enable); });
}
static vrod_status vrod_ctx_profile_impl(vrod_ctx *ctx, int enable) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    ctx->profiling = enable != 0;
    for (vrod_ctx *sub : ctx->subs) sub->profiling = enable != 0;
    return VROD_OK;
}

static vrod_status vrod_ctx_profile_read_impl(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches);
extern "C" vrod_status vrod_ctx_profile_read(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches) {
    return guarded([&]() -> vrod_status { return vrod_ctx_profile_read_impl(ctx, kernel_ms, launches); });
}
static vrod_status vrod_ctx_profile_read_impl(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches) {
    if (!ctx) return fail(VROD_EINVAL, "ctx is NULL");
    if (!ctx->subs.empty()) {   // the slowest device's kernels (the devices run side by side)
        double worst = -1.0;
        uint64_t n = 0;
        for (vrod_ctx *sub : ctx->subs) {
            double ms = 0.0;
            uint64_t l = 0;
            vrod_status st = vrod_ctx_profile_read_impl(sub, &ms, &l);
            if (st != VROD_OK) return st;
            if (ms > worst) {
                worst = ms;
                n = l;
            }
        }
        if (kernel_ms) *kernel_ms = worst < 0.0 ? 0.0 : worst;
        if (launches) *launches = n;
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
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
    if (!ctx->subs.empty()) {   // searches / copies are counted by the parent, the kernels by the devices
        vrod_stats sum = ctx->stats;
        for (vrod_ctx *sub : ctx->subs) {
            vrod_stats t{};
            vrod_status st = vrod_ctx_stats_impl(sub, &t);
            if (st != VROD_OK) return st;
            sum.kernel_launches += t.kernel_launches;
            if (t.fast_scans > sum.fast_scans) sum.fast_scans = t.fast_scans;   // every device scans every query: not a sum
            sum.exact_rescans += t.exact_rescans;
            sum.batched_tiles += t.batched_tiles;
        }
        *out = sum;
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
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
// Names end up as file names (Database::save: name + ".vrc") and as whitespace-delimited tokens of vr_config, so
// the set is closed here, at the one place every collection is born.
static bool valid_collection_name(const char *name) {
    const size_t n = strlen(name);
    if (n == 0 || n > 200 || strcmp(name, ".") == 0 || strcmp(name, "..") == 0) return false;
    for (size_t i = 0; i < n; ++i) {
        const char ch = name[i];
        const bool ok = (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z') || (ch >= '0' && ch <= '9') || ch == '_' || ch == '.' || ch == '-';
        if (!ok) return false;
    }
    return true;
}

static vrod_status vrod_collection_drop_impl(vrod_ctx *ctx, const char *name);
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
    if (!valid_collection_name(name))
        return fail(VROD_EINVAL, "collection names are 1..200 characters of [A-Za-z0-9_.-] and neither '.' nor '..'");
    if (dim == 0 || dim > (1u << 20)) return fail(VROD_EINVAL, "dim must be in [1, 2^20]");
    if (metric != VROD_EUCLIDEAN && metric != VROD_COSINE) return fail(VROD_EINVAL, "unknown metric");
    if (capacity_rows == 0) return fail(VROD_EINVAL, "capacity_rows must be > 0");
    if (ctx->colls.count(name)) return fail(VROD_EEXISTS, std::string("collection '") + name + "' already exists");
    if (!ctx->subs.empty()) {
        // multi-GPU parent: a thin handle over one part per device (part r = shard r of the same capacity)
        std::unique_ptr<vrod_collection> pc(new (std::nothrow) vrod_collection());
        if (!pc) return fail(VROD_ENOMEM, "host allocation failed");
        pc->ctx = ctx;
        pc->name = name;
        pc->dim = dim;
        pc->ld = (dim + 3u) & ~3u;
        pc->metric = metric;
        pc->capacity = capacity_rows;
        for (vrod_ctx *sub : ctx->subs) {
            vrod_collection *part = nullptr;
            vrod_status st = vrod_collection_create_impl(sub, name, dim, metric, capacity_rows, &part);
            if (st != VROD_OK) {
                const std::string msg = g_last_error;
                for (vrod_ctx *s2 : ctx->subs)
                    if (s2->colls.count(name)) vrod_collection_drop_impl(s2, name);
                return fail(st, msg);
            }
            pc->parts.push_back(part);
        }
        VROD_CUDA(cudaSetDevice(ctx->device));
        vrod_collection *raw = pc.release();
        ctx->colls[name] = raw;
        if (out) *out = raw;
        return VROD_OK;
    }
    VROD_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<vrod_collection> c(new (std::nothrow) vrod_collection());
    if (!c) return fail(VROD_ENOMEM, "host allocation failed");
    c->ctx = ctx;
    c->name = name;
    c->dim = dim;
    c->ld = (dim + 3u) & ~3u;
    c->metric = metric;
    c->capacity = capacity_rows;
    // block-cyclic sharding (knn_scan.cuh: row_id): this rank's capacity is what it holds when the collection is full
    c->shard_rows = shard_rows_at(capacity_rows, (uint32_t)ctx->rank, (uint32_t)ctx->world);
    if (c->shard_rows >= 0xFFFFFFFFull) return fail(VROD_EINVAL, "more than 2^32-1 rows per GPU");
    if (c->shard_rows == 0) c->shard_rows = 1;      // (a rank beyond the last block of a small collection: keep the buffers real)
    c->id_base = (uint64_t)ctx->rank * kShardBlock;  // global id of local row 0
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
    if (!ctx->subs.empty()) {
        const std::string nm = name;   // `name` may point into the handle that is deleted here
        for (vrod_ctx *sub : ctx->subs)
            if (sub->colls.count(nm)) vrod_collection_drop_impl(sub, nm.c_str());
        delete it->second;
        ctx->colls.erase(it);
        cudaSetDevice(ctx->device);
        return VROD_OK;
    }
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
    if (id_base) *id_base = c->id_base;    // a multi-GPU parent holds everything: base 0, all rows
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
    for (vrod_collection *part : c->parts) part->path = path;
    return VROD_OK;
}

// Rows [g0, g0 + n) of the global sequence, block by block: calls f(local_row, source_offset, rows) for every piece that
// is dealt to this rank (knn_scan.cuh: block-cyclic sharding).  Appends arrive in id order, so the pieces of one call are
// back to back in the shard.
template <typename F>
static vrod_status for_my_pieces(const vrod_collection *c, uint64_t g0, uint64_t n, F &&f) {
    const uint64_t G = (uint64_t)c->ctx->world, r = (uint64_t)c->ctx->rank, B = kShardBlock;
    for (uint64_t blk = g0 / B; blk * B < g0 + n; ++blk) {
        if (blk % G != r) continue;
        const uint64_t a = blk * B > g0 ? blk * B : g0;
        const uint64_t e = (blk + 1) * B < g0 + n ? (blk + 1) * B : g0 + n;
        vrod_status st = f((blk / G) * B + (a - blk * B), a - g0, e - a);
        if (st != VROD_OK) return st;
    }
    return VROD_OK;
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

// Make room for at least `need` rows in all (the global capacity at least doubles).  Every shard grows on its own: the
// block-cyclic deal just continues, no row changes its shard, so a sharded collection needs no communication here.
static vrod_status collection_grow(vrod_collection *c, uint64_t need) {
    vrod_ctx *ctx = c->ctx;
    const uint32_t rk = (uint32_t)ctx->rank, G = (uint32_t)ctx->world;
    uint64_t gcap = c->capacity * 2 > need ? c->capacity * 2 : need;      // global
    if (shard_rows_at(gcap, 0, G) >= 0xFFFFFFFFull) gcap = need;
    if (shard_rows_at(gcap, 0, G) >= 0xFFFFFFFFull) return fail(VROD_ENOMEM, "more than 2^32-2 rows on one GPU");
    uint64_t cap = shard_rows_at(gcap, rk, G);                            // this shard
    if (cap == 0) cap = 1;
    const uint64_t need_local = shard_rows_at(need, rk, G) ? shard_rows_at(need, rk, G) : 1;
    float *rows = nullptr, *inv = nullptr, *sq = nullptr;
    cudaError_t e = cudaSuccess;
    for (int attempt = 0; attempt < 2; ++attempt) {   // doubled capacity first, exactly `need` rows if that does not fit
        e = cudaMalloc(&rows, (size_t)cap * c->ld * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&inv, (size_t)cap * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&sq, (size_t)cap * sizeof(float));
        if (e != cudaErrorMemoryAllocation || cap == need_local) break;
        cudaGetLastError();
        cudaFree(rows); cudaFree(inv); cudaFree(sq);
        rows = inv = sq = nullptr;
        cap = need_local;
        gcap = need;
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
    c->capacity = gcap;
    c->shard_rows = cap;
    return VROD_OK;
}

// a multi-GPU parent handle mirrors the global counters of its parts (every part tracks them)
static void sync_parent(vrod_collection *pc) {
    pc->count = pc->parts[0]->count;
    pc->capacity = pc->parts[0]->capacity;
    pc->local = pc->count;
}

static vrod_status vrod_collection_insert_impl(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id, bool validated = false);
extern "C" vrod_status vrod_collection_insert(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id) {
    return guarded([&]() -> vrod_status { return vrod_collection_insert_impl(c, rows, n, first_id); });
}
static vrod_status vrod_collection_insert_impl(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id, bool validated) {
    if (!c) return fail(VROD_EINVAL, "collection is NULL");
    if (n == 0) {
        if (first_id) *first_id = c->count;
        return VROD_OK;
    }
    if (!rows) return fail(VROD_EINVAL, "rows is NULL");
    vrod_ctx *ctx = c->ctx;
    if (!c->parts.empty()) {
        // every device takes the rows of its own id range; accept or reject is decided once, here, for all of them
        if (!host_rows_finite(rows, (size_t)n * c->dim)) return fail(VROD_EINVAL, "rows contain NaN or infinity");
        for (vrod_collection *part : c->parts) {
            vrod_status st = vrod_collection_insert_impl(part, rows, n, first_id, true);
            if (st != VROD_OK) return st;
        }
        sync_parent(c);
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
    VROD_CUDA(cudaSetDevice(ctx->device));
    if (ctx->world > 1 && !validated && !host_rows_finite(rows, (size_t)n * c->dim))
        return fail(VROD_EINVAL, "rows contain NaN or infinity");   // decided identically on every rank (same rows)
    if (c->count + n > c->capacity) {
        vrod_status gs = collection_grow(c, c->count + n);
        if (gs != VROD_OK) return gs;
    }
    const uint64_t loc0 = c->local, cnt = shard_rows_at(c->count + n, (uint32_t)ctx->rank, (uint32_t)ctx->world) - c->local;
    vrod_status st = for_my_pieces(c, c->count, n, [&](uint64_t lrow, uint64_t off, uint64_t m) -> vrod_status {
        const float *src = rows + (size_t)off * c->dim;
        float *dst = c->rows + (size_t)lrow * c->ld;
        if (c->ld != c->dim) VROD_CUDA(cudaMemsetAsync(dst, 0, (size_t)m * c->ld * sizeof(float), ctx->stream));
        VROD_CUDA(cudaMemcpy2DAsync(dst, (size_t)c->ld * sizeof(float), src, (size_t)c->dim * sizeof(float),
                                    (size_t)c->dim * sizeof(float), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        return VROD_OK;
    });
    if (st != VROD_OK) return st;
    st = finish_append(c, loc0, cnt, true);
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
    if (!c->parts.empty()) {
        for (vrod_collection *part : c->parts) {
            vrod_status st = vrod_collection_fill_synthetic_impl(part, n, seed);
            if (st != VROD_OK) return st;
        }
        sync_parent(c);
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
    VROD_CUDA(cudaSetDevice(ctx->device));
    const uint64_t loc0 = c->local, cnt = shard_rows_at(c->count + n, (uint32_t)ctx->rank, (uint32_t)ctx->world) - c->local;
    if (cnt) {
        VROD_CUDA(launch_fill_synthetic(c->rows, (uint32_t)loc0, (uint32_t)cnt, c->dim, c->ld, (uint32_t)ctx->rank, (uint32_t)ctx->world, seed,
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
    if (!c->parts.empty()) {   // global row indices: every part returns its blocks of the range
        for (vrod_collection *part : c->parts) {
            vrod_status st = for_my_pieces(part, row0, n, [&](uint64_t lrow, uint64_t off, uint64_t m) -> vrod_status {
                return vrod_collection_read_rows_impl(part, lrow, m, out + (size_t)off * c->dim);
            });
            if (st != VROD_OK) return st;
        }
        VROD_CUDA(cudaSetDevice(ctx->device));
        return VROD_OK;
    }
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
// Collective over the ranks of a process-per-GPU context: *all_ok is true on every rank iff `ok` was true on every rank
// (and nobody gets past it before everybody has arrived: it doubles as a barrier).  One byte per rank through NCCL.
static vrod_status ranks_agree(vrod_ctx *ctx, bool ok, bool *all_ok) {
    *all_ok = ok;
    if (ctx->world <= 1 || !ctx->comm) return VROD_OK;
    const int W = ctx->world;
    VROD_CUDA(cudaSetDevice(ctx->device));
    VROD_CUDA(ctx->agree.ensure((size_t)W + 16));
    unsigned char *d = reinterpret_cast<unsigned char *>(ctx->agree.p);
    const unsigned char mine = ok ? 1 : 0;
    std::vector<unsigned char> all(W);
    VROD_CUDA(cudaMemcpyAsync(d + W, &mine, 1, cudaMemcpyHostToDevice, ctx->stream));
    VROD_NCCL(g_nccl.AllGather(d + W, d, 1, ncclChar, ctx->comm, ctx->stream));
    VROD_CUDA(cudaMemcpyAsync(all.data(), d, W, cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    for (unsigned char v : all) *all_ok = *all_ok && v;
    return VROD_OK;
}

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
    if (ctx->world > 1) {
        // process-per-GPU context (collective): rank 0 lays the file out, then every rank writes the blocks it holds at their
        // offsets -- each row crosses one PCIe link once and nobody holds more than a block on the host
        if (ctx->parent) return fail(VROD_EINVAL, "vrod_collection_save: use the multi-GPU context's own collection handle");
        VROD_CUDA(cudaSetDevice(ctx->device));
        bool ok = true, all_ok = true;
        if (ctx->rank == 0) {
            FILE *f0 = fopen(path, "wb");
            ColFileHeader h{};
            memcpy(h.magic, "VRODCOL1", 8);
            h.dim = c->dim;
            h.metric = (uint32_t)c->metric;
            h.count = c->count;
            ok = f0 && fwrite(&h, sizeof(h), 1, f0) == 1;
            if (f0) ok = (fclose(f0) == 0) && ok;
        }
        vrod_status st = ranks_agree(ctx, ok, &all_ok);
        if (st != VROD_OK) return st;
        if (!all_ok) return fail(VROD_EINVAL, std::string("cannot create '") + path + "'");
        FILE *f = fopen(path, "r+b");
        ok = f != nullptr;
        std::vector<float> buf((size_t)kShardBlock * c->dim);
        if (ok) {
            st = for_my_pieces(c, 0, c->count, [&](uint64_t lrow, uint64_t off, uint64_t m) -> vrod_status {
                vrod_status rs = vrod_collection_read_rows(c, lrow, m, buf.data());
                if (rs != VROD_OK) return rs;
                if (fseeko(f, (off_t)(sizeof(ColFileHeader) + off * c->dim * sizeof(float)), SEEK_SET) != 0 ||
                    fwrite(buf.data(), sizeof(float) * c->dim, m, f) != m)
                    ok = false;
                return VROD_OK;
            });
            ok = (fclose(f) == 0) && ok && st == VROD_OK;
        }
        st = ranks_agree(ctx, ok, &all_ok);
        if (st != VROD_OK) return st;
        return all_ok ? VROD_OK : fail(VROD_EINVAL, std::string("short write to '") + path + "'");
    }
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
    if (ctx->world > 1 && !ctx->parent && ctx->subs.empty()) {
        // process-per-GPU context (collective): every rank reads only the blocks that are dealt to it
        bool ok = true, all_ok = true;
        std::string err;
        std::vector<float> blk((size_t)kShardBlock * h.dim);
        const uint64_t mine = shard_rows_at(h.count, (uint32_t)ctx->rank, (uint32_t)ctx->world);
        st = for_my_pieces(c, 0, h.count, [&](uint64_t lrow, uint64_t off, uint64_t m) -> vrod_status {
            if (!ok) return VROD_OK;
            if (fseeko(f, (off_t)(sizeof(ColFileHeader) + off * h.dim * sizeof(float)), SEEK_SET) != 0 ||
                fread(blk.data(), sizeof(float) * h.dim, m, f) != m) {
                ok = false;
                err = std::string("'") + path + "' is truncated";
                return VROD_OK;
            }
            float *dst = c->rows + (size_t)lrow * c->ld;
            if (c->ld != c->dim) VROD_CUDA(cudaMemsetAsync(dst, 0, (size_t)m * c->ld * sizeof(float), ctx->stream));
            VROD_CUDA(cudaMemcpy2DAsync(dst, (size_t)c->ld * sizeof(float), blk.data(), (size_t)c->dim * sizeof(float),
                                        (size_t)c->dim * sizeof(float), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
            VROD_CUDA(cudaStreamSynchronize(ctx->stream));   // (the staging block is reused)
            return VROD_OK;
        });
        fclose(f);
        if (st == VROD_OK && ok) {
            st = finish_append(c, 0, mine, true);
            if (st != VROD_OK) {
                ok = false;
                err = g_last_error;
                st = VROD_OK;
            }
        }
        if (st != VROD_OK) ok = false;
        vrod_status as = ranks_agree(ctx, ok, &all_ok);
        if (as != VROD_OK || !all_ok) {
            vrod_collection_drop(ctx, name);
            if (as != VROD_OK) return as;
            return fail(VROD_EINVAL, !err.empty() ? err : std::string("loading '") + path + "' failed on another rank");
        }
        c->count = h.count;
        c->local = mine;
        if (out) *out = c;
        return VROD_OK;
    }
    // single GPU, or the single-process multi-GPU context: the file is read once, in chunks, and appended
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
    s.mirror_stats = reinterpret_cast<unsigned int *>(c->flags) + 2;
    s.n = (uint32_t)c->local;
    s.dim = c->dim;
    s.ld = c->ld;
    s.metric = (int)c->metric;
    s.rank = (uint32_t)c->ctx->rank;
    s.world = (uint32_t)c->ctx->world;
    return s;
}

// -------------------------------------------------------------------------------------------------
// One search = local pass (this GPU's shard) [+ read-back of the batched pass's guard flags and rescans] [+ exchange].
// The stages are separate functions so that the single-process multi-GPU context can run each stage over all its
// devices before the next one (one host thread, nothing waits on a device that has not been given its work yet).
// -------------------------------------------------------------------------------------------------
struct LocalPass {
    bool batched = false;     // the tensor-core pass ran: its guard flags must be read back (batched_fetch_status + batched_rescans)
    bool used_scan = false;   // per-query f32 scans ran (host_checks: the caller rescans the flagged queries itself)
    bool exchanged = false;   // the scan kernels exchanged and merged the ranks' lists themselves (fused): d_ids / d_dist hold
                              // the GLOBAL answer and d_status the global guard flags
};

// Fused exchange inside the scan kernels' last CTA: which windows, which sequence numbers.  x.windows == nullptr: none.
struct FusePlan {
    XchgArgs x{};             // rank / world / root / err filled in; seq and qi are set per launch
    uint32_t seq_fast = 0, seq_exact = 0;
    bool on = false;
};

// May the scan kernels of this call exchange the lists themselves?  (single-query scans, small enough for the windows and
// for the scan kernels' shared-memory merge buffer)
static bool can_fuse_scan(const vrod_ctx *ctx, const ScanPlan &fp, const ScanPlan &xp, uint32_t b, uint32_t k) {
    if (ctx->world <= 1 || !ctx->fused_exchange) return false;
    if (b > kXchgMaxB || (size_t)b * k > kXchgMaxHits) return false;
    const uint32_t need = (uint32_t)ctx->world * k;
    return need <= (uint32_t)fp.cap && need <= (uint32_t)xp.cap;
}

static ScanScratch scan_scratch(vrod_ctx *ctx, int *d_status) {
    return ScanScratch{reinterpret_cast<unsigned long long *>(ctx->blk_cand.p), reinterpret_cast<unsigned int *>(ctx_ticket(ctx)),
                       d_status, ctx->dev_counters};
}

// The batched (tensor-core) pass of this rank's shard: mirror upkeep, the phased tile / finish launches.  The guard flags
// land in d_status; the caller reads them back (batched_fetch_status) and re-answers the flagged queries (batched_rescans).
static vrod_status enqueue_batched(vrod_collection *c, const float *d_q, uint32_t b, uint32_t k, uint64_t *d_ids, float *d_dist,
                                   int *d_status) {
    vrod_ctx *ctx = c->ctx;
    const ShardView s = shard_view(c);
    Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
    const bool direct = ctx->world == 1;
    ShardView sb = s;
    if (c->path != 4 && !c->mirror_failed) {
        // bf16 operand mirror: allocate once for the shard's capacity, convert the rows appended since the last time
        if (!c->rows_h) {
            // VROD_NO_MIRROR=1 behaves like a failed allocation (tests of the fallback; memory-tight deployments)
            static const bool no_mirror = getenv("VROD_NO_MIRROR") != nullptr;
            const cudaError_t me = no_mirror ? cudaErrorMemoryAllocation : cudaMalloc(&c->rows_h, mirror_bytes(c->shard_rows, c->dim));
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
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool sample = false;
    if (ctx->profiling) {
        e0 = ctx->prof_event();
        e1 = ctx->prof_event();
    } else if (ctx->cost.want_sample()) {
        sample = ctx->cost.events(&e0, &e1);
    }
    cudaError_t e = launch_batched_search(sb, d_q, b, k, ctx->sms, &ctx->batched.p, &ctx->batched.bytes, d_status, local,
                                          direct ? reinterpret_cast<unsigned long long *>(d_ids) : nullptr, direct ? d_dist : nullptr,
                                          ctx->stream, &bs, e0, e1, c->band_mode, !c->guess_off, c->wide_margin);
    if (e != cudaSuccess) return fail(VROD_ECUDA, std::string("batched search: ") + cudaGetErrorString(e));
    if (sample)
        ctx->cost.pending(2, (double)((b + 255) / 256) * ((double)s.n / 128.0) * ((double)mirror_ld(s.dim) / 144.0) / (double)ctx->sms);
    ctx->stats.kernel_launches += bs.launches;
    ctx->stats.batched_tiles += bs.tiles;
    return VROD_OK;
}

// Stage 1.  Queries already on the device as [b x ld]; enqueue the local pass, no synchronisation.  On a single-GPU
// context the kernels write d_ids / d_dist themselves; otherwise the local hits land in ctx->hits_local.
// d_status: per-query guard flags (device).  host_checks: the caller reads d_status after synchronising and rescans the
// flagged queries itself, so the conditional exact-scan launches are left out (single GPU only).
static vrod_status local_enqueue(vrod_collection *c, const float *d_q, uint32_t b, uint32_t k, uint64_t *d_ids, float *d_dist,
                                 bool force_exact, int *d_status, bool host_checks, LocalPass *lp, FusePlan *fuse = nullptr) {
    vrod_ctx *ctx = c->ctx;
    const ShardView s = shard_view(c);
    const size_t nhits = (size_t)b * k;
    VROD_CUDA(ctx->hits_local.ensure(nhits * sizeof(Hit)));
    Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
    const ScanScratch scr = scan_scratch(ctx, d_status);
    const bool exact_only = force_exact || c->path == 2 || !c->fast_ok;
    // which kind of pass?  (the choice comes first: the fused exchange is for the scan kernels only)
    // automatic choice between b single-query scans and one batched (tensor-core) pass: the context's cost model
    // (CostModel: seeded from round-1 measurements, then corrected by the times this context measures)
    // Sharded contexts: every rank (device) must reach the SAME decision -- the scan path exchanges inside the scan kernels,
    // conditional re-scans included, and a rank on the other path would not answer -- so only rank-invariant inputs
    // count there: the shard size by the partition rule, the model's seeds, no learned rescan share.
    bool prefer_batched = false;
    if (b >= 2) {
        const bool sharded = ctx->world > 1;
        const CostModel seeds;
        const CostModel &cm = sharded ? seeds : ctx->cost;
        const double rows = sharded ? (double)c->capacity / (double)ctx->world : (double)s.n;
        const double share = sharded ? 0.0 : c->rescan_share;
        const double t_scan = cm.scan_seconds(rows * s.ld * 4.0);
        const double groups = (double)((b + 255) / 256);
        const double t_batched = cm.batched_seconds(groups * (rows / 128.0) * ((double)mirror_ld(s.dim) / 144.0) / (double)ctx->sms);
        prefer_batched = (1.0 - share) * (double)b * t_scan > t_batched;
        if (!prefer_batched) c->rescan_share *= 0.995;   // forget slowly: the batched pass is probed again later
    }
    const bool batched = !exact_only && batched_supported(s, b, k) && (c->path == 3 || c->path == 4 || (c->path == 0 && prefer_batched));
    const bool fused = fuse && fuse->on && !batched;
    if (fuse && !fused) fuse->on = false;
    if (ctx->world > 1 && !fused) host_checks = false;   // without the exchange the flags are rank-local: every rank rescans on its own
    // single GPU, or fused exchange: the scan kernels write the final arrays themselves, no merge launch
    const bool direct = ctx->world == 1 || fused;
    auto xf = [&](uint32_t qi, uint32_t seq) {
        fuse->x.qi = qi;
        fuse->x.seq = seq;
        return &fuse->x;
    };
    auto oid = [&](uint32_t qi) { return direct ? reinterpret_cast<unsigned long long *>(d_ids) + (size_t)qi * k : nullptr; };
    auto odd = [&](uint32_t qi) { return direct ? d_dist + (size_t)qi * k : nullptr; };
    if (batched) {
        vrod_status bst = enqueue_batched(c, d_q, b, k, d_ids, d_dist, d_status);
        if (bst != VROD_OK) return bst;
        lp->batched = true;
    } else if (exact_only) {
        const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
        for (uint32_t qi = 0; qi < b; ++qi) {
            const float *q = d_q + (size_t)qi * s.ld;
            VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, nullptr, local + (size_t)qi * k, oid(qi), odd(qi), ctx->stream,
                                        fused ? xf(qi, fuse->seq_fast) : nullptr, d_status + qi));
            ctx->stats.kernel_launches++;
            if (fused && !host_checks) {
                // a peer whose rows allow the f32 pass may have flagged the query: its re-scan round needs this rank's list too
                VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi), ctx->stream,
                                            xf(qi, fuse->seq_exact), d_status + qi));
                ctx->stats.kernel_launches++;
            }
        }
        lp->used_scan = fused;    // (host_checks callers: the flags are meaningful, a peer may have raised one)
        lp->exchanged = fused;
    } else {
        const ScanPlan fp = make_scan_plan(s, k, ctx->sms, false);
        const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
        for (uint32_t qi = 0; qi < b; ++qi) {
            const float *q = d_q + (size_t)qi * s.ld;
            static const bool scan_dbg = kDbg && getenv("VROD_SCAN_DEBUG") != nullptr;
            unsigned long long *dbgbuf = nullptr;
            if (scan_dbg) {
                dbgbuf = scan_debug_enable();
                const unsigned long long init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
                cudaMemcpyAsync(dbgbuf, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
            }
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            bool sample = false;
            if (ctx->profiling) {
                e0 = ctx->prof_event();
                e1 = ctx->prof_event();
            } else if (qi == 0 && ctx->cost.want_sample()) {
                sample = ctx->cost.events(&e0, &e1);
            }
            if (e0) VROD_CUDA(cudaEventRecord(e0, ctx->stream));
            VROD_CUDA(launch_fast_scan(s, q, k, fp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi),
                                       ctx->stream, fused ? xf(qi, fuse->seq_fast) : nullptr));
            if (scan_dbg) {
                unsigned long long h[8];
                cudaStreamSynchronize(ctx->stream);
                cudaMemcpy(h, dbgbuf, sizeof(h), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[scan dbg] scan loops %.1f us | wait+ticket %.1f | merge %.1f (heads+T0 %.1f, lists %.1f, sort %.1f) | rerank %.1f | sort+write %.1f | total %.1f us\n",
                        (h[1] - h[0]) / 1e3, (h[2] - h[1]) / 1e3, (h[3] - h[2]) / 1e3, (h[6] - h[2]) / 1e3, (h[7] - h[6]) / 1e3, (h[3] - h[7]) / 1e3,
                        (h[4] - h[3]) / 1e3, (h[5] - h[4]) / 1e3, (h[5] - h[0]) / 1e3);
            }
            if (e1) VROD_CUDA(cudaEventRecord(e1, ctx->stream));
            if (sample) ctx->cost.pending(1, (double)s.n * s.ld * 4.0);
            if (!host_checks) {
                // (fused: the flag is the OR over all ranks, so every rank runs this re-scan or none does, and the re-scan
                // exchanges again under the call's second sequence number)
                VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, d_status + qi, local + (size_t)qi * k, oid(qi), odd(qi),
                                            ctx->stream, fused ? xf(qi, fuse->seq_exact) : nullptr, d_status + qi));
                ctx->stats.kernel_launches++;
            }
            ctx->stats.kernel_launches++;
        }
        ctx->stats.fast_scans += b;
        lp->used_scan = true;
        lp->exchanged = fused;
    }
    return VROD_OK;
}

// Stage 2a (after a batched local pass).  The number of queries whose guard failed is only known on the device:
// copy the flags to pinned memory and mark the point in the stream.
static vrod_status batched_fetch_status(vrod_collection *c, uint32_t b, const int *d_status) {
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(ctx->status_host.ensure(sizeof(int) * b));
    if (!ctx->ev_status) VROD_CUDA(cudaEventCreateWithFlags(&ctx->ev_status, cudaEventDisableTiming));
    VROD_CUDA(cudaMemcpyAsync(ctx->status_host.p, d_status, sizeof(int) * b, cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaEventRecord(ctx->ev_status, ctx->stream));
    return VROD_OK;
}

// Stage 2b.  Wait for the flags (the one synchronisation of the batched path; the rescans are shard-local, no
// collective is involved) and answer the flagged queries with the single-query scan.
static vrod_status batched_rescans(vrod_collection *c, const float *d_q, uint32_t b, uint32_t k, uint64_t *d_ids, float *d_dist,
                                   int *d_status) {
    vrod_ctx *ctx = c->ctx;
    VROD_CUDA(cudaEventSynchronize(ctx->ev_status));
    const int *hs = reinterpret_cast<const int *>(ctx->status_host.p);
    const ShardView s = shard_view(c);
    const bool direct = ctx->world == 1;
    Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
    const ScanScratch scr = scan_scratch(ctx, d_status);
    uint32_t flagged = 0, misguessed = 0;
    for (uint32_t qi = 0; qi < b; ++qi) {
        flagged += hs[qi] ? 1u : 0u;
        misguessed += hs[qi] == 2 ? 1u : 0u;
    }
    static const bool verbose = getenv("VROD_VERBOSE") != nullptr;
    if (verbose && flagged)
        fprintf(stderr, "[vrod] batched pass of %u queries on '%s': %u flagged (%u by a guessed threshold); wide margin %d, band mode %d, guessing %d\n", b,
                c->name.c_str(), flagged, misguessed, (int)c->wide_margin, (int)c->band_mode, (int)!c->guess_off);
    if (misguessed > 2u + b / 64u && !c->guess_off) {
        // Guessed thresholds failed for more than a stray query: the rows of this collection do not arrive in an order in
        // which the rows seen so far predict the rows to come.  Its batched passes filter at the k'-th best key from now
        // on (more candidates, more phases, no assumption) -- starting with this batch.
        c->guess_off = true;
        vrod_status st = enqueue_batched(c, d_q, b, k, d_ids, d_dist, d_status);
        if (st == VROD_OK) st = batched_fetch_status(c, b, d_status);
        if (st != VROD_OK) return st;
        VROD_CUDA(cudaEventSynchronize(ctx->ev_status));
        flagged = 0;
        for (uint32_t qi = 0; qi < b; ++qi) flagged += hs[qi] ? 1u : 0u;
    }
    const bool ran_tight = !c->wide_margin && !c->band_mode;   // (what the pass just read back ran with)
    if (flagged > misguessed && ran_tight) {
        // A proof failed at the tight margin (k' = 1.5k + 16): the k-th and the k'-th neighbour of that query are closer than
        // the surrogate's error.  One rescan costs a whole scan -- as much as the batched pass itself on a 1M-row
        // collection -- while the wide margin costs a few per cent per batch, so this collection keeps the wide margin
        // from now on; the batch at hand is answered again right away if more than 5 % of it failed.
        c->wide_margin = true;
    }
    if (flagged * 20u > b && ran_tight) {
        vrod_status st = enqueue_batched(c, d_q, b, k, d_ids, d_dist, d_status);
        if (st == VROD_OK) st = batched_fetch_status(c, b, d_status);
        if (st != VROD_OK) return st;
        VROD_CUDA(cudaEventSynchronize(ctx->ev_status));
        flagged = 0;
        for (uint32_t qi = 0; qi < b; ++qi) flagged += hs[qi] ? 1u : 0u;
    }
    if (flagged * 20u > b && !c->band_mode && c->rows_h && c->path != 4) {
        // more than 5 % of the fixed-k' proofs failed: this collection holds neighbours the bf16 contraction cannot tell
        // apart.  From now on its batched passes keep every candidate inside the error band (up to 1024 per query) -- and
        // this batch is answered that way at once: one more tensor-core pass instead of `flagged` single-query scans.
        c->band_mode = true;
        vrod_status st = enqueue_batched(c, d_q, b, k, d_ids, d_dist, d_status);
        if (st == VROD_OK) st = batched_fetch_status(c, b, d_status);
        if (st != VROD_OK) return st;
        VROD_CUDA(cudaEventSynchronize(ctx->ev_status));
        flagged = 0;
        for (uint32_t qi = 0; qi < b; ++qi) flagged += hs[qi] ? 1u : 0u;
    }
    {   // rises at once, falls slowly
        const double share = (double)flagged / (double)b, ema = 0.5 * c->rescan_share + 0.5 * share;
        c->rescan_share = share > ema ? share : ema;
    }
    if (!flagged) return VROD_OK;
    const ScanPlan fp = make_scan_plan(s, k, ctx->sms, false);
    const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
    for (uint32_t qi = 0; qi < b; ++qi) {
        if (!hs[qi]) continue;
        const float *q = d_q + (size_t)qi * s.ld;
        unsigned long long *oi = direct ? reinterpret_cast<unsigned long long *>(d_ids) + (size_t)qi * k : nullptr;
        float *od = direct ? d_dist + (size_t)qi * k : nullptr;
        VROD_CUDA(launch_fast_scan(s, q, k, fp, scr, d_status + qi, local + (size_t)qi * k, oi, od, ctx->stream));
        VROD_CUDA(launch_exact_scan(s, q, k, xp, scr, d_status + qi, local + (size_t)qi * k, oi, od, ctx->stream));
        ctx->stats.kernel_launches += 2;
        ctx->stats.fast_scans++;
    }
    return VROD_OK;
}

// Stage 3 (one process per GPU).  Every rank ends up with the global answer in d_ids / d_dist.
static vrod_status exchange_enqueue(vrod_collection *c, uint32_t b, uint32_t k, uint64_t *d_ids, float *d_dist, int *d_xerr,
                                    bool *used_fused) {
    vrod_ctx *ctx = c->ctx;
    const size_t nhits = (size_t)b * k;
    Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
    if (used_fused) *used_fused = false;
    if (ctx->world == 1) return VROD_OK;   // the kernels of the local pass wrote the final arrays
    if (ctx->fused_exchange && b <= kXchgMaxB && nhits <= kXchgMaxHits) {
        // one kernel: push the local lists into every rank's window over NVLink, wait for the peers', merge
        XchgArgs x{ctx->d_windows, (uint32_t)ctx->rank, (uint32_t)ctx->world, ++ctx->xchg_seq, kXchgAllRanks, 0u,
                   d_xerr ? d_xerr : ctx->d_xchg_err};
        VROD_CUDA(launch_exchange_merge(x, local, b, k, reinterpret_cast<unsigned long long *>(d_ids), d_dist, ctx->stream));
        ctx->stats.kernel_launches++;
        if (used_fused) *used_fused = true;
        return VROD_OK;
    }
    if (!ctx->comm) return fail(VROD_ENCCL, "no communicator on this context");
    VROD_CUDA(ctx->hits_all.ensure(nhits * sizeof(Hit) * ctx->world));
    VROD_NCCL(g_nccl.AllGather(local, ctx->hits_all.p, nhits * sizeof(Hit), ncclChar, ctx->comm, ctx->stream));
    VROD_CUDA(launch_merge_hits(reinterpret_cast<const Hit *>(ctx->hits_all.p), (uint32_t)ctx->world, b, k,
                                reinterpret_cast<unsigned long long *>(d_ids), d_dist, ctx->stream));
    ctx->stats.kernel_launches++;
    return VROD_OK;
}

// All stages on one (single-GPU or per-rank) context.
static vrod_status search_enqueue(vrod_collection *c, const float *d_q, uint32_t b, uint32_t k, uint64_t *d_ids,
                                  float *d_dist, bool force_exact, int *d_status = nullptr, bool host_checks = false,
                                  bool *used_scan = nullptr, int *d_xerr = nullptr, bool *used_fused = nullptr,
                                  bool *scan_fused = nullptr) {
    vrod_ctx *ctx = c->ctx;
    if (!d_status) d_status = ctx_status(ctx);
    ctx->cost.collect();
    LocalPass lp;
    FusePlan fuse;
    if (ctx->world > 1) {
        const ShardView s = shard_view(c);
        fuse.on = can_fuse_scan(ctx, make_scan_plan(s, k, ctx->sms, false), make_scan_plan(s, k, ctx->sms, true), b, k);
        fuse.x = XchgArgs{ctx->d_windows, (uint32_t)ctx->rank, (uint32_t)ctx->world, 0u, kXchgAllRanks, 0u, d_xerr ? d_xerr : ctx->d_xchg_err};
        // two sequence numbers per call, used or not (all ranks count alike): the first pass, the conditional exact re-scans
        fuse.seq_fast = ctx->xchg_seq + 1;
        fuse.seq_exact = ctx->xchg_seq + 2;
    }
    vrod_status st = local_enqueue(c, d_q, b, k, d_ids, d_dist, force_exact, d_status, host_checks, &lp, &fuse);
    if (st != VROD_OK) return st;
    if (lp.batched) {
        st = batched_fetch_status(c, b, d_status);
        if (st == VROD_OK) st = batched_rescans(c, d_q, b, k, d_ids, d_dist, d_status);
        if (st != VROD_OK) return st;
    }
    if (used_scan) *used_scan = lp.used_scan;
    if (scan_fused) *scan_fused = lp.exchanged;
    if (lp.exchanged) {
        ctx->xchg_seq += 2;
        if (used_fused) *used_fused = true;
    } else {
        st = exchange_enqueue(c, b, k, d_ids, d_dist, d_xerr, used_fused);
        if (st != VROD_OK) return st;
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
    if (!c->parts.empty())
        return fail(VROD_EINVAL, "the device-pointer search needs a single-device context (a multi-GPU context answers host buffers)");
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

// validate + pack (zero padded to ld) into pinned staging.  Queries outside the range the f32 / tensor-core passes'
// error bounds cover are marked in unsafe_q: they are answered by the exact f64 scan.
static vrod_status pack_queries(const vrod_collection *c, const float *queries, uint32_t b, float *qh,
                                std::vector<unsigned char> &unsafe_q, uint32_t *n_unsafe) {
    unsafe_q.assign(b, 0);
    *n_unsafe = 0;
    for (uint32_t i = 0; i < b; ++i) {
        const float *src = queries + (size_t)i * c->dim;
        float *dst = qh + (size_t)i * c->ld;
        // four independent sums and branch-free checks: one serial f64 chain with an early exit per element was ~150 us of
        // host time for a 1024 x 128 batch (a third of the host path's overhead at configs[2])
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        float big = 0.f;            // max |v| (a NaN never wins a fmaxf: non-finite values are caught by their exponent)
        uint32_t nonfinite = 0;
        uint32_t j = 0;
        for (; j + 4 <= c->dim; j += 4) {
            for (int e = 0; e < 4; ++e) {
                const float v = src[j + e];
                uint32_t bits;
                memcpy(&bits, &v, sizeof bits);
                nonfinite |= ((bits & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
                big = fmaxf(big, fabsf(v));
                acc[e] += (double)v * (double)v;
                dst[j + e] = v;
            }
        }
        for (; j < c->dim; ++j) {
            const float v = src[j];
            nonfinite |= isfinite(v) ? 0u : 1u;
            big = fmaxf(big, fabsf(v));
            acc[0] += (double)v * (double)v;
            dst[j] = v;
        }
        if (nonfinite) return fail(VROD_EINVAL, "query contains NaN or infinity");
        const double nq = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        bool u = big > 0x1p40f;
        for (uint32_t j = c->dim; j < c->ld; ++j) dst[j] = 0.f;
        if (nq > 0.0 && (nq < 0x1p-80 || nq > 0x1p100)) u = true;
        if (c->metric == VROD_COSINE && nq == 0.0) u = true;  // all distances are 1: answered by the exact scan
        unsafe_q[i] = u ? 1 : 0;
        *n_unsafe += u ? 1 : 0;
    }
    return VROD_OK;
}

// layout of the packed result buffer [ids | dist | status (+1: peer-exchange error flag)]: results and guard flags
// come back in ONE device-to-host copy
struct PackLayout {
    size_t nres, off_dist, off_stat, bytes;
    PackLayout(uint32_t b, uint32_t k) {
        nres = (size_t)b * k;
        off_dist = nres * sizeof(uint64_t);
        off_stat = off_dist + ((nres * sizeof(float) + 15) & ~(size_t)15);
        bytes = off_stat + ((size_t)b + 1) * sizeof(int);
    }
};

// SEARCH on a single-process multi-GPU context: every stage runs over all devices before the next one starts, from
// the caller's one thread.  Device 0 receives the per-device top-k lists (fused NVLink push + merge kernel, or peer
// copies + merge for large batches) and is the only device the host reads.
static vrod_status multi_search(vrod_collection *pc, const float *queries, uint32_t b, uint32_t k, uint64_t *out_ids, float *out_dist) {
    vrod_ctx *P = pc->ctx;
    const int W = (int)P->subs.size();
    vrod_ctx *s0 = P->subs[0];
    const size_t qbytes = (size_t)b * pc->ld * sizeof(float);
    VROD_CUDA(cudaSetDevice(s0->device));
    VROD_CUDA(P->q_host.ensure(qbytes));
    float *qh = reinterpret_cast<float *>(P->q_host.p);
    std::vector<unsigned char> unsafe_q;
    uint32_t n_unsafe = 0;
    vrod_status st = pack_queries(pc, queries, b, qh, unsafe_q, &n_unsafe);
    if (st != VROD_OK) return st;
    const bool unsafe = n_unsafe != 0;   // the whole call goes the exact way
    const PackLayout L(b, k);
    VROD_CUDA(P->ids_host.ensure(L.bytes));
    unsigned char *hpack = reinterpret_cast<unsigned char *>(P->ids_host.p);
    const bool fused = s0->fused_exchange && b <= kXchgMaxB && L.nres <= kXchgMaxHits;
    std::vector<LocalPass> lp(W);
    std::vector<FusePlan> fuse(W);
    auto ids_of = [&](vrod_ctx *sub) { return reinterpret_cast<uint64_t *>(sub->out_ids.p); };
    auto dist_of = [&](vrod_ctx *sub) { return reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sub->out_ids.p) + L.off_dist); };
    auto stat_of = [&](vrod_ctx *sub) { return reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(sub->out_ids.p) + L.off_stat); };
    // single-query scans: the scan kernels push their lists to device 0 themselves (one kernel per device and query); two
    // sequence numbers per call as in the per-rank mode (first pass, exact re-scans)
    const uint32_t seq_fast = P->xchg_seq + 1, seq_exact = P->xchg_seq + 2;
    // stage 1: queries in, local pass on every device
    for (int r = 0; r < W; ++r) {
        vrod_ctx *sub = P->subs[r];
        VROD_CUDA(cudaSetDevice(sub->device));
        sub->cost.collect();
        VROD_CUDA(sub->q_dev.ensure(qbytes));
        VROD_CUDA(sub->out_ids.ensure(L.bytes));
        VROD_CUDA(cudaMemcpyAsync(sub->q_dev.p, qh, qbytes, cudaMemcpyHostToDevice, sub->stream));
        const ShardView sv = shard_view(pc->parts[r]);
        fuse[r].on = can_fuse_scan(sub, make_scan_plan(sv, k, sub->sms, false), make_scan_plan(sv, k, sub->sms, true), b, k);
        fuse[r].x = XchgArgs{sub->d_windows, (uint32_t)r, (uint32_t)W, 0u, 0u, 0u, stat_of(sub) + b};
        fuse[r].seq_fast = seq_fast;
        fuse[r].seq_exact = seq_exact;
        // host_checks: with the fused exchange device 0 ends up with the OR of all devices' guard flags, and the host (which
        // reads only device 0) starts the exact re-scans on every device
        st = local_enqueue(pc->parts[r], reinterpret_cast<const float *>(sub->q_dev.p), b, k, ids_of(sub), dist_of(sub), unsafe,
                           stat_of(sub), true, &lp[r], &fuse[r]);
        if (st != VROD_OK) return st;
        if (lp[r].exchanged != lp[0].exchanged || lp[r].batched != lp[0].batched)
            return fail(VROD_EINVAL, "internal error: the devices of a multi-GPU context chose different search paths");
        if (lp[r].batched) {
            st = batched_fetch_status(pc->parts[r], b, stat_of(sub));
            if (st != VROD_OK) return st;
        }
    }
    // stage 2: the batched passes' guard flags, one device after the other (the others keep computing)
    for (int r = 0; r < W; ++r) {
        if (!lp[r].batched) continue;
        vrod_ctx *sub = P->subs[r];
        VROD_CUDA(cudaSetDevice(sub->device));
        st = batched_rescans(pc->parts[r], reinterpret_cast<const float *>(sub->q_dev.p), b, k, ids_of(sub), dist_of(sub), stat_of(sub));
        if (st != VROD_OK) return st;
    }
    // stage 3: the lists meet on device 0 (unless the scan kernels have seen to that)
    const bool scan_fused = lp[0].exchanged;
    if (scan_fused) {
        P->xchg_seq += 2;
    } else if (fused) {
        const uint32_t seq = ++P->xchg_seq;
        for (int r = 0; r < W; ++r) {
            vrod_ctx *sub = P->subs[r];
            VROD_CUDA(cudaSetDevice(sub->device));
            XchgArgs x{sub->d_windows, (uint32_t)r, (uint32_t)W, seq, 0u, 0u, stat_of(sub) + b};
            VROD_CUDA(launch_exchange_merge(x, reinterpret_cast<const Hit *>(sub->hits_local.p), b, k,
                                            reinterpret_cast<unsigned long long *>(ids_of(sub)), dist_of(sub), sub->stream));
            sub->stats.kernel_launches++;
        }
    } else {
        VROD_CUDA(cudaSetDevice(s0->device));
        VROD_CUDA(s0->hits_all.ensure(L.nres * sizeof(Hit) * W));
        Hit *all = reinterpret_cast<Hit *>(s0->hits_all.p);
        for (int r = 0; r < W; ++r) {
            vrod_ctx *sub = P->subs[r];
            VROD_CUDA(cudaSetDevice(sub->device));
            if (r == 0) {
                VROD_CUDA(cudaMemcpyAsync(all, sub->hits_local.p, L.nres * sizeof(Hit), cudaMemcpyDeviceToDevice, sub->stream));
            } else {
                VROD_CUDA(cudaMemcpyPeerAsync(all + (size_t)r * L.nres, s0->device, sub->hits_local.p, sub->device, L.nres * sizeof(Hit), sub->stream));
                VROD_CUDA(cudaEventRecord(sub->ev_done, sub->stream));
            }
        }
        VROD_CUDA(cudaSetDevice(s0->device));
        for (int r = 1; r < W; ++r) VROD_CUDA(cudaStreamWaitEvent(s0->stream, P->subs[r]->ev_done, 0));
        VROD_CUDA(launch_merge_hits(all, (uint32_t)W, b, k, reinterpret_cast<unsigned long long *>(ids_of(s0)), dist_of(s0), s0->stream));
        s0->stats.kernel_launches++;
    }
    // stage 4: one copy back from device 0
    VROD_CUDA(cudaSetDevice(s0->device));
    VROD_CUDA(cudaMemcpyAsync(hpack, s0->out_ids.p, L.bytes, cudaMemcpyDeviceToHost, s0->stream));
    VROD_CUDA(cudaStreamSynchronize(s0->stream));
    if ((fused || scan_fused) && reinterpret_cast<const int *>(hpack + L.off_stat)[b] == 1)
        return fail(VROD_ENCCL, "peer exchange timed out: a device did not deliver its lists");
    if (scan_fused && !unsafe) {
        // device 0 holds the OR of all devices' guard flags: the flagged queries are re-scanned exactly on every device, the
        // lists meet on device 0 again
        const int *hs = reinterpret_cast<const int *>(hpack + L.off_stat);
        std::vector<uint32_t> flagged;
        for (uint32_t qi = 0; qi < b; ++qi)
            if (hs[qi]) flagged.push_back(qi);
        if (!flagged.empty()) {
            for (int r = 0; r < W; ++r) {
                vrod_ctx *sub = P->subs[r];
                VROD_CUDA(cudaSetDevice(sub->device));
                const ShardView sv = shard_view(pc->parts[r]);
                const ScanPlan xp = make_scan_plan(sv, k, sub->sms, true);
                const ScanScratch scr = scan_scratch(sub, stat_of(sub));
                XchgArgs x{sub->d_windows, (uint32_t)r, (uint32_t)W, seq_exact, 0u, 0u, stat_of(sub) + b};
                Hit *local = reinterpret_cast<Hit *>(sub->hits_local.p);
                for (uint32_t qi : flagged) {
                    x.qi = qi;
                    VROD_CUDA(launch_exact_scan(sv, reinterpret_cast<const float *>(sub->q_dev.p) + (size_t)qi * sv.ld, k, xp, scr, nullptr,
                                                local + (size_t)qi * k, reinterpret_cast<unsigned long long *>(ids_of(sub)) + (size_t)qi * k,
                                                dist_of(sub) + (size_t)qi * k, sub->stream, &x));
                    sub->stats.kernel_launches++;
                }
            }
            VROD_CUDA(cudaSetDevice(s0->device));
            VROD_CUDA(cudaMemcpyAsync(hpack, s0->out_ids.p, L.bytes, cudaMemcpyDeviceToHost, s0->stream));
            VROD_CUDA(cudaStreamSynchronize(s0->stream));
            if (reinterpret_cast<const int *>(hpack + L.off_stat)[b] == 1)
                return fail(VROD_ENCCL, "peer exchange timed out: a device did not deliver its lists");
        }
    }
    memcpy(out_ids, hpack, L.nres * sizeof(uint64_t));
    memcpy(out_dist, hpack + L.off_dist, L.nres * sizeof(float));
    P->stats.searches += b;
    P->stats.h2d_bytes += qbytes * W;
    P->stats.d2h_bytes += L.bytes;
    return VROD_OK;
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
    if (!c->parts.empty()) return multi_search(c, queries, b, k, out_ids, out_dist);
    VROD_CUDA(cudaSetDevice(ctx->device));
    const size_t qbytes = (size_t)b * c->ld * sizeof(float);
    VROD_CUDA(ctx->q_host.ensure(qbytes));
    float *qh = reinterpret_cast<float *>(ctx->q_host.p);
    // out-of-range queries: each on its own on a single GPU (the rest of the batch keeps its fast path), the whole
    // call on sharded contexts (every rank must enqueue the same collectives)
    std::vector<unsigned char> unsafe_q;
    uint32_t n_unsafe = 0;
    st = pack_queries(c, queries, b, qh, unsafe_q, &n_unsafe);
    if (st != VROD_OK) return st;
    const bool unsafe = n_unsafe != 0 && (ctx->world > 1 || n_unsafe == b);   // the whole call goes the exact way
    const PackLayout L(b, k);
    VROD_CUDA(ctx->q_dev.ensure(qbytes));
    VROD_CUDA(ctx->out_ids.ensure(L.bytes));
    VROD_CUDA(ctx->ids_host.ensure(L.bytes));
    unsigned char *dpack = reinterpret_cast<unsigned char *>(ctx->out_ids.p);
    unsigned char *hpack = reinterpret_cast<unsigned char *>(ctx->ids_host.p);
    uint64_t *d_ids = reinterpret_cast<uint64_t *>(dpack);
    float *d_dist = reinterpret_cast<float *>(dpack + L.off_dist);
    int *d_stat = reinterpret_cast<int *>(dpack + L.off_stat);
    VROD_CUDA(cudaMemcpyAsync(ctx->q_dev.p, qh, qbytes, cudaMemcpyHostToDevice, ctx->stream));
    bool used_scan = false, used_fused = false, scan_fused = false;
    st = search_enqueue(c, reinterpret_cast<const float *>(ctx->q_dev.p), b, k, d_ids, d_dist, unsafe, d_stat, true, &used_scan,
                        d_stat + b, &used_fused, &scan_fused);
    if (st != VROD_OK) return st;
    VROD_CUDA(cudaMemcpyAsync(hpack, dpack, L.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VROD_CUDA(cudaStreamSynchronize(ctx->stream));
    // (the flag word is written only by the fused exchange kernel: the all-gather path leaves stale bytes there)
    if (used_fused && reinterpret_cast<const int *>(hpack + L.off_stat)[b] == 1)
        return fail(VROD_ENCCL, "peer exchange timed out: a rank did not take part in this search");
    if ((ctx->world == 1 || scan_fused) && !unsafe && (used_scan || n_unsafe)) {
        // the scans ran without their device-side conditional rescans: answer the flagged queries exactly now,
        // and with them the queries that were out of range for the fast paths.  (Sharded context with the exchange
        // fused into the scans: the flags are the OR over all ranks, so every rank re-scans the same queries, and the
        // re-scans exchange again under the call's second sequence number.)
        int *hs = reinterpret_cast<int *>(hpack + L.off_stat);
        bool any = false;
        for (uint32_t qi = 0; qi < b; ++qi) {
            hs[qi] = (used_scan && hs[qi] != 0) || unsafe_q[qi] ? 1 : 0;
            any = any || hs[qi] != 0;
        }
        if (any) {
            const ShardView s = shard_view(c);
            const ScanPlan xp = make_scan_plan(s, k, ctx->sms, true);
            const ScanScratch scr = scan_scratch(ctx, d_stat);
            Hit *local = reinterpret_cast<Hit *>(ctx->hits_local.p);
            XchgArgs x{ctx->d_windows, (uint32_t)ctx->rank, (uint32_t)ctx->world, ctx->xchg_seq, kXchgAllRanks, 0u, d_stat + b};
            for (uint32_t qi = 0; qi < b; ++qi) {
                if (!hs[qi]) continue;
                x.qi = qi;
                VROD_CUDA(launch_exact_scan(s, reinterpret_cast<const float *>(ctx->q_dev.p) + (size_t)qi * s.ld, k, xp, scr, nullptr,
                                            local + (size_t)qi * k, reinterpret_cast<unsigned long long *>(d_ids) + (size_t)qi * k,
                                            d_dist + (size_t)qi * k, ctx->stream, scan_fused ? &x : nullptr));
                ctx->stats.kernel_launches++;
            }
            VROD_CUDA(cudaMemcpyAsync(hpack, dpack, L.bytes, cudaMemcpyDeviceToHost, ctx->stream));
            VROD_CUDA(cudaStreamSynchronize(ctx->stream));
            if (scan_fused && reinterpret_cast<const int *>(hpack + L.off_stat)[b] == 1)
                return fail(VROD_ENCCL, "peer exchange timed out: a rank did not take part in this search");
        }
    }
    memcpy(out_ids, hpack, L.nres * sizeof(uint64_t));
    memcpy(out_dist, hpack + L.off_dist, L.nres * sizeof(float));
    ctx->stats.h2d_bytes += qbytes;
    ctx->stats.d2h_bytes += L.bytes;
    return VROD_OK;
}

extern "C" const char *vrod_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *vrod_version(void) { return "vrod_knn_b200 0.2.0 (sm_100a)"; }
