/*
 * vrod_knn.h -- C ABI of the B200-native exact k-nearest-neighbour scan that fills vRod's SEARCH
 * hot path.  Plain C: opaque handles, pointers and sizes only; no C++ exception crosses it.
 *
 * Reference interfaces these entry points stand behind (paths relative to the vRod repo):
 *   - src/command/types.rs:5-7      trait Command { fn execute(&self); }          (call shape)
 *   - src/command/types.rs:114-119  SearchCommand::execute  -- EMPTY body; vrod_collection_search
 *                                   is what a maintainer would call from it
 *   - src/command/types.rs:14-19    CreateCollectionCommand::execute -> vrod_collection_create
 *   - src/command/types.rs:27-32    DropCollectionCommand::execute   -> vrod_collection_drop
 *   - src/command/types.rs:38-43    ListCollectionsCommand::execute  -> vrod_collection_list
 *   - src/command/types.rs:62-67, 75-80  Insert/BulkInsertCommand::execute -> vrod_collection_insert
 *   - src/database/mod.rs:6-10      struct Database (collections are a TODO) -> vrod_ctx owns them
 *   - src/utils/embeddings.rs:29-31 Vec<Vec<f32>> rows: element type f32, one dimension per set
 * The reference has no FFI of its own; INTEGRATION.md shows the Rust
 * `extern "C"` block and the command bodies that bind these symbols.
 *
 * Rules: the caller owns every host buffer; the library copies in/out and retains no pointer.
 * Handles are freed only by vrod_ctx_destroy / vrod_collection_drop.  One calling thread per
 * vrod_ctx (the reference's Rc<RefCell<Database>> is !Send + !Sync, types.rs:10).  Every function
 * returns a vrod_status; vrod_last_error() gives the thread-local message of the last failure.
 * There is no CPU fallback: without a CUDA device vrod_ctx_create fails with VROD_ENOGPU.
 *
 * Search semantics (DESIGN.md "Search semantics"; parity is unpinned by the reference):
 *   Euclidean  dist = (f32) sqrt( SUM_j (x_j - q_j)^2 )
 *   Cosine     dist = (f32) (1 - dot / (sqrt(nx) sqrt(nq))),  dist = 1 if nx == 0 or nq == 0
 *   sums in f64 with fma in the fixed 128-way interleaved order + adjacent-pair tree,
 *   results ordered by (f32 dist ascending, id ascending), ids = insertion index (u64) from 0,
 *   slots beyond min(k, N) padded with id = UINT64_MAX, dist = +inf.
 *   Non-finite inputs are rejected with VROD_EINVAL.
 */
#ifndef VROD_KNN_H
#define VROD_KNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VROD_API __attribute__((visibility("default")))
#else
#define VROD_API
#endif

typedef struct vrod_ctx vrod_ctx;               /* device, stream, NCCL communicator, collections */
typedef struct vrod_collection vrod_collection; /* this rank's row shard: rows, norms, scratch     */

typedef enum {
    VROD_OK = 0,
    VROD_EINVAL = 1,    /* bad argument, non-finite input, dimension mismatch, k out of range */
    VROD_ENOTFOUND = 2, /* no collection of that name */
    VROD_EEXISTS = 3,   /* collection name already taken */
    VROD_ENOMEM = 4,    /* host or device allocation failed, or capacity exceeded */
    VROD_ECUDA = 5,     /* CUDA runtime error */
    VROD_ENCCL = 6,     /* NCCL error or NCCL library not loadable */
    VROD_ENOGPU = 7     /* no usable CUDA device: there is no CPU path */
} vrod_status;

typedef enum { VROD_EUCLIDEAN = 0, VROD_COSINE = 1 } vrod_metric;

#define VROD_MAX_K 1024u          /* largest k a search accepts */
#define VROD_COMM_ID_BYTES 128u   /* size of the opaque communicator id (an ncclUniqueId) */
#define VROD_PAD_ID UINT64_MAX
#define VROD_SHARD_BLOCK 4096u    /* rows per block of the block-cyclic deal of a sharded collection */

/* Counters of one context, for benches and tests (monotonic since ctx creation). */
typedef struct {
    uint64_t searches;         /* queries answered */
    uint64_t kernel_launches;  /* CUDA kernels this library launched */
    uint64_t fast_scans;       /* queries answered by the f32 scan + exact rerank */
    uint64_t exact_rescans;    /* queries whose guard failed and were re-answered by the f64 scan */
    uint64_t batched_tiles;    /* tensor-core tiles issued by the batched path */
    uint64_t h2d_bytes;        /* bytes copied host->device by search calls */
    uint64_t d2h_bytes;        /* bytes copied device->host by search calls */
} vrod_stats;

/* ---- context -------------------------------------------------------------------------- */

/* One GPU, no sharding.  `device` is a CUDA ordinal. */
VROD_API vrod_status vrod_ctx_create(int device, vrod_ctx **out);

/* Fill `out` (VROD_COMM_ID_BYTES) with a fresh communicator id; rank 0 calls it and ships the
 * bytes to the other ranks by any means (the benches use torch.distributed). */
VROD_API vrod_status vrod_comm_unique_id(void *out);

/* One process per GPU: this rank holds shard `rank` of `world` of every collection.
 * Rows are dealt to the shards BLOCK-CYCLICALLY: ids in blocks of VROD_SHARD_BLOCK consecutive rows, block j to
 * shard j mod world, a shard's blocks stored back to back -- every shard holds the same number of rows (within one
 * block) at any fill level, and a collection grows without moving a row between shards.  Collective: all ranks call
 * it with the same id.  A search scans every shard and merges the per-rank top-k lists under (dist, id) -- inside the
 * scan kernels' last CTA over NVLink peer memory for single queries, a fused push + merge kernel after the batched
 * pass, ncclAllGather + merge for calls too large for the exchange windows. */
VROD_API vrod_status vrod_ctx_create_sharded(int device, int rank, int world, const void *comm_id,
                                             vrod_ctx **out);

/* ONE process, n_devices GPUs: the shape the reference's caller needs -- a single-threaded
 * process holding Rc<RefCell<Database>> (src/command/types.rs:10; fn main, src/main.rs:42).  Every collection of
 * the context is row-sharded over the devices (device i of the list holds shard i of the block-cyclic deal, see
 * vrod_ctx_create_sharded); the calling thread drives all devices (one stream per device), the per-device top-k lists meet on the first device
 * through direct NVLink peer access (fused push + merge kernel; peer copies for large batches or when the devices
 * have no peer access) and only that device is read by the host.  All collection calls work on such a context
 * with GLOBAL meaning (read_rows takes global row indices, shard reports base 0 and every row); the device-pointer
 * search does not (VROD_EINVAL).  n_devices == 1 is vrod_ctx_create(device_ids[0]). */
VROD_API vrod_status vrod_ctx_create_multi(const int *device_ids, int n_devices, vrod_ctx **out);
/* Number of GPUs behind the context (1 unless it came from vrod_ctx_create_multi). */
VROD_API int vrod_ctx_devices(vrod_ctx *ctx);

VROD_API void vrod_ctx_destroy(vrod_ctx *ctx);
VROD_API vrod_status vrod_ctx_synchronize(vrod_ctx *ctx);
/* The cudaStream_t every kernel and copy of this context is issued on (for event timing). */
VROD_API void *vrod_ctx_stream(vrod_ctx *ctx);
VROD_API vrod_status vrod_ctx_stats(vrod_ctx *ctx, vrod_stats *out);
/* Kernel timing for benches: while enabled, every scan kernel launch (the f32 scan of the single-query
 * path, the tile kernel of the batched path) is bracketed by CUDA events on the context's stream.
 * vrod_ctx_profile_read synchronises, returns the summed device time (ms) and the number of bracketed
 * launches since the last read, and clears both. */
VROD_API vrod_status vrod_ctx_profile(vrod_ctx *ctx, int enable);
VROD_API vrod_status vrod_ctx_profile_read(vrod_ctx *ctx, double *kernel_ms, uint64_t *launches);
VROD_API int vrod_ctx_rank(vrod_ctx *ctx);
VROD_API int vrod_ctx_world(vrod_ctx *ctx);

/* ---- collections (Database surface) ---------------------------------------------------- */

/* CREATE: `capacity_rows` is the GLOBAL row capacity (an initial size: collections grow); a sharded context keeps
 * about capacity/world rows per rank.  dim >= 1.  Names are 1..200 characters of [A-Za-z0-9_.-], not "." or ".." (they become file
 * names and whitespace-delimited config tokens in the Database layer). */
VROD_API vrod_status vrod_collection_create(vrod_ctx *ctx, const char *name, uint32_t dim, vrod_metric metric,
                                            uint64_t capacity_rows, vrod_collection **out);
VROD_API vrod_status vrod_collection_get(vrod_ctx *ctx, const char *name, vrod_collection **out);
VROD_API vrod_status vrod_collection_drop(vrod_ctx *ctx, const char *name);
/* LISTCOLLECTIONS: writes the names, '\n'-separated and NUL-terminated, into buf (cap bytes);
 * *needed receives the size required including the NUL. */
VROD_API vrod_status vrod_collection_list(vrod_ctx *ctx, char *buf, size_t cap, size_t *needed);

VROD_API vrod_status vrod_collection_info(vrod_collection *c, uint32_t *dim, vrod_metric *metric,
                                          uint64_t *count, uint64_t *capacity);

/* INSERT / BULKINSERT: append n rows (row-major n x dim f32).  Ids are insertion indices; the id of
 * the first appended row is written to *first_id (may be NULL).  In a sharded context every rank
 * passes the same rows and keeps the blocks that are dealt to it; a batch with a NaN / infinity is rejected by every
 * rank alike.  A collection that is full grows: the capacity at least doubles, every shard re-allocates on its own
 * (device-to-device move), no row changes its shard. */
VROD_API vrod_status vrod_collection_insert(vrod_collection *c, const float *rows, uint64_t n, uint64_t *first_id);

/* Persistence (the step after INSERT; the reference's Database::load is a todo!(), src/database/mod.rs:19-21).
 * vrod_collection_save writes this collection to `path`: a 64-byte header ("VRODCOL1", dim, metric, count)
 * followed by count x dim f32 rows in id order, row-major, unpadded, little-endian -- the same file whatever the
 * context.  vrod_collection_load creates collection `name` from such a file; capacity_rows = 0 means "as many as
 * the file holds".  In a process-per-GPU context both are collective (same path on every rank, a shared file
 * system): rank 0 lays the file out and every rank writes / reads only the blocks it holds. */
VROD_API vrod_status vrod_collection_save(vrod_collection *c, const char *path);
VROD_API vrod_status vrod_collection_load(vrod_ctx *ctx, const char *name, const char *path, uint64_t capacity_rows,
                                          vrod_collection **out);

/* Append n synthetic rows generated ON THE DEVICE: element (i, j) of a collection is
 * u2f(philox4x32_10(counter = (i*dim + j) >> 2, key = seed)[(i*dim + j) & 3]), uniform [-1, 1);
 * i is the global row index.  The CPU oracle replays it bit for bit. */
VROD_API vrod_status vrod_collection_fill_synthetic(vrod_collection *c, uint64_t n, uint64_t seed);

/* Copy rows [row0, row0+n) of THIS RANK's shard (local indices; a single-GPU or multi-GPU context: global row
 * indices) back to the host (n x dim). */
VROD_API vrod_status vrod_collection_read_rows(vrod_collection *c, uint64_t row0, uint64_t n, float *out);
/* Global id of this rank's local row 0 (rank * VROD_SHARD_BLOCK) and the number of rows it holds; local row l has
 * id ((l / VROD_SHARD_BLOCK) * world + rank) * VROD_SHARD_BLOCK + l % VROD_SHARD_BLOCK. */
VROD_API vrod_status vrod_collection_shard(vrod_collection *c, uint64_t *id_base, uint64_t *local_rows);

/* ---- SEARCH ----------------------------------------------------------------------------- */

/* Exact top-k of b queries (row-major b x dim f32, host memory).  out_ids / out_dist: b x k, host.
 * Synchronous: results are in the out buffers on return.  1 <= k <= VROD_MAX_K.  In a sharded
 * context it is collective (same queries on every rank) and every rank receives the global result. */
VROD_API vrod_status vrod_collection_search(vrod_collection *c, const float *queries, uint32_t b, uint32_t k,
                                            uint64_t *out_ids, float *out_dist);

/* Same, with queries and outputs already resident in device memory of the context's GPU.  The work is enqueued on
 * vrod_ctx_stream(); single-query scans return without synchronising, a batch answered by the tensor-core pass
 * synchronises once (its guard flags are read on the host).  Precondition (checked by the host-buffer call, NOT
 * here): queries are finite, |q_j| <= 2^40 and 2^-40 <= ||q|| <= 2^50, and not all-zero under the cosine metric --
 * outside that range the f32 / bf16 error bounds of the fast passes do not hold; use vrod_collection_search or
 * vrod_collection_set_path(c, 2).  A peer that never takes part in a sharded search is reported by
 * vrod_ctx_synchronize (VROD_ENCCL). */
VROD_API vrod_status vrod_collection_search_device(vrod_collection *c, const float *d_queries, uint32_t b,
                                                   uint32_t k, uint64_t *d_out_ids, float *d_out_dist);

/* Force a path for tests and benches: 0 = automatic (default), 1 = f32 scan + rerank only where the
 * guard holds else exact (same as auto but never the batched path), 2 = always the exact f64 scan,
 * 3 = always the tensor-core batched path (bf16 operand mirror of the rows, built on first use; falls back to
 * feeding the f32 rows as tf32 when the mirror does not fit in device memory), 4 = the batched path with
 * tf32 operands and no mirror. */
VROD_API vrod_status vrod_collection_set_path(vrod_collection *c, int path);

VROD_API const char *vrod_last_error(void);
VROD_API const char *vrod_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VROD_KNN_H */
