// knn_scan.cuh -- host-visible launch interface of the scan kernels (knn_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrod {

struct ShardView {
    const float *rows;       // n x ld f32, row-major, rows zero-padded from dim to ld (ld % 4 == 0)
    const float *inv_norm;   // n f32: 1/||x|| (0 for a zero row)               [cosine fast scan]
    const float *sq_norm;    // n f32: ||x||^2                                  [batched L2 path]
    const unsigned int *maxnorm_bits;  // f32 bits of max ||x||^2 over the shard [batched guard]
    unsigned int *mirror_stats;        // [2] f32 bits of max ||x~ - x||^2 and max ||x~||^2 over the mirrored rows (x~: the bf16 mirror row) [bf16 guard]
    const unsigned short *rows_h;      // tiled bf16 mirror of the rows (knn_batched.cu: mirror_bytes), nullptr if none [batched bf16 mode]
    uint32_t n;              // local rows (< 2^32)
    uint32_t dim, ld;
    int metric;              // 0 Euclidean, 1 cosine
    uint32_t rank, world;    // which shard of how many: ids are dealt to the shards in blocks of kShardBlock rows (row_id)
};

// Row sharding: BLOCK-CYCLIC.  Global ids are dealt out in blocks of kShardBlock consecutive rows, block j to shard j mod G;
// a shard stores its blocks back to back.  Every shard then holds the same number of rows (within one block) at ANY fill
// level, so a half-filled collection scans as fast as a full one, and it grows on its own: more rows just continue the
// deal, no row ever moves between shards (round 1 cut the id space into G contiguous ranges sized by the capacity: a
// collection filled shard 0 first and could not grow without re-dealing everything).  Inside a shard local order is id
// order; across shards ids interleave, so lists are merged under the full (dist, id) order, not by shard number.
constexpr uint32_t kShardBlock = 4096;
__host__ __device__ inline uint64_t row_id(uint64_t local_row, uint32_t rank, uint32_t world) {
    return ((local_row / kShardBlock) * world + rank) * kShardBlock + local_row % kShardBlock;
}
// rows shard `rank` holds when the collection holds `count` rows
inline uint64_t shard_rows_at(uint64_t count, uint32_t rank, uint32_t world) {
    const uint64_t cycle = (uint64_t)kShardBlock * world, full = count / cycle, rem = count % cycle;
    const uint64_t lo = (uint64_t)rank * kShardBlock;
    return full * kShardBlock + (rem <= lo ? 0 : (rem - lo < kShardBlock ? rem - lo : kShardBlock));
}

// Per-launch scratch owned by the collection (sized by scan_scratch_bytes()).
struct ScanScratch {
    unsigned long long *blk_cand;  // [max_grid][kprime] per-CTA best keys
    unsigned int *ticket;          // self-resetting last-CTA ticket
    int *status;                   // [b] 0 = answered exactly, 1 = guard failed -> exact rescan needed
    unsigned long long *counters;  // [0] += 1 per query whose guard failed
};

struct ScanPlan {
    int grid;        // CTAs
    int kprime;      // candidates kept by the f32 scan (>= k + 16, power of two)
    int cap;         // candidate store capacity (keys) per CTA
    size_t smem;     // dynamic shared memory per CTA
    double eps;      // error bound of the f32 pass (relative for L2, absolute on cos-sim for cosine)
};

// Result entry exchanged between ranks: 16 bytes.
struct Hit {
    unsigned long long id;
    float dist;
    uint32_t pad;
};

// Fused exchange + merge over NVLink peer memory (knn_scan.cu: exchange_device).  It serves searches with
// b <= kXchgMaxB queries and b*k <= kXchgMaxHits hits on at most kXchgMaxWorld ranks.
constexpr uint32_t kXchgMaxWorld = 16, kXchgMaxB = 256, kXchgMaxHits = 4096, kXchgSlots = 3;
size_t xchg_window_bytes();
constexpr uint32_t kXchgAllRanks = 0xffffffffu;   // root: every rank merges; else only rank `root` receives and merges
struct XchgArgs {
    unsigned char *const *windows;   // device table of every rank's window (peer mappings); nullptr = no exchange
    uint32_t rank, world, seq, root;
    uint32_t qi;                     // query index inside the search call (its flag / data slot in the windows)
    int *err;                        // set to 1 when a peer never arrived
};
// exchange + merge as a kernel of its own (one CTA per query): after the batched pass
cudaError_t launch_exchange_merge(const XchgArgs &x, const Hit *local, uint32_t b, uint32_t k, unsigned long long *out_ids,
                                  float *out_dist, cudaStream_t st);
int scan_sm_count(int device);
unsigned long long *scan_debug_enable();   // development aid: device buffer of 8 globaltimer stamps (VROD_SCAN_DEBUG)
ScanPlan make_scan_plan(const ShardView &s, uint32_t k, int sm_count, bool exact);
size_t scan_cand_bytes(int sm_count);

// f32 scan + exact rerank + guard of ONE query (q: ld floats, device).  Writes k hits to `out` and, when
// out_ids/out_dist are not null, the final ids / distances as well (single-GPU: no merge kernel needed).
// x (optional): the last CTA also pushes the k hits to the peers, waits for theirs and writes the GLOBAL answer to
// out_ids / out_dist; *status then is the OR of all ranks' guard flags.  Needs world * k <= plan.cap.
cudaError_t launch_fast_scan(const ShardView &s, const float *q, uint32_t k, const ScanPlan &plan,
                             const ScanScratch &scr, int *status, Hit *out, unsigned long long *out_ids, float *out_dist,
                             cudaStream_t st, const XchgArgs *x = nullptr);
// exact f64 scan of ONE query; if only_if_flag != nullptr the grid returns at once unless *only_if_flag != 0.
// With x, *gstatus (optional) receives the OR of the ranks' guard flags that travelled with the lists.
cudaError_t launch_exact_scan(const ShardView &s, const float *q, uint32_t k, const ScanPlan &plan,
                              const ScanScratch &scr, const int *only_if_flag, Hit *out, unsigned long long *out_ids,
                              float *out_dist, cudaStream_t st, const XchgArgs *x = nullptr, int *gstatus = nullptr);

// rows [row0, row0+n) of the shard: inv_norm / sq_norm, and flags[0] |= 1 if a value is not finite,
// |= 2 if a value or norm is outside the range the f32 scan's error bound covers.
cudaError_t launch_row_norms(const float *rows, uint32_t row0, uint32_t n, uint32_t ld, float *inv_norm,
                             float *sq_norm, int *flags, cudaStream_t st);
// synthetic rows into the local rows [row0, row0 + n) of shard `rank` of `world` (their global ids by row_id)
cudaError_t launch_fill_synthetic(float *rows, uint32_t row0, uint32_t n, uint32_t dim, uint32_t ld, uint32_t rank, uint32_t world,
                                  uint64_t seed, cudaStream_t st);
// [b x dim] -> [b x ld] zero padded
cudaError_t launch_pad_queries(const float *src, float *dst, uint32_t b, uint32_t dim, uint32_t ld, cudaStream_t st);
// merge g lists of [b][k] hits (layout [g][b][k]) into ids/dist [b][k] by (dist, id)
cudaError_t launch_merge_hits(const Hit *lists, uint32_t g, uint32_t b, uint32_t k, unsigned long long *out_ids,
                              float *out_dist, cudaStream_t st);

}  // namespace vrod
