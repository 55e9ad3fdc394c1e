// knn_batched.cuh -- batched-query path (tensor cores): Q . X^T tiles with a fused candidate filter,
// followed by the exact f64 rerank.  See knn_batched.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "knn_scan.cuh"

namespace vrod {

struct BatchedStats {
    uint64_t launches;
    uint64_t tiles;
};

// bf16 operand mirror of the rows (see knn_batched.cu): kd = dim rounded up to 16 data columns + 16 aux columns
uint32_t mirror_kd(uint32_t dim);
uint32_t mirror_ld(uint32_t dim);
// bytes of the TILED mirror of `rows` rows (128-row tiles x mirror_ld/16 K-step blocks of 4 KB, see knn_batched.cu)
size_t mirror_bytes(uint64_t rows, uint32_t dim);
// rows [row0, row0 + n) of the shard -> the tiles of rows_h that hold them (whole tiles are rewritten; rows beyond s.n
// become zero rows); needs the rows' norms (launch_row_norms)
cudaError_t launch_build_mirror(const ShardView &s, unsigned short *rows_h, uint32_t row0, uint32_t n, cudaStream_t st);

// Can the tensor-core path answer this (shape, k) at all?
bool batched_supported(const ShardView &s, uint32_t b, uint32_t k);

// Enqueue the batched search of b queries (d_q: b x ld) on `st`.  `scratch`/`scratch_bytes` is a
// device buffer the callee may grow.  status[qi] = 1 marks a query whose guard failed (the caller
// rescans it); out receives b x k hits and, when out_ids/out_dist are not null, the final ids / distances as well
// (single-GPU contexts: no merge kernel needed).  ev_start/ev_stop, when given, bracket the tile kernels.
// band_mode: keep every candidate within the approximate surrogate's error band above the k-th best one (up to 1024 per
// query) instead of a fixed k' of them -- for collections whose neighbours the bf16 contraction cannot tell apart (tight
// clusters under the Euclidean metric), where the fixed-k' proof fails for every query.
// guess: filter every phase at a GUESSED threshold -- the rank of the keys seen so far below which the phase is expected
// to find its k' keys if the rows still to come resemble the rows already seen -- instead of the k'-th best key so far
// (2-3x fewer candidates, half as many phases).  A guess that did not hold is detected (status[qi] = 2: the caller
// rescans the query, and stops guessing for a collection whose row order defeats it).
// wide_margin: k' = pow2 >= 2k + 16 (round 1's rule) instead of 1.5k + 16: the first thing to try when the proofs of a
// collection fail (high-dimensional data whose distances concentrate), before band mode.
cudaError_t launch_batched_search(const ShardView &s, const float *d_q, uint32_t b, uint32_t k, int sm_count,
                                  void **scratch, size_t *scratch_bytes, int *status, Hit *out, unsigned long long *out_ids,
                                  float *out_dist, cudaStream_t st, BatchedStats *stats, cudaEvent_t ev_start = nullptr,
                                  cudaEvent_t ev_stop = nullptr, bool band_mode = false, bool guess = false, bool wide_margin = false);

}  // namespace vrod
