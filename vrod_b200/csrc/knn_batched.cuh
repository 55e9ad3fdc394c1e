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

// Can the tensor-core path answer this (shape, k) at all?
bool batched_supported(const ShardView &s, uint32_t b, uint32_t k);

// Enqueue the batched search of b queries (d_q: b x ld) on `st`.  `scratch`/`scratch_bytes` is a
// device buffer the callee may grow.  status[qi] = 1 marks a query whose guard failed (the caller
// rescans it); out receives b x k hits.  ev_start/ev_stop, when given, bracket the tile kernel.
cudaError_t launch_batched_search(const ShardView &s, const float *d_q, uint32_t b, uint32_t k, int sm_count,
                                  void **scratch, size_t *scratch_bytes, int *status, Hit *out, cudaStream_t st,
                                  BatchedStats *stats, cudaEvent_t ev_start = nullptr, cudaEvent_t ev_stop = nullptr);

}  // namespace vrod
