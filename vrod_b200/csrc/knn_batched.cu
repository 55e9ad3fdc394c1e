// knn_batched.cu -- batched-query path.  Not built yet in this revision: batched_supported() says no
// and every batch is answered by per-query scans (knn_scan.cu).
#include "knn_batched.cuh"

namespace vrod {

bool batched_supported(const ShardView &, uint32_t, uint32_t) { return false; }

cudaError_t launch_batched_search(const ShardView &, const float *, uint32_t, uint32_t, int, void **, size_t *, int *,
                                  Hit *, cudaStream_t, BatchedStats *) {
    return cudaErrorNotSupported;
}

}  // namespace vrod
