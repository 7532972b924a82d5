// Internal interface of the exact-kNN pipeline (rg_knn.cu) for the multi-GPU layer (rg_knn_sharded.cu).
#pragma once
#include <cstring>

#include "rg_common.cuh"

namespace rg {
namespace knn {

// K2/K2s/K3 (+ second pass / exact scan) for one base shard; all pointers are device memory on the current device.  Work is
// enqueued on `st`; the call returns after synchronising it.  reuse_base: the FP16 copy of the shard made by the previous
// call on this device is still valid (same pointer, row count and dim, rows unchanged).  stats[3] = {kernel launches,
// queries redone by the exact FP32 scan, queries that needed the conservative second pass}.
rg_status knn_device(const float *d_base, uint64_t n, uint64_t id_base, const float *d_queries, uint64_t nq, uint32_t dim,
                     int metric, uint32_t K, uint32_t *d_ids, float *d_dists, cudaStream_t st, uint64_t *stats, bool reuse_base);
void release_scratch(int device);
void set_last_stats(const uint64_t stats[3]);  // what rg_knn_last_stats() reports for the calling thread

}  // namespace knn
}  // namespace rg
