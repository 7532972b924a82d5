// K2/K3/K4 - build-time exact kNN (placeholder until the tcgen05 path lands; fails loudly).
#include "rg_common.cuh"

extern "C" {
rg_status rg_knn_exact(const float *, uint64_t, uint64_t, const float *, uint64_t, uint32_t, int, uint32_t,
                       uint32_t *, float *, int) {
    return rg::fail(RG_ERR_INTERNAL, "rg_knn_exact: not implemented yet");
}
rg_status rg_knn_exact_device(const float *, uint64_t, uint64_t, const float *, uint64_t, uint32_t, int, uint32_t,
                              uint32_t *, float *, int, void *) {
    return rg::fail(RG_ERR_INTERNAL, "rg_knn_exact_device: not implemented yet");
}
rg_status rg_knn_merge_device(const uint32_t *, const float *, uint32_t, uint64_t, uint32_t, int, uint32_t *,
                              float *, int, void *) {
    return rg::fail(RG_ERR_INTERNAL, "rg_knn_merge_device: not implemented yet");
}
}
