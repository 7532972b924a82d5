// Drop-in for the reference's efanna2e::IndexBipartite (include/index_bipartite.h:23-171), restricted to the
// RoarGraph path (tests/test_search_roargraph.cpp, tests/test_build_roargraph.cpp).  Same public names and
// argument meaning; search runs on the GPU through the C ABI in include/roargraph_b200.h; graph construction runs on
// the host CPU like the reference (edge-identical at one thread) or, with Parameters "gpu_build" = 1, on the GPU.  The legacy bipartite methods of the reference are out of scope.
#pragma once
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "efanna2e/index.h"
#include "efanna2e/neighbor.h"
#include "efanna2e/parameters.h"
#include "efanna2e/util.h"

struct rg_index;  // include/roargraph_b200.h

namespace efanna2e {

class IndexBipartite : public Index {
   public:
    using CompactGraph = std::vector<std::vector<uint32_t>>;

    // index_bipartite.h:27 - n is ignored by the RoarGraph path (sizes come from the loaded files)
    explicit IndexBipartite(size_t dimension, size_t n, Metric m, Index *initializer);
    ~IndexBipartite() override;

    // ---- load / save (src/index_bipartite.cpp:2661-2692, 2097-2117, 2606-2619, 2622-2639) ----
    void LoadSearchNeededData(const char *base_file, const char *sampled_query_file) {
        LoadVectorData(base_file, sampled_query_file);
    }
    void LoadVectorData(const char *base_file, const char *sampled_query_file);
    void LoadProjectionGraph(const char *filename);
    void SaveProjectionGraph(const char *filename);
    void LoadLearnBaseKNN(const char *filename);

    // ---- search (src/index_bipartite.cpp:2311-2420) ----
    // Kept for API parity: the GPU path needs no visited-list pool; this uploads the index to the device.
    void InitVisitedListPool(uint32_t num_threads);
    // One query (parameters must hold "L_pq").  Returns {cmps, hops}.  Re-entrant like the reference's: concurrent callers
    // are micro-batched into one GPU launch (see PendingQuery below); a lone caller pays a launch + sync per query, so
    // batch callers should use SearchRoarGraphBatch.
    std::pair<uint32_t, uint32_t> SearchRoarGraph(const float *query, size_t k, size_t &qid,
                                                  const Parameters &parameters, unsigned *indices,
                                                  std::vector<float> &res_dists);
    // What the drop-in driver calls in place of the reference's OpenMP loop: all queries in one launch.
    // queries: nq rows of GetDimension() floats; indices/dists: nq*k; cmps/hops: nq (may be null).
    void SearchRoarGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters,
                              unsigned *indices, float *dists, uint32_t *cmps, uint32_t *hops);

    // ---- build (src/index_bipartite.cpp:143-233) ----
    void BuildRoarGraph(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data,
                        const Parameters &parameters);

    CompactGraph &GetProjectionGraph() { return projection_graph_; }
    std::vector<std::vector<uint32_t>> &GetLearnBaseKNN() { return learn_base_knn_; }
    // the pruned pivot list one training query produces in P1 of LinkProjection (test hook, see index_bipartite.cpp)
    static std::vector<uint32_t> PivotProjectionList(const float *base, size_t dim, const Distance *dist, uint32_t M_pjbp,
                                                     const uint32_t *nn, size_t n_nn);
    uint32_t GetProjectionEp() const { return projection_ep_; }
    void SetProjectionGraph(uint32_t ep, CompactGraph graph);  // adopt an externally built graph
    void SetBaseData(const float *base, size_t n);             // adopt caller-owned padded rows
    void SetDevice(int device) { device_ = device; }
    // Search on `count` GPUs (devices device_ .. device_ + count - 1): the index is replicated on each of them and every
    // SearchRoarGraphBatch call splits its queries into `count` contiguous slices, one host thread per GPU - the
    // query sharding the reference does with OpenMP threads (tests/test_search_roargraph.cpp:203).  No collective.
    void SetDeviceCount(int count) { device_count_ = count < 1 ? 1 : count; }

    bool need_normalize = false;  // index_bipartite.h:145 (COSINE)

   private:
    void upload_to_device();
    void release_device();
    // graph construction steps (see src/index_bipartite.cpp file:line in index_bipartite.cpp)
    void calculate_projection_ep();
    void link_projection(const Parameters &parameters);
    void build_on_device(const Parameters &parameters);  // Parameters "gpu_build" != 0

    Index *initializer_;
    CompactGraph projection_graph_, supply_nbrs_, learn_base_knn_;
    std::vector<std::mutex> locks_;
    uint32_t projection_ep_ = 0;
    float *owned_base_ = nullptr;  // allocated by LoadVectorData, never freed by the reference either
    rg_index *device_index_ = nullptr;          // replica on device_
    std::vector<rg_index *> extra_replicas_;    // replicas on device_ + 1 .. device_ + device_count_ - 1
    int device_ = 0, device_count_ = 1;
    std::mutex device_mutex_;
    std::mutex search_mutex_;  // the GPU replicas own one set of scratch buffers each: concurrent callers take turns

    // Micro-batching of concurrent SearchRoarGraph callers (the reference's OpenMP loop, tests/test_search_roargraph.cpp:
    // 203-209, kept as is by a caller): the first caller to arrive leads - it waits microbatch_us_ for others, packs every
    // pending query with the same (k, L_pq) into ONE rg_search_batch launch and hands the results back.
    struct PendingQuery {
        const float *query;
        size_t k;
        uint32_t L;
        unsigned *indices;
        float *dists;
        uint32_t cmps = 0, hops = 0;
        bool done = false, short_result = false;
        std::string error;
    };
    void lead_microbatch(std::unique_lock<std::mutex> &lk);
    std::mutex mb_mutex_;
    std::condition_variable mb_cv_;
    std::vector<PendingQuery *> mb_pending_;
    bool mb_leader_active_ = false;
    int microbatch_us_ = 50;             // RG_MICROBATCH_US; 0 = every call is its own batch of one
    float *mb_queries_ = nullptr;        // page-locked staging (queries in, results out) for up to kMicrobatchMax queries
    unsigned *mb_ids_ = nullptr;
    float *mb_dists_ = nullptr;
    uint32_t *mb_cmps_ = nullptr, *mb_hops_ = nullptr;
    size_t mb_k_cap_ = 0;
    static constexpr size_t kMicrobatchMax = 4096;
};

}  // namespace efanna2e
