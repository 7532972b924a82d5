// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// A thin extern "C" wrapper over the UNMODIFIED reference classes, compiled together with the
// reference's own translation units where they lie (/root/reference/src/index_bipartite.cpp,
// /root/reference/src/index.cpp) by oracle/Makefile into oracle/_ref/libroargraph_ref.so.
// It repeats the call sequences of the reference's CLI drivers
//   tests/test_build_roargraph.cpp:105-136   (ref_build_index)
//   tests/test_search_roargraph.cpp:160-209  (ref_index_open / ref_index_search)
// so that the C restatement in oracle/roargraph_oracle.c and the CUDA path can be checked
// against the real thing, and so that bench.py --impl reference can time the reference's own
// OpenMP search loop.
#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "efanna2e/distance.h"
#include "efanna2e/neighbor.h"
#include "efanna2e/parameters.h"
#include "efanna2e/util.h"
#include "index_bipartite.h"

namespace {
// The reference prints progress to std::cout; silence it while we are inside it.
struct CoutMute {
    std::streambuf *old;
    std::ostringstream sink;
    CoutMute() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutMute() { std::cout.rdbuf(old); }
};
thread_local std::string g_err;
efanna2e::Metric to_metric(int m) { return static_cast<efanna2e::Metric>(m); }
}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

int ref_omp_num_procs() { return omp_get_num_procs(); }

// Distance::compare through the same virtual dispatch the reference uses (src/index.cpp:8-26).
float ref_distance(int metric, const float *a, const float *b, unsigned dim) {
    static efanna2e::DistanceL2 l2;
    static efanna2e::DistanceInnerProduct ip;
    const efanna2e::Distance *d = (metric == efanna2e::L2) ? (const efanna2e::Distance *)&l2 : &ip;
    return d->compare(a, b, dim);
}

void ref_distance_batch(int metric, const float *a, const float *b, unsigned dim, uint64_t n, float *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = ref_distance(metric, a + i * dim, b + i * dim, dim);
}

// NeighborPriorityQueue driven by a script: for each op, kind 0 = insert(id, dist),
// kind 1 = closest_unexpanded() (only when has_unexpanded_node()).  Dumps the final pool.
// Returns pool size.  out_pop[i] receives the id returned by the i-th pop.
uint32_t ref_pool_script(uint32_t capacity, uint32_t nops, const uint8_t *kind, const uint32_t *ids,
                         const float *dists, uint32_t *out_ids, float *out_dists, uint8_t *out_flags,
                         uint32_t *out_pop, uint32_t *n_pop) {
    efanna2e::NeighborPriorityQueue q(capacity);
    uint32_t np = 0;
    for (uint32_t i = 0; i < nops; ++i) {
        if (kind[i] == 0) {
            q.insert(efanna2e::Neighbor(ids[i], dists[i], false));
        } else if (q.has_unexpanded_node()) {
            out_pop[np++] = q.closest_unexpanded().id;
        }
    }
    *n_pop = np;
    for (size_t i = 0; i < q.size(); ++i) {
        out_ids[i] = q[i].id;
        out_dists[i] = q[i].distance;
        out_flags[i] = q[i].flag ? 1 : 0;
    }
    return (uint32_t)q.size();
}

// tests/test_build_roargraph.cpp:105-136.  All inputs are files in the reference's formats.
int ref_build_index(const char *base_fbin, const char *train_fbin, const char *knn_ibin, const char *out_index,
                    int metric, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq, uint32_t num_threads,
                    double *seconds) {
    try {
        CoutMute mute;
        uint32_t base_num, base_dim, sq_num, sq_dim;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        efanna2e::load_meta<float>(train_fbin, sq_num, sq_dim);
        float *data_bp = nullptr, *data_sq = nullptr;
        efanna2e::load_data<float>(base_fbin, base_num, base_dim, data_bp);
        efanna2e::load_data<float>(train_fbin, sq_num, sq_dim, data_sq);
        efanna2e::Parameters parameters;
        efanna2e::IndexBipartite index(base_dim, base_num + sq_num, to_metric(metric), nullptr);
        parameters.Set<uint32_t>("M_sq", M_sq);
        parameters.Set<uint32_t>("M_pjbp", M_pjbp);
        parameters.Set<uint32_t>("L_pjpq", L_pjpq);
        parameters.Set<uint32_t>("num_threads", num_threads);
        index.LoadLearnBaseKNN(knn_ibin);
        omp_set_num_threads(num_threads);
        auto s = std::chrono::high_resolution_clock::now();
        index.BuildRoarGraph(sq_num, data_sq, base_num, data_bp, parameters);
        auto e = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(e - s).count();
        index.SaveProjectionGraph(out_index);
        free(data_bp);
        free(data_sq);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return 1;
    }
}

struct RefIndex {
    efanna2e::IndexBipartite *index;
    uint32_t dim;
    bool need_normalize;
};

// tests/test_search_roargraph.cpp:160-173
void *ref_index_open(const char *base_fbin, const char *index_path, int metric, uint32_t num_threads) {
    try {
        CoutMute mute;
        uint32_t base_num, base_dim;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        auto *ri = new RefIndex;
        uint32_t dim_aligned = (base_dim + 7) / 8 * 8;
        ri->index = new efanna2e::IndexBipartite(dim_aligned, base_num, to_metric(metric), nullptr);
        ri->index->LoadSearchNeededData(base_fbin, "");
        ri->index->LoadProjectionGraph(index_path);
        ri->index->InitVisitedListPool(num_threads);
        ri->dim = dim_aligned;
        ri->need_normalize = ri->index->need_normalize;
        return ri;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return nullptr;
    }
}

uint32_t ref_index_dim(void *h) { return static_cast<RefIndex *>(h)->dim; }

// tests/test_search_roargraph.cpp:196-209: 100 sequential warm-up queries (if warmup != 0), then the
// timed `#pragma omp parallel for schedule(dynamic, 1)` loop.  `queries` must be rows of
// ref_index_dim() floats (zero padded like data_align does).  Returns 0 on success.
int ref_index_search(void *h, const float *queries, uint64_t nq, uint32_t k, uint32_t L, uint32_t num_threads,
                     int warmup, uint32_t *ids, float *dists, uint32_t *cmps, uint32_t *hops, double *seconds) {
    RefIndex *ri = static_cast<RefIndex *>(h);
    try {
        efanna2e::Parameters parameters;
        parameters.Set<uint32_t>("L_pq", L);
        parameters.Set<uint32_t>("num_threads", num_threads);
        omp_set_num_threads(num_threads);
        std::vector<std::vector<float>> res_dists(nq, std::vector<float>(k, 0.0f));
        if (warmup) {
            for (size_t i = 0; i < 100 && i < nq; ++i) {
                ri->index->SearchRoarGraph(queries + i * ri->dim, k, i, parameters, ids + i * k, res_dists[i]);
            }
        }
        std::string err;
        auto s = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t i = 0; i < nq; ++i) {
            try {
                auto r = ri->index->SearchRoarGraph(queries + i * ri->dim, k, i, parameters, ids + i * k,
                                                    res_dists[i]);
                if (cmps) cmps[i] = r.first;
                if (hops) hops[i] = r.second;
            } catch (const std::exception &ex) {
#pragma omp critical
                err = ex.what();
            }
        }
        auto e = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(e - s).count();
        if (!err.empty()) {
            g_err = err;
            return 2;
        }
        for (size_t i = 0; i < nq; ++i) memcpy(dists + i * k, res_dists[i].data(), k * sizeof(float));
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return 1;
    }
}

void ref_index_close(void *h) {
    RefIndex *ri = static_cast<RefIndex *>(h);
    // the reference never frees its base copy (src/index_bipartite.cpp:40); we leak likewise.
    delete ri;
}

}  // extern "C"
