// Small extern "C" handle over the host C++ class so the Python harness (tests/, bench.py) can drive the
// drop-in layer exactly as the CLI drivers do: file in -> BuildRoarGraph -> file out, and file in -> batched search.
#include <cstring>
#include <sstream>
#include <string>

#include "index_bipartite.h"

namespace {
thread_local std::string g_err;
struct Mute {  // the build prints progress like the reference; keep harness logs clean
    std::streambuf *old;
    std::ostringstream sink;
    explicit Mute(bool on) : old(on ? std::cout.rdbuf(sink.rdbuf()) : nullptr) {}
    ~Mute() {
        if (old) std::cout.rdbuf(old);
    }
};
}  // namespace

extern "C" {
__attribute__((visibility("default"))) const char *rgh_last_error() { return g_err.c_str(); }

// tests/test_build_roargraph.cpp:105-136 on in-memory arrays.  base/train: padded rows (dim % 8 == 0);
// knn_ids: n_train x knn_k.  Writes the projection index file.  Returns 0 on success.
__attribute__((visibility("default"))) int rgh_build_index(const float *base, uint64_t n, const float *train,
                                                            uint64_t n_train, uint32_t dim, int metric,
                                                            const uint32_t *knn_ids, uint32_t knn_k, uint32_t M_sq,
                                                            uint32_t M_pjbp, uint32_t L_pjpq, uint32_t num_threads,
                                                            const char *out_index, int quiet, double *seconds) {
    try {
        Mute mute(quiet != 0);
        efanna2e::IndexBipartite index(dim, n + n_train, static_cast<efanna2e::Metric>(metric), nullptr);
        auto &knn = index.GetLearnBaseKNN();
        knn.resize(n_train);
        for (uint64_t i = 0; i < n_train; ++i) knn[i].assign(knn_ids + i * knn_k, knn_ids + (i + 1) * knn_k);
        efanna2e::Parameters p;
        p.Set<uint32_t>("M_sq", M_sq);
        p.Set<uint32_t>("M_pjbp", M_pjbp);
        p.Set<uint32_t>("L_pjpq", L_pjpq);
        p.Set<uint32_t>("num_threads", num_threads);
        auto s = std::chrono::high_resolution_clock::now();
        index.BuildRoarGraph(n_train, train, n, base, p);
        auto e = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(e - s).count();
        index.SaveProjectionGraph(out_index);
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return 1;
    }
}

// P1 lists of the host build (PruneBiSearchBaseGetBase per training query): out [n_train][M_pjbp + 1], word 0 = length.
__attribute__((visibility("default"))) int rgh_projection_lists(const float *base, uint32_t dim, int metric, const uint32_t *knn_ids,
                                                                 uint64_t n_train, uint32_t knn_k, uint32_t M_sq, uint32_t M_pjbp,
                                                                 uint32_t *out) {
    try {
        efanna2e::IndexBipartite index(dim, 1, static_cast<efanna2e::Metric>(metric), nullptr);
        const efanna2e::Distance *dist = index.GetDistance();
        const size_t take = std::min<size_t>(knn_k, M_sq);
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t i = 0; i < (int64_t)n_train; ++i) {
            const std::vector<uint32_t> l =
                efanna2e::IndexBipartite::PivotProjectionList(base, dim, dist, M_pjbp, knn_ids + (size_t)i * knn_k, take);
            uint32_t *row = out + (size_t)i * (M_pjbp + 1);
            row[0] = (uint32_t)l.size();
            for (size_t j = 0; j < l.size(); ++j) row[1 + j] = l[j];
            for (size_t j = l.size(); j < M_pjbp; ++j) row[1 + j] = 0;
        }
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return 1;
    }
}

// The reference driver's own search loop (tests/test_search_roargraph.cpp:160-209) on files: load base + index, then ONE
// SearchRoarGraph CALL PER QUERY from num_threads OpenMP threads (schedule(dynamic,1)) - the way existing callers of the
// reference use the class.  queries: nq padded rows.  Outputs: ids/dists [nq*k], cmps/hops [nq].
__attribute__((visibility("default"))) int rgh_search_per_query(const char *base_fbin, const char *index_file, int metric,
                                                                 const float *queries, uint64_t nq, uint32_t k, uint32_t L_pq,
                                                                 uint32_t num_threads, uint32_t *ids, float *dists,
                                                                 uint32_t *cmps, uint32_t *hops, double *loop_seconds) {
    try {
        Mute mute(true);
        uint32_t base_num = 0, base_dim = 0;
        efanna2e::load_meta<float>(base_fbin, base_num, base_dim);
        const uint32_t dim = (base_dim + 7) / 8 * 8;
        efanna2e::IndexBipartite index(dim, base_num, static_cast<efanna2e::Metric>(metric), nullptr);
        index.LoadSearchNeededData(base_fbin, "");
        index.LoadProjectionGraph(index_file);
        index.InitVisitedListPool(num_threads);
        efanna2e::Parameters p;
        p.Set<uint32_t>("L_pq", L_pq);
        std::string first_error;
        {   // warm-up like the reference's first 100 sequential queries (tests/test_search_roargraph.cpp:198-200)
            size_t qid = 0;
            std::vector<float> res_dists(k);
            std::vector<unsigned> tmp(k);
            for (uint64_t i = 0; i < std::min<uint64_t>(nq, 8); ++i) {
                try {
                    index.SearchRoarGraph(queries + (size_t)i * dim, k, qid, p, tmp.data(), res_dists);
                } catch (const std::exception &) {
                }
            }
        }
        const auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads)
        for (int64_t i = 0; i < (int64_t)nq; ++i) {
            try {
                size_t qid = (size_t)i;
                std::vector<float> res_dists(k);
                auto ch = index.SearchRoarGraph(queries + (size_t)i * dim, k, qid, p, ids + (size_t)i * k, res_dists);
                std::memcpy(dists + (size_t)i * k, res_dists.data(), k * sizeof(float));
                cmps[i] = ch.first;
                hops[i] = ch.second;
            } catch (const std::exception &ex) {
#pragma omp critical
                if (first_error.empty()) first_error = ex.what();
            }
        }
        if (loop_seconds) *loop_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        if (!first_error.empty()) {
            g_err = first_error;
            return 1;
        }
        return 0;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return 1;
    }
}
}
