// Host layer of the drop-in efanna2e::IndexBipartite (RoarGraph path only).
//   * load / save of the reference's file formats
//   * search: forwards to the CUDA library through the C ABI (include/roargraph_b200.h)
//   * BuildRoarGraph: CPU graph construction, a from-scratch restatement of the reference's rules
//     (src/index_bipartite.cpp:143-233, 1043-1277 and the helpers cited below) that reproduces the
//     reference's adjacency byte for byte at num_threads = 1 (tests/test_host_build.py).
// All reference citations are /root/reference/src/index_bipartite.cpp unless another file is named.
#include "index_bipartite.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/roargraph_b200.h"

namespace efanna2e {

// ---- Index base (src/index.cpp:8-26) -------------------------------------------------------------
Index::Index(size_t dimension, size_t n, Metric metric) : dimension_(dimension), nd_(n), metric_(metric) {
    switch (metric_) {
        case INNER_PRODUCT:
        case COSINE:
            distance_ = new DistanceInnerProduct();
            break;
        default:
            distance_ = new DistanceL2();
            break;
    }
}
Index::~Index() { delete distance_; }

IndexBipartite::IndexBipartite(size_t dimension, size_t n, Metric m, Index *initializer)
    : Index(dimension, n, m), initializer_(initializer) {
    if (m == COSINE) need_normalize = true;  // :30-38
    if (const char *e = std::getenv("RG_MICROBATCH_US")) microbatch_us_ = std::max(0, atoi(e));
}

IndexBipartite::~IndexBipartite() {
    release_device();
    for (void *p : {(void *)mb_queries_, (void *)mb_ids_, (void *)mb_dists_, (void *)mb_cmps_, (void *)mb_hops_})
        if (p) {
            rg_host_unregister(p);
            free(p);
        }
}

namespace {
[[noreturn]] void throw_rg(const char *what) {
    throw std::runtime_error(std::string(what) + ": " + rg_last_error_string());
}
int rg_metric(Metric m) { return m == L2 ? RG_METRIC_L2 : (m == COSINE ? RG_METRIC_COSINE : RG_METRIC_INNER_PRODUCT); }
}  // namespace

// ---- files -----------------------------------------------------------------------------------------
// :2661-2692.  Reads the base fbin into 8-float padded rows, L2-normalises them for COSINE.
void IndexBipartite::LoadVectorData(const char *base_file, const char *sampled_query_file) {
    uint32_t base_num = 0, base_dim = 0, sq_num = 0, q_dim = 0;
    load_meta<float>(base_file, base_num, base_dim);
    if (sampled_query_file && strlen(sampled_query_file) != 0) {
        load_meta<float>(sampled_query_file, sq_num, q_dim);
        if (base_dim != q_dim) throw std::runtime_error("base and query dimension mismatch");
    }
    float *base = nullptr;
    load_data<float>(base_file, base_num, base_dim, base);  // rows already padded to padded_dim(base_dim)
    const uint64_t stride = padded_dim(base_dim);
    if (need_normalize) {
        std::cout << "Normalizing base data" << std::endl;
        for (size_t i = 0; i < base_num; ++i) normalize<float>(base + i * stride, base_dim);
    }
    if (stride != dimension_) {
        free(base);
        throw std::runtime_error("base dimension does not match the index dimension");
    }
    release_device();
    owned_base_ = base;
    data_bp_ = base;
    nd_ = base_num;
    nd_sq_ = sq_num;
}

void IndexBipartite::SetBaseData(const float *base, size_t n) {
    release_device();
    data_bp_ = base;
    nd_ = n;
}

// :2097-2117
void IndexBipartite::LoadProjectionGraph(const char *filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error(std::string("cannot open file ") + filename);
    uint32_t npts = 0;
    in.read(reinterpret_cast<char *>(&projection_ep_), sizeof(uint32_t));
    in.read(reinterpret_cast<char *>(&npts), sizeof(uint32_t));
    std::cout << "Projection graph, ep: " << projection_ep_ << std::endl;
    // header and degree words are checked against the bytes that are really there: a truncated or foreign file must not
    // turn into a huge resize here or into out-of-range row gathers on the GPU
    in.seekg(0, std::ios::end);
    const uint64_t file_bytes = static_cast<uint64_t>(in.tellg());
    in.seekg(2 * sizeof(uint32_t), std::ios::beg);
    if (!in || file_bytes < 2 * sizeof(uint32_t) || uint64_t(npts) * sizeof(uint32_t) > file_bytes - 2 * sizeof(uint32_t))
        throw std::runtime_error("projection graph file truncated");
    uint64_t left = file_bytes - 2 * sizeof(uint32_t);
    projection_graph_.assign(npts, {});
    double total = 0;
    for (uint32_t i = 0; i < npts; ++i) {
        uint32_t deg = 0;
        in.read(reinterpret_cast<char *>(&deg), sizeof(uint32_t));
        if (!in || left < sizeof(uint32_t) || uint64_t(deg) * sizeof(uint32_t) > left - sizeof(uint32_t))
            throw std::runtime_error("projection graph file truncated");
        left -= sizeof(uint32_t) + uint64_t(deg) * sizeof(uint32_t);
        projection_graph_[i].resize(deg);
        in.read(reinterpret_cast<char *>(projection_graph_[i].data()), std::streamsize(deg) * sizeof(uint32_t));
        total += deg;
    }
    if (!in) throw std::runtime_error("projection graph file truncated");
    std::cout << "Projection graph, avg_degree: " << (npts ? total / npts : 0.0) << std::endl;
    release_device();
}

// :2606-2619
void IndexBipartite::SaveProjectionGraph(const char *filename) {
    std::ofstream out(filename, std::ios::binary | std::ios::out);
    if (!out.is_open()) throw std::runtime_error("cannot open file");
    const uint32_t n = static_cast<uint32_t>(projection_graph_.size());
    out.write(reinterpret_cast<const char *>(&projection_ep_), sizeof(uint32_t));
    out.write(reinterpret_cast<const char *>(&n), sizeof(uint32_t));
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t deg = static_cast<uint32_t>(projection_graph_[i].size());
        out.write(reinterpret_cast<const char *>(&deg), sizeof(uint32_t));
        out.write(reinterpret_cast<const char *>(projection_graph_[i].data()), std::streamsize(deg) * sizeof(uint32_t));
    }
}

// :2622-2639 - header + ids block; the distances behind it are never read
void IndexBipartite::LoadLearnBaseKNN(const char *filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("learn base knn file error");
    uint32_t npts = 0, k = 0;
    in.read(reinterpret_cast<char *>(&npts), sizeof(uint32_t));
    in.read(reinterpret_cast<char *>(&k), sizeof(uint32_t));
    std::cout << "learn base knn npts: " << npts << ", k_dim: " << k << std::endl;
    learn_base_knn_.assign(npts, std::vector<uint32_t>(k));
    for (uint32_t i = 0; i < npts; ++i)
        in.read(reinterpret_cast<char *>(learn_base_knn_[i].data()), std::streamsize(k) * sizeof(uint32_t));
    if (!in || npts == 0) throw std::runtime_error("learn base knn file error");
}

void IndexBipartite::SetProjectionGraph(uint32_t ep, CompactGraph graph) {
    projection_ep_ = ep;
    projection_graph_ = std::move(graph);
    release_device();
}

// ---- search -----------------------------------------------------------------------------------------
void IndexBipartite::release_device() {
    if (device_index_) {
        rg_index_destroy(device_index_);
        device_index_ = nullptr;
    }
    for (rg_index *r : extra_replicas_) rg_index_destroy(r);
    extra_replicas_.clear();
}

void IndexBipartite::upload_to_device() {
    std::lock_guard<std::mutex> g(device_mutex_);
    if (device_index_) return;
    if (!data_bp_ || projection_graph_.empty()) throw std::runtime_error("index not loaded");
    const size_t n = projection_graph_.size();
    if (n != nd_) throw std::runtime_error("graph and base sizes differ");
    std::vector<uint64_t> offsets(n + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + projection_graph_[i].size();
    std::vector<uint32_t> adj(offsets[n]);
    for (size_t i = 0; i < n; ++i) std::copy(projection_graph_[i].begin(), projection_graph_[i].end(), adj.begin() + offsets[i]);
    if (device_ + device_count_ > rg_device_count() && rg_device_count() > 0)
        throw std::runtime_error("SetDeviceCount: only " + std::to_string(rg_device_count()) + " CUDA device(s) visible");
    // one replica per GPU, uploaded concurrently (the host arrays are read-only here)
    std::vector<rg_index *> replicas(device_count_, nullptr);
    std::vector<std::string> errors(device_count_);
    auto upload = [&](int r) {
        if (rg_index_create(&replicas[r], data_bp_, n, (uint32_t)dimension_, rg_metric(metric_), offsets.data(), adj.data(),
                            projection_ep_, device_ + r, 0) != RG_OK)
            errors[r] = rg_last_error_string();
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < device_count_; ++r) threads.emplace_back(upload, r);
    upload(0);
    for (auto &t : threads) t.join();
    for (int r = 0; r < device_count_; ++r)
        if (!replicas[r]) {
            for (rg_index *x : replicas) rg_index_destroy(x);
            throw std::runtime_error("rg_index_create: " + errors[r]);
        }
    device_index_ = replicas[0];
    extra_replicas_.assign(replicas.begin() + 1, replicas.end());
}

void IndexBipartite::InitVisitedListPool(uint32_t) { upload_to_device(); }  // index_bipartite.h:133

void IndexBipartite::SearchRoarGraphBatch(const float *queries, size_t nq, size_t k, const Parameters &parameters,
                                          unsigned *indices, float *dists, uint32_t *cmps, uint32_t *hops) {
    const uint32_t L_pq = parameters.Get<uint32_t>("L_pq");  // :2313
    if (!device_index_) upload_to_device();
    // SearchRoarGraph is re-entrant in the reference (called from OpenMP threads, tests/test_search_roargraph.cpp:203);
    // here every call is a GPU batch on the index's own stream and scratch, so concurrent callers are serialised.
    std::lock_guard<std::mutex> serial(search_mutex_);
    const size_t G = 1 + extra_replicas_.size();
    if (G == 1 || nq < G) {
        const rg_status s = rg_search_batch(device_index_, queries, nq, (uint32_t)k, L_pq, indices, dists, cmps, hops);
        if (s == RG_ERR_NOT_ENOUGH_RESULTS) throw std::runtime_error(rg_last_error_string());  // :2408-2412
        if (s != RG_OK) throw_rg("rg_search_batch");
        return;
    }
    // contiguous query slices, one host thread per replica; results land in disjoint slices of the caller's arrays
    std::vector<rg_status> status(G, RG_OK);
    std::vector<std::string> errors(G);
    auto run = [&](size_t r) {
        const size_t lo = nq * r / G, hi = nq * (r + 1) / G;
        rg_index *ix = r == 0 ? device_index_ : extra_replicas_[r - 1];
        status[r] = rg_search_batch(ix, queries + lo * dimension_, hi - lo, (uint32_t)k, L_pq, indices + lo * k, dists + lo * k,
                                    cmps ? cmps + lo : nullptr, hops ? hops + lo : nullptr);
        if (status[r] != RG_OK) errors[r] = rg_last_error_string();  // the message is thread-local
    };
    std::vector<std::thread> threads;
    for (size_t r = 1; r < G; ++r) threads.emplace_back(run, r);
    run(0);
    for (auto &t : threads) t.join();
    for (size_t r = 0; r < G; ++r) {
        if (status[r] == RG_ERR_NOT_ENOUGH_RESULTS) throw std::runtime_error(errors[r]);  // :2408-2412
        if (status[r] != RG_OK) throw std::runtime_error("rg_search_batch: " + errors[r]);
    }
}

// Leader side of the micro-batch (mb_mutex_ held on entry and on return): wait a moment for more callers, then run every
// pending query that shares (k, L_pq) with the oldest one as ONE batch.
void IndexBipartite::lead_microbatch(std::unique_lock<std::mutex> &lk) {
    if (microbatch_us_ > 0 && mb_pending_.size() < kMicrobatchMax) {
        lk.unlock();
        std::this_thread::sleep_for(std::chrono::microseconds(microbatch_us_));
        lk.lock();
    }
    if (mb_pending_.empty()) return;
    const size_t k = mb_pending_.front()->k;
    const uint32_t L = mb_pending_.front()->L;
    std::vector<PendingQuery *> batch, rest;
    for (PendingQuery *q : mb_pending_) (q->k == k && q->L == L && batch.size() < kMicrobatchMax ? batch : rest).push_back(q);
    mb_pending_.swap(rest);
    lk.unlock();
    const size_t nb = batch.size(), dim = dimension_;
    std::string error;
    try {
        auto pinned = [&](auto *&ptr, size_t count) {  // page-locked once: rg_search_batch then works on it in place
            using T = std::remove_reference_t<decltype(*ptr)>;
            void *raw = nullptr;
            if (posix_memalign(&raw, 4096, count * sizeof(T)) != 0) throw std::bad_alloc();
            ptr = static_cast<T *>(raw);
            (void)rg_host_register(ptr, count * sizeof(T));  // failure only means the staged path is taken
        };
        if (!mb_queries_) {
            pinned(mb_queries_, kMicrobatchMax * dim);
            pinned(mb_cmps_, kMicrobatchMax);
            pinned(mb_hops_, kMicrobatchMax);
        }
        if (mb_k_cap_ < k) {
            for (void *p : {(void *)mb_ids_, (void *)mb_dists_})
                if (p) {
                    rg_host_unregister(p);
                    free(p);
                }
            mb_ids_ = nullptr;
            mb_dists_ = nullptr;
            pinned(mb_ids_, kMicrobatchMax * k);
            pinned(mb_dists_, kMicrobatchMax * k);
            mb_k_cap_ = k;
        }
        for (size_t i = 0; i < nb; ++i) memcpy(mb_queries_ + i * dim, batch[i]->query, dim * sizeof(float));
        if (!device_index_) upload_to_device();
        rg_status s;
        {
            std::lock_guard<std::mutex> serial(search_mutex_);
            s = rg_search_batch(device_index_, mb_queries_, nb, (uint32_t)k, L, mb_ids_, mb_dists_, mb_cmps_, mb_hops_);
        }
        if (s != RG_OK && s != RG_ERR_NOT_ENOUGH_RESULTS) throw_rg("rg_search_batch");
        const std::string short_msg = s == RG_ERR_NOT_ENOUGH_RESULTS ? rg_last_error_string() : "";
        for (size_t i = 0; i < nb; ++i) {
            PendingQuery *q = batch[i];
            memcpy(q->indices, mb_ids_ + i * k, k * sizeof(unsigned));
            memcpy(q->dists, mb_dists_ + i * k, k * sizeof(float));
            q->cmps = mb_cmps_[i];
            q->hops = mb_hops_[i];
            // a query that ended with fewer than k pool entries has its ids filled with 0xFFFFFFFF: only ITS caller gets the
            // reference's "not enough results" exception (:2408-2412)
            if (s == RG_ERR_NOT_ENOUGH_RESULTS && k > 0 && mb_ids_[i * k + k - 1] == 0xFFFFFFFFu) {
                q->short_result = true;
                q->error = short_msg;
            }
        }
    } catch (const std::exception &ex) {
        error = ex.what();
    }
    lk.lock();
    for (PendingQuery *q : batch) {
        if (!error.empty()) q->error = error;
        q->done = true;
    }
}

std::pair<uint32_t, uint32_t> IndexBipartite::SearchRoarGraph(const float *query, size_t k, size_t &,
                                                              const Parameters &parameters, unsigned *indices,
                                                              std::vector<float> &res_dists) {
    if (res_dists.size() < k) res_dists.resize(k);
    PendingQuery me;
    me.query = query;
    me.k = k;
    me.L = parameters.Get<uint32_t>("L_pq");
    me.indices = indices;
    me.dists = res_dists.data();
    {
        std::unique_lock<std::mutex> lk(mb_mutex_);
        mb_pending_.push_back(&me);
        while (!me.done) {
            if (!mb_leader_active_) {  // nobody is collecting: lead one batch (mine, unless an older (k, L) group goes first)
                mb_leader_active_ = true;
                lead_microbatch(lk);
                mb_leader_active_ = false;
                mb_cv_.notify_all();
            } else {
                mb_cv_.wait(lk);
            }
        }
    }
    if (!me.error.empty()) throw std::runtime_error(me.error);
    return {me.cmps, me.hops};
}

// ====================================================================================================
// Graph construction (CPU).  Vocabulary: "occluded(p, R)" = p.id is in R, or some r in R (scanned in
// insertion order) has dist(p.id, r) < p.distance, where p.distance is p's distance to the list owner.
// ====================================================================================================
namespace {

struct BuildCtx {
    const float *base;
    size_t dim;
    const Distance *dist;
    uint32_t M;  // M_pjbp
    float d(uint32_t a, uint32_t b) const { return dist->compare(base + dim * (size_t)a, base + dim * (size_t)b, (unsigned)dim); }
};

inline bool contains(const std::vector<uint32_t> &v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// The occlusion walk shared by every prune routine: visit sorted[begin..end), keep a candidate if it is not
// occluded by what has been kept so far and is not the owner itself (e.g. :1635-1656).
void occlusion_pass(const BuildCtx &c, const std::vector<Neighbor> &sorted, size_t begin, size_t end, uint32_t owner,
                    std::vector<uint32_t> &kept) {
    for (size_t i = begin; i < end && kept.size() < c.M; ++i) {
        const Neighbor &p = sorted[i];
        bool occluded = false;
        for (uint32_t r : kept) {
            if (p.id == r || c.d(p.id, r) < p.distance) {
                occluded = true;
                break;
            }
        }
        if (!occluded && p.id != owner) kept.push_back(p.id);
    }
}
// Every routine then repeats the walk from index 1 with an additional "not already kept" test (:1658-1683,
// :1571-1594, :1480-1511, :1896-1926).  A candidate the first walk already judged can never be added by the
// second one: it is kept, the owner, or occluded by a member that is still in `kept`; and the second walk only
// runs when the first one exhausted the list.  So the second walk matters only for candidates the first walk
// never looked at - which exist in prune_base_search alone (the leading projection neighbours it skips).

// PruneBiSearchBaseGetBase, :1612-1694
std::vector<uint32_t> prune_projection(const BuildCtx &c, const std::vector<Neighbor> &pool, uint32_t owner) {
    std::vector<Neighbor> uniq;
    uniq.reserve(pool.size());
    for (const Neighbor &b : pool) {
        if (b.id == owner) continue;
        if (std::find_if(uniq.begin(), uniq.end(), [&](const Neighbor &u) { return u.id == b.id; }) == uniq.end())
            uniq.push_back(b);
    }
    std::vector<uint32_t> kept;
    if (uniq.empty()) return kept;
    std::sort(uniq.begin(), uniq.end());
    kept.reserve(c.M * 2);
    kept.push_back(uniq[0].id);
    occlusion_pass(c, uniq, 1, uniq.size(), owner, kept);
    for (size_t i = 1; i < uniq.size() && kept.size() < c.M; ++i)  // nearest-first fill, :1685-1691
        if (!contains(kept, uniq[i].id) && uniq[i].id != owner) kept.push_back(uniq[i].id);
    return kept;
}

// Scores `ids` against `owner`, dropping repeated ids (first occurrence wins; Neighbor::operator== is id-only)
std::vector<Neighbor> score_unique(const BuildCtx &c, uint32_t owner, const std::vector<uint32_t> &ids,
                                   size_t phantom_slots) {
    // phantom_slots reproduces `std::vector<Neighbor> prune_queue(pruned_list.size())` at :1438: the queue starts
    // with that many value-initialised entries {id 0, distance 0} which also suppress a real entry for id 0.
    std::vector<Neighbor> q(phantom_slots);
    q.reserve(phantom_slots + ids.size());
    for (uint32_t id : ids) {
        const Neighbor nn(id, c.d(owner, id), false);
        if (std::find(q.begin(), q.end(), nn) == q.end()) q.push_back(nn);
    }
    return q;
}

// PruneProjectionReverseCandidates, :1527-1610 (with_fill = true, no phantoms) and
// PruneProjectionInternalReverseCandidates, :1434-1525 (with_fill = false, phantoms)
void prune_reverse(const BuildCtx &c, uint32_t owner, std::vector<uint32_t> &list, bool internal) {
    static const bool no_phantoms = std::getenv("RG_HOST_NO_PHANTOMS") != nullptr;  // experiment switch (graph construction)
    std::vector<Neighbor> q = score_unique(c, owner, list, internal && !no_phantoms ? list.size() : 0);
    std::sort(q.begin(), q.end());
    std::vector<uint32_t> kept;
    kept.reserve(c.M * 2);
    size_t first = 0;
    if (q[first].id == owner) ++first;
    if (first < q.size()) {
        kept.push_back(q[first].id);
        occlusion_pass(c, q, first + 1, q.size(), owner, kept);
    }
    if (!internal) {  // fill from the ORIGINAL order, :1596-1600
        for (size_t i = 0; i < list.size() && kept.size() < c.M; ++i)
            if (!contains(kept, list[i])) kept.push_back(list[i]);
    }
    list = kept;
}

// PruneProjectionBaseSearchCandidates, :1846-1940.  `pool` gets sorted in place like the reference's.
std::vector<uint32_t> prune_base_search(const BuildCtx &c, std::vector<Neighbor> &pool, uint32_t owner,
                                        const std::vector<uint32_t> &owner_projection) {
    std::vector<uint32_t> kept;
    if (pool.empty()) return kept;
    std::sort(pool.begin(), pool.end());
    size_t first = 0;
    if (pool[first].id == owner) ++first;
    // skip leading candidates that are already projection neighbours (:1862-1864; the reference has no bounds
    // check here - running off the end is undefined there, we return an empty list)
    while (first < pool.size() && contains(owner_projection, pool[first].id)) ++first;
    if (first >= pool.size()) return kept;
    kept.reserve(c.M * 2);
    kept.push_back(pool[first].id);
    occlusion_pass(c, pool, first + 1, pool.size(), owner, kept);
    // second walk (:1896-1926) restarts at index 1: the skipped leading projection neighbours get their turn
    occlusion_pass(c, pool, 1, first, owner, kept);
    return kept;  // no fill pass: may stay shorter than M
}

}  // namespace

// P1 of LinkProjection for ONE training query (:1059-1084 + PruneBiSearchBaseGetBase): nn = its learn->base list (already
// cut to M_sq), element 0 the pivot.  Exposed so that the GPU build's projection prune can be checked list by list.
std::vector<uint32_t> IndexBipartite::PivotProjectionList(const float *base, size_t dim, const Distance *dist, uint32_t M_pjbp,
                                                          const uint32_t *nn, size_t n_nn) {
    if (n_nn == 0) return {};
    const BuildCtx c{base, dim, dist, M_pjbp};
    const uint32_t pivot = nn[0];
    std::vector<Neighbor> pool;
    pool.reserve(n_nn);
    for (size_t i = 0; i < n_nn; ++i)
        if (nn[i] != pivot) pool.emplace_back(nn[i], c.d(nn[i], pivot), false);
    return prune_projection(c, pool, pivot);
}

// CalculateProjectionep, :2004-2041: centroid in FP32 (row order), squared L2 to it, first minimum.
void IndexBipartite::calculate_projection_ep() {
    const size_t dim = dimension_;
    std::vector<float> center(dim, 0.f);
    for (size_t i = 0; i < nd_; ++i) {
        const float *row = data_bp_ + i * dim;
        for (size_t d = 0; d < dim; ++d) center[d] += row[d];
    }
    for (size_t d = 0; d < dim; ++d) center[d] /= (float)nd_;
    std::vector<float> dist(nd_);
#pragma omp parallel for
    for (size_t i = 0; i < nd_; ++i) {
        const float *row = data_bp_ + i * dim;
        float acc = 0;
        for (size_t j = 0; j < dim; ++j) acc += (center[j] - row[j]) * (center[j] - row[j]);
        dist[i] = acc;
    }
    uint32_t closest = 0;
    for (size_t i = 1; i < nd_; ++i)
        if (dist[i] < dist[closest]) closest = (uint32_t)i;
    projection_ep_ = closest;
    std::cout << "projection ep: " << projection_ep_ << std::endl;
}

// LinkProjection, :1043-1277
void IndexBipartite::link_projection(const Parameters &parameters) {
    const uint32_t M = parameters.Get<uint32_t>("M_pjbp");
    const uint32_t L_pjpq = parameters.Get<uint32_t>("L_pjpq");
    const uint32_t M_sq = parameters.Get<uint32_t>("M_sq");
    const uint32_t n = (uint32_t)nd_, n_sq = (uint32_t)nd_sq_;
    omp_set_num_threads((int)parameters.Get<uint32_t>("num_threads"));
    const BuildCtx c{data_bp_, dimension_, distance_, M};
    using Guard = std::lock_guard<std::mutex>;
    auto t1 = std::chrono::high_resolution_clock::now();

    // Reverse edges: for every out-neighbour `des` of `src`, make `src` a neighbour of `des`; a full list is
    // re-pruned with the newcomer appended.  ProjectionAddReverse :1391-1432 (cap M, prune_reverse with fill)
    // and SupplyAddReverse :1352-1389 (cap 2M, internal prune).
    auto add_reverse = [&](CompactGraph &g, uint32_t src, uint32_t cap, bool internal) {
        std::vector<uint32_t> out;  // snapshot under the owner's lock: other threads may be rewriting g[src]
        {
            Guard lk(locks_[src]);
            out = g[src];
        }
        for (uint32_t des : out) {
            std::vector<uint32_t> merged;
            {
                Guard lk(locks_[des]);
                std::vector<uint32_t> &dn = g[des];
                if (contains(dn, src)) continue;
                if (dn.size() < cap) {
                    dn.push_back(src);
                    continue;
                }
                merged = dn;
            }
            merged.push_back(src);
            prune_reverse(c, des, merged, internal);
            Guard lk(locks_[des]);
            g[des] = std::move(merged);
        }
    };

    // P1 :1059-1097 - every training query projects onto its nearest base point (the pivot)
#pragma omp parallel for schedule(static, 100)
    for (uint32_t sq = 0; sq < n_sq; ++sq) {
        std::vector<uint32_t> &nn = learn_base_knn_[sq];
        if (nn.size() > M_sq) nn.resize(M_sq);
        if (nn.empty()) continue;
        const uint32_t pivot = nn[0];
        std::vector<Neighbor> pool;
        pool.reserve(nn.size());
        for (uint32_t b : nn)
            if (b != pivot) pool.emplace_back(b, c.d(b, pivot), false);
        std::vector<uint32_t> pruned = prune_projection(c, pool, pivot);
        {
            Guard lk(locks_[pivot]);
            projection_graph_[pivot] = std::move(pruned);  // overwrites reverse edges collected so far (:1090)
        }
        add_reverse(projection_graph_, pivot, M, false);
        if (sq % 1000 == 0)
            std::cout << "\r" << (100.0 * sq) / n_sq << "% of projection search bipartite by base completed." << std::flush;
    }
    std::cout << std::endl;

    // P2 :1100-1104
#pragma omp parallel for schedule(static, 100)
    for (uint32_t i = 0; i < n; ++i) add_reverse(projection_graph_, i, M, false);

    // P3 :1107-1136 - lists that grew beyond M (possible only through races at >1 thread) are re-pruned
#pragma omp parallel for schedule(static, 2048)
    for (uint32_t node = 0; node < n; ++node) {
        if (projection_graph_[node].size() <= M) continue;
        std::vector<Neighbor> pool;
        std::vector<uint32_t> seen;
        for (uint32_t x : projection_graph_[node]) {
            if (contains(seen, x)) continue;
            seen.push_back(x);
            if (x != node) pool.emplace_back(x, c.d(x, node), false);
        }
        std::vector<uint32_t> pruned = prune_projection(c, pool, node);
        Guard lk(locks_[node]);
        projection_graph_[node] = std::move(pruned);
    }
    auto t2 = std::chrono::high_resolution_clock::now();
    std::cout << "Projection time: " << std::chrono::duration<double>(t2 - t1).count() << std::endl;

    // :1183-1188
    supply_nbrs_ = projection_graph_;
    t1 = std::chrono::high_resolution_clock::now();

    // P4 :1192-1220 - connectivity enhancement: beam-search every base point in the live supply graph
    // (SearchProjectionGraphInternal :1279-1350) and link it to an occlusion-pruned subset of the EXPANDED nodes.
#pragma omp parallel
    {
        std::vector<uint32_t> stamp(n, 0);  // replaces the per-node dynamic_bitset (:1195): entry == epoch <=> visited
        uint32_t epoch = 0;
        NeighborPriorityQueue pool;
        pool.reserve(L_pjpq);
        std::vector<Neighbor> expanded;
        std::vector<uint32_t> nbrs;
#pragma omp for schedule(dynamic, 2048)
        for (uint32_t node = 0; node < n; ++node) {
            if (++epoch == 0) {
                std::fill(stamp.begin(), stamp.end(), 0u);
                epoch = 1;
            }
            const float *query = data_bp_ + dimension_ * (size_t)node;
            pool.clear();
            expanded.clear();
            pool.insert(Neighbor(projection_ep_, distance_->compare(data_bp_ + dimension_ * (size_t)projection_ep_, query, (unsigned)dimension_), false));
            stamp[projection_ep_] = epoch;  // unlike SearchRoarGraph the entry point IS marked (:1309)
            while (pool.has_unexpanded_node()) {
                const Neighbor cur = pool.closest_unexpanded();
                expanded.push_back(cur);  // :1318
                {   // the live list may be rewritten by another thread's SupplyAddReverse: copy it under its lock
                    Guard lk(locks_[cur.id]);
                    nbrs = supply_nbrs_[cur.id];
                }
                for (uint32_t nbr : nbrs) {
                    if (stamp[nbr] == epoch || nbr == node) continue;  // :1327
                    stamp[nbr] = epoch;
                    pool.insert(Neighbor(nbr, distance_->compare(data_bp_ + dimension_ * (size_t)nbr, query, (unsigned)dimension_), false));
                }
            }
            expanded.erase(std::remove_if(expanded.begin(), expanded.end(), [&](const Neighbor &x) { return x.id == node; }),
                           expanded.end());  // :1203-1208
            std::vector<uint32_t> pruned = prune_base_search(c, expanded, node, projection_graph_[node]);
            {
                Guard lk(locks_[node]);
                supply_nbrs_[node] = std::move(pruned);
            }
            add_reverse(supply_nbrs_, node, 2 * M, true);
            if (node % 1000 == 0)
                std::cout << "\r" << (100.0 * node) / n << "% of projection graph base search completed." << std::flush;
        }
    }
    std::cout << "finish connectivity enhancement" << std::endl;

    // P5 :1224-1248 - supply lists longer than M are re-pruned (no fill)
#pragma omp parallel for schedule(dynamic, 2048)
    for (uint32_t node = 0; node < n; ++node) {
        if (supply_nbrs_[node].size() <= M) continue;
        std::vector<Neighbor> pool;
        std::vector<uint32_t> seen;
        for (uint32_t x : supply_nbrs_[node]) {
            if (contains(seen, x)) continue;
            seen.push_back(x);
            pool.emplace_back(x, c.d(x, node), false);
        }
        std::vector<uint32_t> pruned = prune_base_search(c, pool, node, projection_graph_[node]);
        Guard lk(locks_[node]);
        supply_nbrs_[node] = std::move(pruned);
    }

    // P6 :1251-1269 - append the supply edges that are not projection edges yet
#pragma omp parallel for schedule(dynamic, 100)
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<uint32_t> extra;
        for (uint32_t x : supply_nbrs_[i]) {
            if (extra.size() >= 2 * M) break;
            if (!contains(projection_graph_[i], x)) extra.push_back(x);
        }
        projection_graph_[i].insert(projection_graph_[i].end(), extra.begin(), extra.end());
    }
    t2 = std::chrono::high_resolution_clock::now();
    std::cout << "Connectivity enhancement time: " << std::chrono::duration<double>(t2 - t1).count() << std::endl;
}

// Parameters "gpu_build" != 0: all construction phases run on the GPU (rg_build_roargraph, csrc/rg_build.cu).  Same
// rules, applied phase-wise to all nodes at once, so - like a multi-threaded reference build - the adjacency is not
// edge-identical to the one-thread CPU build; recall parity is tested in tests/test_build_gpu.py.
void IndexBipartite::build_on_device(const Parameters &parameters) {
    const uint32_t M = parameters.Get<uint32_t>("M_pjbp");
    const uint32_t L_pjpq = parameters.Get<uint32_t>("L_pjpq");
    const uint32_t M_sq = parameters.Get<uint32_t>("M_sq");
    size_t knn_k = learn_base_knn_.empty() ? 0 : learn_base_knn_[0].size();
    for (size_t i = 0; i < nd_sq_; ++i) knn_k = std::min(knn_k, learn_base_knn_[i].size());
    if (knn_k == 0) throw std::runtime_error("learn base knn file error");
    std::vector<uint32_t> flat(nd_sq_ * knn_k);
    for (size_t i = 0; i < nd_sq_; ++i) std::copy_n(learn_base_knn_[i].begin(), knn_k, flat.begin() + i * knn_k);
    rg_graph *g = nullptr;
    std::cout << "begin link projection (GPU " << device_ << ")" << std::endl;
    if (rg_build_roargraph(data_bp_, nd_, (uint32_t)dimension_, rg_metric(metric_), flat.data(), nd_sq_, (uint32_t)knn_k, M_sq, M,
                           L_pjpq, &g, device_) != RG_OK)
        throw_rg("rg_build_roargraph");
    uint64_t n = 0, nnz = 0;
    uint32_t dmax = 0, ep = 0;
    double phases[6];
    if (rg_graph_info(g, &n, &dmax, &nnz, &ep, phases) != RG_OK) throw_rg("rg_graph_info");
    std::vector<uint64_t> offsets(n + 1);
    std::vector<uint32_t> adj(nnz);
    const rg_status s = rg_graph_download(g, offsets.data(), adj.data());
    rg_graph_destroy(g);
    if (s != RG_OK) throw_rg("rg_graph_download");
    projection_ep_ = ep;
    projection_graph_.assign(nd_, {});
    for (size_t i = 0; i < nd_; ++i) projection_graph_[i].assign(adj.begin() + offsets[i], adj.begin() + offsets[i + 1]);
    std::cout << "GPU build phases (s): ep " << phases[0] << ", projection " << phases[1] << ", reverse " << phases[2]
              << ", enhancement search " << phases[3] << ", enhancement prune " << phases[4] << ", merge " << phases[5] << std::endl;
}

// BuildRoarGraph, :143-233
void IndexBipartite::BuildRoarGraph(size_t n_sq, const float *sq_data, size_t n_bp, const float *bp_data,
                                    const Parameters &parameters) {
    std::cout << "start build bipartite index" << std::endl;
    auto s = std::chrono::high_resolution_clock::now();
    release_device();
    data_bp_ = bp_data;
    data_sq_ = sq_data;
    nd_ = n_bp;
    nd_sq_ = n_sq;
    if (learn_base_knn_.size() < n_sq) throw std::runtime_error("learn base knn file error");
    locks_ = std::vector<std::mutex>(nd_);
    if (need_normalize) {  // :176-182 - in place, like the reference
        std::cout << "normalizing base data" << std::endl;
        float *data = const_cast<float *>(data_bp_);
        for (size_t i = 0; i < nd_; ++i) normalize(data + i * dimension_, dimension_);
    }
    if (parameters.Get<uint32_t>("gpu_build", 0u) != 0) {
        build_on_device(parameters);
    } else {
        projection_graph_.assign(nd_, {});  // BipartiteProjectionReserveSpace :951-958
        supply_nbrs_.assign(nd_, {});
        calculate_projection_ep();
        std::cout << "begin link projection" << std::endl;
        link_projection(parameters);
    }
    std::cout << std::endl;
    auto e = std::chrono::high_resolution_clock::now();
    std::cout << "Build projection graph time: " << std::chrono::duration<double>(e - s).count() << std::endl;
    size_t total = 0, dmax = 0, dmin = std::numeric_limits<size_t>::max();
    for (const auto &l : projection_graph_) {
        total += l.size();
        dmax = std::max(dmax, l.size());
        dmin = std::min(dmin, l.size());
    }
    std::cout << "total degree: " << total << std::endl;
    std::cout << "Projection degree avg: " << (double)total / (double)nd_ << std::endl;
    std::cout << "Projection degree max: " << dmax << std::endl;
    std::cout << "Projection degree min: " << dmin << std::endl;
    supply_nbrs_.clear();
    supply_nbrs_.shrink_to_fit();
    has_built = true;
}

}  // namespace efanna2e
