// Drop-in for the reference's tests/test_search_roargraph.cpp: same flags, same stdout table and CSV columns
// (L_pq,qps,avg_cmps,mean_latency_ms,recall,avg_hops), same load sequence (:118-173).  The OpenMP loop over
// queries (:203-209) is replaced by ONE batched GPU call per L_pq; timing includes host<->device copies.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "cli_args.h"
#include "index_bipartite.h"
#include "roargraph_b200.h"

template <typename T>
struct PageAligned {  // zero-initialised, 4096-byte aligned array with the vector calls the driver uses
    T *p = nullptr;
    size_t n = 0;
    explicit PageAligned(size_t count) : n(count) {
        if (posix_memalign(reinterpret_cast<void **>(&p), 4096, std::max<size_t>(count, 1) * sizeof(T)) != 0) throw std::bad_alloc();
        memset(p, 0, std::max<size_t>(count, 1) * sizeof(T));
    }
    ~PageAligned() { free(p); }
    PageAligned(const PageAligned &) = delete;
    PageAligned &operator=(const PageAligned &) = delete;
    T *data() { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
};

// tests/test_search_roargraph.cpp:23-36
static float ComputeRecall(uint32_t q_num, uint32_t k, uint32_t gt_dim, const uint32_t *res, const uint32_t *gt) {
    uint32_t hit = 0;
    for (uint32_t i = 0; i < q_num; i++) {
        const uint32_t *g = gt + (size_t)i * gt_dim, *r = res + (size_t)i * k;
        for (uint32_t a = 0; a < k; ++a)
            if (std::find(r, r + k, g[a]) != r + k) ++hit;
    }
    return static_cast<float>(hit) / (float)(k * q_num);
}

int main(int argc, char **argv) {
    std::string base_data_file, query_file, gt_file, projection_index_save_file, data_type, dist, evaluation_save_path;
    std::vector<uint32_t> L_vec;
    uint32_t num_threads, k;
    int device, devices;
    try {
        CliArgs args(argc, argv, {{"-T", "--num_threads"}, {"-h", "--help"}});
        if (args.has("help")) {
            std::cout << "Arguments: --data_type <float> --dist <l2/ip/cosine> --base_data_path F --query_path F --gt_path F\n"
                         "  --projection_index_save_path F --L_pq <L...> --k K [--evaluation_save_path F] [-T threads] [--device D] [--devices N]\n";
            return 0;
        }
        data_type = args.get<std::string>("data_type");
        dist = args.get<std::string>("dist");
        base_data_file = args.get<std::string>("base_data_path");
        query_file = args.get<std::string>("query_path");
        gt_file = args.get<std::string>("gt_path");
        projection_index_save_file = args.get<std::string>("projection_index_save_path");
        L_vec = args.get_all<uint32_t>("L_pq");
        k = args.get<uint32_t>("k", 1);
        evaluation_save_path = args.get<std::string>("evaluation_save_path", "");
        num_threads = args.get<uint32_t>("num_threads", (uint32_t)omp_get_num_procs());
        device = args.get<int>("device", 0);
        devices = args.get<int>("devices", 1);  // replicate the index on N GPUs and shard the queries
    } catch (const std::exception &ex) {
        std::cerr << ex.what() << '\n';
        return -1;
    }
    uint32_t base_num, base_dim;
    efanna2e::load_meta<float>(base_data_file.c_str(), base_num, base_dim);
    efanna2e::Parameters parameters;
    parameters.Set<uint32_t>("num_threads", num_threads);
    uint32_t q_pts, q_dim;
    efanna2e::load_meta<float>(query_file.c_str(), q_pts, q_dim);
    float *query_data = nullptr;
    efanna2e::load_data<float>(query_file.c_str(), q_pts, q_dim, query_data);
    // load_data already returns rows padded to 8 floats, so the reference's follow-up data_align(query_data, q_pts, q_dim)
    // (tests/test_search_roargraph.cpp:166) would re-read the padded buffer with the unpadded stride and scramble every row
    // after the first when dim % 8 != 0 (the reference has that latent bug for base and queries alike; all its datasets have
    // dim % 8 == 0).  The loaded buffer IS the aligned one; only the row length changes.
    float *aligned_query_data = query_data;
    q_dim = (uint32_t)efanna2e::padded_dim(q_dim);
    std::cout << "new_dim: " << q_dim << std::endl;

    uint32_t gt_pts, gt_dim;
    uint32_t *gt_ids = nullptr;
    float *gt_dists = nullptr;
    efanna2e::load_gt_meta<uint32_t>(gt_file.c_str(), gt_pts, gt_dim);
    efanna2e::load_gt_data_with_dist<uint32_t, float>(gt_file.c_str(), gt_pts, gt_dim, gt_ids, gt_dists);
    efanna2e::Metric dist_metric = efanna2e::INNER_PRODUCT;
    if (dist == "l2") {
        dist_metric = efanna2e::L2;
        std::cout << "Using l2 as distance metric" << std::endl;
    } else if (dist == "ip") {
        std::cout << "Using inner product as distance metric" << std::endl;
    } else if (dist == "cosine") {
        dist_metric = efanna2e::COSINE;
        std::cout << "Using cosine as distance metric" << std::endl;
    } else {
        std::cout << "Unknown distance type: " << dist << std::endl;
        return -1;
    }
    if (!std::filesystem::exists(projection_index_save_file.c_str())) {
        std::cout << "projection index file does not exist." << std::endl;
        return -1;
    }
    efanna2e::IndexBipartite index(q_dim, base_num, dist_metric, nullptr);
    index.SetDevice(device);
    index.SetDeviceCount(devices);
    index.LoadSearchNeededData(base_data_file.c_str(), "");
    std::cout << "Load graph index: " << projection_index_save_file << std::endl;
    index.LoadProjectionGraph(projection_index_save_file.c_str());
    if (index.need_normalize) {
        std::cout << "Normalizing query data" << std::endl;
        for (uint32_t i = 0; i < q_pts; i++) efanna2e::normalize<float>(aligned_query_data + (size_t)i * q_dim, q_dim);
    }
    index.InitVisitedListPool(num_threads);  // uploads base + graph to the GPU

    std::cout << "k: " << k << std::endl;
    // result arrays on their own pages: two small heap vectors sharing a page cannot both be page-locked
    PageAligned<uint32_t> res((size_t)q_pts * k), cmps(q_pts), hops(q_pts);
    PageAligned<float> res_dists((size_t)q_pts * k);
    // page-lock the query and result arrays: rg_search_batch then reads/writes them in place (no staging copies)
    const std::pair<void *, size_t> pinned[] = {{aligned_query_data, (size_t)q_pts * q_dim * sizeof(float)},
                                                {res.data(), res.size() * sizeof(uint32_t)},
                                                {res_dists.data(), res_dists.size() * sizeof(float)},
                                                {cmps.data(), cmps.size() * sizeof(uint32_t)},
                                                {hops.data(), hops.size() * sizeof(uint32_t)}};
    for (const auto &b : pinned)
        if (b.second && rg_host_register(b.first, b.second) != RG_OK)
            std::cout << "note: buffer not page-locked (" << rg_last_error_string() << "), results are staged" << std::endl;
    std::ofstream evaluation_out;
    if (!evaluation_save_path.empty()) evaluation_out.open(evaluation_save_path, std::ios::out);
    std::cout << "Using GPU device: " << device << (devices > 1 ? " .. " + std::to_string(device + devices - 1) + " (index replicated, queries sharded)" : std::string()) << std::endl;
    std::cout << "L_pq" << "\t\tQPS" << "\t\t\tavg_visited" << "\tmean_latency" << "\trecall@" << k << "\tavg_hops" << std::endl;
    for (uint32_t L_pq : L_vec) {
        if (k > L_pq) {
            std::cout << "L_pq must greater or equal than k" << std::endl;
            exit(1);
        }
        parameters.Set<uint32_t>("L_pq", L_pq);
        // warm-up like the reference's 100 sequential queries (:198-200)
        index.SearchRoarGraphBatch(aligned_query_data, std::min<uint32_t>(100, q_pts), k, parameters, res.data(),
                                   res_dists.data(), cmps.data(), hops.data());
        auto start = std::chrono::high_resolution_clock::now();
        index.SearchRoarGraphBatch(aligned_query_data, q_pts, k, parameters, res.data(), res_dists.data(), cmps.data(),
                                   hops.data());
        auto end = std::chrono::high_resolution_clock::now();
        const double us = (double)std::chrono::duration_cast<std::chrono::microseconds>(end - start).count();
        const float qps = (float)(q_pts / (us * 1e-6));
        const float recall = ComputeRecall(q_pts, k, gt_dim, res.data(), gt_ids);
        double avg_cmps = 0, avg_hops = 0;
        for (uint32_t i = 0; i < q_pts; ++i) {
            avg_cmps += cmps[i];
            avg_hops += hops[i];
        }
        avg_cmps /= q_pts;
        avg_hops /= q_pts;
        const float mean_latency_ms = (float)(us * 1e-3 / q_pts);
        std::cout << L_pq << "\t\t" << qps << "\t\t" << avg_cmps << "\t\t" << mean_latency_ms << "\t\t" << recall << "\t\t"
                  << avg_hops << std::endl;
        if (evaluation_out.is_open())
            evaluation_out << L_pq << "," << qps << "," << avg_cmps << "," << mean_latency_ms << "," << recall << ","
                           << avg_hops << std::endl;
    }
    for (const auto &b : pinned)
        if (b.second) rg_host_unregister(b.first);
    free(aligned_query_data);
    delete[] gt_ids;
    delete[] gt_dists;
    return 0;
}
