// Drop-in for the reference's tests/test_build_roargraph.cpp: same flags and call sequence (:105-136).
#include <omp.h>

#include <chrono>
#include <iostream>
#include <string>

#include "cli_args.h"
#include "index_bipartite.h"

int main(int argc, char **argv) {
    std::string base_data_file, sampled_query_data_file, projection_index_save_file, learn_base_nn_file, data_type, dist;
    uint32_t M_sq, M_pjbp, L_pjpq, num_threads, gpu_build;
    try {
        CliArgs args(argc, argv, {{"-T", "--num_threads"}, {"-h", "--help"}});
        if (args.has("help")) {
            std::cout << "Arguments: --data_type <float> --dist <l2/ip/cosine> --base_data_path F --sampled_query_data_path F\n"
                         "  --projection_index_save_path F --learn_base_nn_path F [--M_sq 32] [--M_pjbp 32] [--L_pjpq 32] [-T threads]\n"
                         "  [--gpu_build 1]   (not in the reference: run the graph construction on the GPU)\n";
            return 0;
        }
        data_type = args.get<std::string>("data_type");
        dist = args.get<std::string>("dist");
        base_data_file = args.get<std::string>("base_data_path");
        sampled_query_data_file = args.get<std::string>("sampled_query_data_path");
        projection_index_save_file = args.get<std::string>("projection_index_save_path");
        learn_base_nn_file = args.get<std::string>("learn_base_nn_path");
        M_sq = args.get<uint32_t>("M_sq", 32);
        M_pjbp = args.get<uint32_t>("M_pjbp", 32);
        L_pjpq = args.get<uint32_t>("L_pjpq", 32);
        num_threads = args.get<uint32_t>("num_threads", (uint32_t)omp_get_num_procs());
        gpu_build = args.get<uint32_t>("gpu_build", 0);
    } catch (const std::exception &ex) {
        std::cerr << ex.what() << '\n';
        return -1;
    }
    std::cout << "sampled query: " << sampled_query_data_file << std::endl;
    uint32_t base_num, base_dim, sq_num, sq_dim;
    efanna2e::load_meta<float>(base_data_file.c_str(), base_num, base_dim);
    efanna2e::load_meta<float>(sampled_query_data_file.c_str(), sq_num, sq_dim);
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
    float *data_bp = nullptr, *data_sq = nullptr;
    efanna2e::Parameters parameters;
    efanna2e::load_data<float>(base_data_file.c_str(), base_num, base_dim, data_bp);
    efanna2e::load_data<float>(sampled_query_data_file.c_str(), sq_num, sq_dim, data_sq);
    std::cout << "Index save path: " << projection_index_save_file << std::endl;
    // NB the reference passes the unpadded base_dim here (:117) while load_data pads rows to 8 floats; every
    // supported dataset has dim % 8 == 0.  We pass the padded length so both agree.
    efanna2e::IndexBipartite index_bipartite(efanna2e::padded_dim(base_dim), base_num + sq_num, dist_metric, nullptr);
    parameters.Set<uint32_t>("M_sq", M_sq);
    parameters.Set<uint32_t>("M_pjbp", M_pjbp);
    parameters.Set<uint32_t>("L_pjpq", L_pjpq);
    parameters.Set<uint32_t>("num_threads", num_threads);
    parameters.Set<uint32_t>("gpu_build", gpu_build);
    index_bipartite.LoadLearnBaseKNN(learn_base_nn_file.c_str());
    omp_set_num_threads((int)num_threads);
    auto s = std::chrono::high_resolution_clock::now();
    index_bipartite.BuildRoarGraph(sq_num, data_sq, base_num, data_bp, parameters);
    auto e = std::chrono::high_resolution_clock::now();
    std::cout << "indexing time: " << std::chrono::duration<double>(e - s).count() << "\n";
    index_bipartite.SaveProjectionGraph(projection_index_save_file.c_str());
    std::cout << "Save index to " << projection_index_save_file << std::endl;
    return 0;
}
