// Drop-in for the DiskANN tool the reference uses to produce the learn->base kNN file and the test ground truth
// (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp): same flags (:465-486), same part-and-merge structure
// (aux_main :345-453) and the same output file (save_groundtruth_as_one_file :325-343).  The dense scoring + heap
// scan of exact_knn (:126-248, MKL sgemm on the host) is replaced by K2/K3 on the GPU (rg_knn_exact), the
// concat + std::sort merge (:424-448) by K4 (rg_knn_merge).  With several GPUs visible the base parts are dealt
// round-robin to one host thread per device (the reference walks them sequentially).
//
// Differences a caller can see: ties between equal distances are ordered by id (the reference's heap / std::sort
// leave them unspecified); K <= 128; rows are zero-padded to a multiple of 8 floats on the way to the device (does
// not change any score).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cli_args.h"
#include "roargraph_b200.h"

namespace {

constexpr uint64_t kPartSize = 20000000;  // PARTSIZE, compute_groundtruth.cpp:32
constexpr uint32_t kNoId = 0xFFFFFFFFu;

uint64_t padded(uint64_t d) { return (d + 7) / 8 * 8; }

struct BinHeader {
    uint64_t npts = 0, ndims = 0;
};

BinHeader read_header(const std::string &file) {
    std::ifstream reader;
    reader.exceptions(std::ios::failbit | std::ios::badbit);
    reader.open(file, std::ios::binary);
    int32_t h[2];
    reader.read(reinterpret_cast<char *>(h), sizeof(h));
    BinHeader out;
    out.npts = uint64_t(h[0]);
    out.ndims = uint64_t(h[1]);
    return out;
}

// load_bin_as_float (:267-305): rows [start, start + count) of a bin file as float, here zero-padded to `dpad`
template <typename T>
void load_part_as_float(const std::string &file, uint64_t start, uint64_t count, uint64_t ndims, uint64_t dpad,
                        std::vector<float> &out) {
    std::ifstream reader;
    reader.exceptions(std::ios::failbit | std::ios::badbit);
    reader.open(file, std::ios::binary);
    reader.seekg(std::streamoff(start * ndims * sizeof(T) + 2 * sizeof(uint32_t)), std::ios::beg);
    out.assign(count * dpad, 0.f);
    const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / (ndims * sizeof(T)));
    std::vector<T> buf(chunk_rows * ndims);
    for (uint64_t r0 = 0; r0 < count; r0 += chunk_rows) {
        const uint64_t rows = std::min(chunk_rows, count - r0);
        reader.read(reinterpret_cast<char *>(buf.data()), std::streamsize(rows * ndims * sizeof(T)));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < int64_t(rows); ++i)
            for (uint64_t j = 0; j < ndims; ++j) out[(r0 + uint64_t(i)) * dpad + j] = float(buf[uint64_t(i) * ndims + j]);
    }
}

// exact_knn's COSINE branch (:146-175): rows divided by their L2 norm (epsilon for a zero row), then L2 scores
void normalize_rows(std::vector<float> &x, uint64_t rows, uint64_t ndims, uint64_t dpad) {
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < int64_t(rows); ++i) {
        float *r = x.data() + uint64_t(i) * dpad;
        float sq = 0.f;
        for (uint64_t j = 0; j < ndims; ++j) sq += r[j] * r[j];
        float norm = std::sqrt(sq);
        if (norm == 0) norm = std::numeric_limits<float>::epsilon();
        for (uint64_t j = 0; j < ndims; ++j) r[j] = r[j] / norm;
    }
}

// save_groundtruth_as_one_file (:325-343)
void save_groundtruth_as_one_file(const std::string &filename, const uint32_t *ids, const float *dists, uint64_t npts,
                                  uint64_t ndims) {
    std::ofstream writer(filename, std::ios::binary | std::ios::out);
    if (!writer) throw std::runtime_error("cannot open " + filename + " for writing");
    const int32_t h[2] = {int32_t(npts), int32_t(ndims)};
    writer.write(reinterpret_cast<const char *>(h), sizeof(h));
    std::cout << "Saving truthset in one file (npts, dim, npts*dim id-matrix, npts*dim dist-matrix) with npts = " << npts
              << ", dim = " << ndims << ", size = " << 2 * npts * ndims * sizeof(unsigned) + 2 * sizeof(int) << "B" << std::endl;
    writer.write(reinterpret_cast<const char *>(ids), std::streamsize(npts * ndims * sizeof(uint32_t)));
    writer.write(reinterpret_cast<const char *>(dists), std::streamsize(npts * ndims * sizeof(float)));
    writer.close();
    std::cout << "Finished writing truthset" << std::endl;
}

void check(rg_status s, const char *what) {
    if (s != RG_OK) throw std::runtime_error(std::string(what) + ": " + rg_last_error_string());
}

template <typename T>
int aux_main(const std::string &base_file, const std::string &query_file, const std::string &gt_file, uint64_t k, int metric,
             bool cosine, const std::string &tags_file, int n_devices, uint64_t part_rows, bool explicit_parts, int base_shards) {
    const BinHeader bh = read_header(base_file), qh = read_header(query_file);
    std::cout << "Reading bin file " << base_file << " ...\n#pts = " << bh.npts << ", #dims = " << bh.ndims << std::endl;
    if (bh.ndims != qh.ndims) throw std::runtime_error("base and query dimensions differ");
    const uint64_t ndims = bh.ndims, dpad = padded(ndims);
    uint64_t nqueries = qh.npts;
    if (nqueries > kPartSize) {  // :359-363
        std::cerr << "WARNING: #Queries provided (" << nqueries << ") is greater than " << kPartSize
                  << ". Computing GT only for the first " << kPartSize << " queries." << std::endl;
        nqueries = kPartSize;
    }
    const uint64_t num_parts = (bh.npts + part_rows - 1) / part_rows;
    std::cout << "Number of parts: " << num_parts << std::endl;

    std::vector<float> queries;
    load_part_as_float<T>(query_file, 0, nqueries, ndims, dpad, queries);
    if (cosine) normalize_rows(queries, nqueries, ndims, dpad);

    std::vector<uint32_t> location_to_tag;  // :366-391
    if (!tags_file.empty()) {
        const BinHeader th = read_header(tags_file);
        if (th.ndims != 1) throw std::runtime_error("tag file error");
        if (th.npts != bh.npts) throw std::runtime_error("point num in tags file mismatch");
        std::vector<float> dummy;
        location_to_tag.resize(th.npts);
        std::ifstream reader(tags_file, std::ios::binary);
        reader.seekg(8);
        reader.read(reinterpret_cast<char *>(location_to_tag.data()), std::streamsize(th.npts * sizeof(uint32_t)));
    }

    std::cout << "Going to compute " << k << " NNs for " << nqueries << " queries over " << bh.npts << " points in " << ndims
              << " dimensions using" << (metric == RG_METRIC_INNER_PRODUCT ? " MIPS " : cosine ? " Cosine " : " L2 ")
              << "distance fn. on " << n_devices << " GPU(s)" << std::endl;

    std::vector<uint32_t> closest_points(nqueries * k);
    std::vector<float> dist_closest_points(nqueries * k);

    // Several GPUs, default part size, no tags: ONE base shard per GPU and the merge on the devices (rg_knn_exact_sharded:
    // K2/K3 per shard, grouped ncclSend/ncclRecv exchange of the per-shard lists over NVLink, K4) instead of the
    // part-by-part walk with a host-side merge below.  One host thread per GPU, communicators from ncclCommInitAll.
    // --base_shards B (default: one shard per GPU) arranges the GPUs as B base shards x devices / B query groups
    // (rg_knn_exact_grid): fewer, larger shards keep the GEMM in its efficient regime when the base fits B GPUs.
    const int B = (base_shards > 0 && n_devices % base_shards == 0) ? base_shards : n_devices;
    const int G = n_devices / B;
    const bool sharded = n_devices > 1 && location_to_tag.empty() && !explicit_parts && bh.npts >= uint64_t(B) &&
                         nqueries >= uint64_t(G) && uint64_t(B) * k <= 1024 &&
                         (bh.npts / B + 1) * dpad * sizeof(float) <= (64ull << 30);
    if (sharded) {
        std::vector<void *> comms(size_t(n_devices), nullptr);
        check(rg_nccl_comm_init_all(comms.data(), n_devices, nullptr), "rg_nccl_comm_init_all");
        std::cout << "Base sharded over " << B << " GPUs x " << G << " query group(s), per-shard lists exchanged with NCCL "
                  << rg_nccl_version() << " and merged on the devices" << std::endl;
        std::mutex err_mu;
        std::string err;
        auto worker = [&](int dev) {
            try {
                const uint64_t shard = uint64_t(dev % B), group = uint64_t(dev / B);
                const uint64_t q = bh.npts / B, r = bh.npts % B;
                const uint64_t start_id = shard * q + std::min<uint64_t>(shard, r), npoints = q + (shard < r ? 1 : 0);
                std::vector<float> base;
                load_part_as_float<T>(base_file, start_id, npoints, ndims, dpad, base);
                if (cosine) normalize_rows(base, npoints, ndims, dpad);
                uint64_t g0 = 0, g1 = 0, lo = 0, hi = 0;
                rg_knn_sharded_slice(nqueries, int(group), G, &g0, &g1);          // this group's queries
                rg_knn_sharded_slice(g1 - g0, int(shard), B, &lo, &hi);           // the slice of them this GPU ends up with
                check(rg_knn_exact_grid_host(base.data(), npoints, start_id, queries.data() + g0 * dpad, g1 - g0, uint32_t(dpad),
                                             metric, uint32_t(k), closest_points.data() + (g0 + lo) * k,
                                             dist_closest_points.data() + (g0 + lo) * k, comms[size_t(dev)], dev, n_devices, B, dev),
                      "rg_knn_exact_grid_host");
            } catch (const std::exception &ex) {
                std::lock_guard<std::mutex> lock(err_mu);
                if (err.empty()) err = ex.what();
            }
        };
        std::vector<std::thread> threads;
        for (int d = 1; d < n_devices; ++d) threads.emplace_back(worker, d);
        worker(0);
        for (auto &t : threads) t.join();
        for (void *c : comms) rg_nccl_comm_destroy(c);
        if (!err.empty()) throw std::runtime_error(err);
    }

    // per-part lists, [part][nq][k]; empty slots carry kNoId
    if (!sharded) {
    std::vector<uint32_t> part_ids(num_parts * nqueries * k, kNoId);
    std::vector<float> part_dists(num_parts * nqueries * k, 0.f);
    std::mutex err_mu;
    std::string err;
    auto worker = [&](int dev) {
        try {
            std::vector<float> base;
            for (uint64_t p = uint64_t(dev); p < num_parts; p += uint64_t(n_devices)) {
                const uint64_t start_id = p * part_rows, npoints = std::min(part_rows, bh.npts - start_id);
                load_part_as_float<T>(base_file, start_id, npoints, ndims, dpad, base);
                if (cosine) normalize_rows(base, npoints, ndims, dpad);
                uint32_t *ids = part_ids.data() + p * nqueries * k;
                float *dists = part_dists.data() + p * nqueries * k;
                check(rg_knn_exact(base.data(), npoints, start_id, queries.data(), nqueries, uint32_t(dpad), metric, uint32_t(k),
                                   ids, dists, dev),
                      "rg_knn_exact");
                if (!location_to_tag.empty())  // :409-412: points whose tag is 0 are dropped from the part's list
                    for (uint64_t i = 0; i < nqueries * k; ++i)
                        if (ids[i] != kNoId && location_to_tag[ids[i]] == 0) ids[i] = kNoId;
                std::cout << "Computed exact k-NN of part " << p << " [" << start_id << "," << start_id + npoints << ") on GPU "
                          << dev << std::endl;
            }
        } catch (const std::exception &ex) {
            std::lock_guard<std::mutex> lock(err_mu);
            if (err.empty()) err = ex.what();
        }
    };
    std::vector<std::thread> threads;
    for (int d = 1; d < n_devices; ++d) threads.emplace_back(worker, d);
    worker(0);
    for (auto &t : threads) t.join();
    if (!err.empty()) throw std::runtime_error(err);

    // merge (:424-448): K4 takes up to 1024 / k lists at a time; more parts are folded group by group
    const uint64_t group = std::max<uint64_t>(2, 1024 / k);
    uint64_t done = 0;
    bool have_acc = false;
    std::vector<uint32_t> stage_ids;
    std::vector<float> stage_d;
    while (done < num_parts) {
        const uint64_t take = std::min(num_parts - done, have_acc ? group - 1 : group);
        const uint64_t G = take + (have_acc ? 1 : 0);
        stage_ids.resize(G * nqueries * k);
        stage_d.resize(G * nqueries * k);
        uint64_t g = 0;
        if (have_acc) {
            std::copy(closest_points.begin(), closest_points.end(), stage_ids.begin());
            std::copy(dist_closest_points.begin(), dist_closest_points.end(), stage_d.begin());
            g = 1;
        }
        std::copy(part_ids.begin() + done * nqueries * k, part_ids.begin() + (done + take) * nqueries * k,
                  stage_ids.begin() + g * nqueries * k);
        std::copy(part_dists.begin() + done * nqueries * k, part_dists.begin() + (done + take) * nqueries * k,
                  stage_d.begin() + g * nqueries * k);
        check(rg_knn_merge(stage_ids.data(), stage_d.data(), uint32_t(G), nqueries, uint32_t(k), metric, closest_points.data(),
                           dist_closest_points.data(), 0),
              "rg_knn_merge");
        have_acc = true;
        done += take;
    }
    }  // !sharded
    for (uint64_t i = 0; i < nqueries; ++i) {
        bool short_list = false;
        for (uint64_t j = 0; j < k; ++j) {
            uint32_t &id = closest_points[i * k + j];
            if (id == kNoId) {
                short_list = true;
                id = 0;  // the reference leaves the slot uninitialised
            } else if (!location_to_tag.empty()) {
                id = location_to_tag[id];  // :431-433
            }
        }
        if (short_list) std::cout << "WARNING: found less than k GT entries for query " << i << std::endl;
    }
    save_groundtruth_as_one_file(gt_file, closest_points.data(), dist_closest_points.data(), nqueries, k);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    std::string data_type, dist_fn, base_file, query_file, gt_file, tags_file;
    uint64_t K = 0, part_rows = kPartSize;
    int devices = 0, base_shards = 0;
    bool explicit_parts = false;
    try {
        CliArgs args(argc, argv, {{"-h", "--help"}});
        if (args.has("help")) {
            std::cout << "Arguments:\n  --data_type <int8/uint8/float>\n  --dist_fn <l2/mips/cosine>\n  --base_file F\n  --query_file F\n"
                         "  --gt_file F\n  --K N\n  [--tags_file F]\n  [--devices N (default: all visible GPUs)] [--base_shards B (default N)] [--part_size rows]\n";
            return 0;
        }
        data_type = args.get<std::string>("data_type");
        dist_fn = args.get<std::string>("dist_fn");
        base_file = args.get<std::string>("base_file");
        query_file = args.get<std::string>("query_file");
        gt_file = args.get<std::string>("gt_file");
        K = args.get<uint64_t>("K");
        tags_file = args.get<std::string>("tags_file", std::string());
        devices = args.get<int>("devices", 0);
        base_shards = args.get<int>("base_shards", 0);
        part_rows = args.get<uint64_t>("part_size", kPartSize);
        explicit_parts = args.has("part_size");
    } catch (const std::exception &ex) {
        std::cerr << ex.what() << '\n';
        return -1;
    }
    if (data_type != "float" && data_type != "int8" && data_type != "uint8") {
        std::cout << "Unsupported type. float, int8 and uint8 types are supported." << std::endl;
        return -1;
    }
    int metric;
    bool cosine = false;
    if (dist_fn == "l2") {
        metric = RG_METRIC_L2;
    } else if (dist_fn == "mips") {
        metric = RG_METRIC_INNER_PRODUCT;
    } else if (dist_fn == "cosine") {
        metric = RG_METRIC_L2;  // "we convert cosine distance as normalized L2 distance" (:146)
        cosine = true;
    } else {
        std::cerr << "Unsupported distance function. Use l2/mips/cosine." << std::endl;
        return -1;
    }
    if (K == 0 || K > 128) {
        std::cerr << "K must be in [1, 128]" << std::endl;
        return -1;
    }
    if (part_rows == 0 || part_rows > kPartSize) part_rows = kPartSize;
    const int visible = rg_device_count();
    if (visible <= 0) {
        std::cerr << "no CUDA device available (there is no CPU fallback)" << std::endl;
        return -1;
    }
    if (devices <= 0 || devices > visible) devices = visible;
    try {
        auto t0 = std::chrono::steady_clock::now();
        int rc;
        if (data_type == "float") rc = aux_main<float>(base_file, query_file, gt_file, K, metric, cosine, tags_file, devices, part_rows, explicit_parts, base_shards);
        else if (data_type == "int8") rc = aux_main<int8_t>(base_file, query_file, gt_file, K, metric, cosine, tags_file, devices, part_rows, explicit_parts, base_shards);
        else rc = aux_main<uint8_t>(base_file, query_file, gt_file, K, metric, cosine, tags_file, devices, part_rows, explicit_parts, base_shards);
        std::cout << "Total time: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s" << std::endl;
        return rc;
    } catch (const std::exception &e) {
        std::cout << std::string(e.what()) << std::endl;
        std::cerr << "Compute GT failed." << std::endl;
        return -1;
    }
}
