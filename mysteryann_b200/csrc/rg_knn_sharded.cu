// Base-sharded exact kNN across the GPUs of one box behind the C ABI (replaces the sequential part loop + merge of
// /root/reference thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:396-448, where the parts are 20M-point slices of
// the base walked one after the other on one host).
//
// Every rank (one GPU; a process under torchrun, or a thread of the compute_groundtruth driver) holds ONE base shard and
// ALL queries.  Rank g owns the merged answer of the contiguous query slice [qb[g], qb[g+1]).  The work is cut into
// chunks; chunk c covers segment c of EVERY rank's slice, so that per chunk
//   1. K2/K3 score the G segments against the local shard (one call per segment, the FP16 copy of the shard is reused),
//   2. one grouped ncclSend/ncclRecv exchange over NVLink hands rank g the G partial lists of its segment,
//   3. K4 merges them straight into rank g's output rows.
// The exchange of a chunk is ~G*m*K*8 bytes (m = segment rows) against ~2*G*m*n_shard*dim FLOP of GEMM, i.e. milliseconds
// against seconds, so it is issued on the same stream rather than overlapped.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the process - torch's - or the system one), so
// libroargraph_b200.so has no link-time NCCL dependency and the single-GPU paths never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "rg_knn.cuh"

namespace rg {
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

static NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {  // a copy that is already loaded (torch's) wins over opening a second one
            api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (api.handle) break;
        }
        for (const char *nm : names) {
            if (api.handle) break;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!api.handle) return;
#define RG_NCCL_SYM(field, sym)                                                      \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));       \
    if (!api.field) return;
        RG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        RG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        RG_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        RG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        RG_NCCL_SYM(Send, "ncclSend")
        RG_NCCL_SYM(Recv, "ncclRecv")
        RG_NCCL_SYM(GroupStart, "ncclGroupStart")
        RG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        RG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        RG_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef RG_NCCL_SYM
        api.ok = true;
    });
    return api;
}

static rg_status need_nccl() {
    if (!nccl().ok) return fail(RG_ERR_INTERNAL, "NCCL is not available: libnccl.so.2 could not be loaded (%s)", dlerror());
    return RG_OK;
}

#define RG_NCCL_OK(expr)                                                                                       \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess)                                                                                 \
            return rg::fail(RG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, rg::nccl().GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

// contiguous ranges: part g = [b[g], b[g+1]); the first total % parts ranges get one extra element
static std::vector<uint64_t> split_bounds(uint64_t total, int parts) {
    std::vector<uint64_t> b(size_t(parts) + 1, 0);
    const uint64_t q = total / uint64_t(parts), r = total % uint64_t(parts);
    for (int g = 0; g < parts; ++g) b[size_t(g) + 1] = b[size_t(g)] + q + (uint64_t(g) < r ? 1 : 0);
    return b;
}

struct ShardedScratch {
    uint32_t *part_ids = nullptr, *recv_ids = nullptr;
    float *part_d = nullptr, *recv_d = nullptr;
    uint64_t cap = 0;  // entries per buffer
    ~ShardedScratch() { release(); }
    void release() {
        cudaFree(part_ids);
        cudaFree(recv_ids);
        cudaFree(part_d);
        cudaFree(recv_d);
        part_ids = recv_ids = nullptr;
        part_d = recv_d = nullptr;
        cap = 0;
    }
    rg_status ensure(uint64_t entries) {
        if (cap >= entries) return RG_OK;
        release();
        RG_CUDA_OK(cudaMalloc(&part_ids, entries * sizeof(uint32_t)));
        RG_CUDA_OK(cudaMalloc(&recv_ids, entries * sizeof(uint32_t)));
        RG_CUDA_OK(cudaMalloc(&part_d, entries * sizeof(float)));
        RG_CUDA_OK(cudaMalloc(&recv_d, entries * sizeof(float)));
        cap = entries;
        return RG_OK;
    }
};

}  // namespace
}  // namespace rg

extern "C" {

int rg_nccl_version(void) {
    int v = 0;
    if (!rg::nccl().ok || rg::nccl().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

rg_status rg_nccl_get_unique_id(void *id128) {
    if (!id128) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_nccl_get_unique_id: null argument");
    rg_status s = rg::need_nccl();
    if (s != RG_OK) return s;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    RG_NCCL_OK(rg::nccl().GetUniqueId(static_cast<ncclUniqueId *>(id128)));
    return RG_OK;
}

rg_status rg_nccl_comm_init_rank(void **comm, int world, int rank, const void *id128, int device) {
    if (!comm || !id128 || world <= 0 || rank < 0 || rank >= world)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_nccl_comm_init_rank: bad argument");
    rg_status s = rg::need_nccl();
    if (s != RG_OK) return s;
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    RG_NCCL_OK(rg::nccl().CommInitRank(&c, world, id, rank));
    *comm = c;
    return RG_OK;
}

rg_status rg_nccl_comm_init_all(void **comms, int ndev, const int *devices) {
    if (!comms || ndev <= 0) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_nccl_comm_init_all: bad argument");
    rg_status s = rg::need_nccl();
    if (s != RG_OK) return s;
    if (rg_device_count() < ndev) return rg::fail(RG_ERR_NO_DEVICE, "rg_nccl_comm_init_all: %d devices requested, %d present", ndev, rg_device_count());
    std::vector<ncclComm_t> c(size_t(ndev), nullptr);
    RG_NCCL_OK(rg::nccl().CommInitAll(c.data(), ndev, devices));
    for (int i = 0; i < ndev; ++i) comms[i] = c[size_t(i)];
    return RG_OK;
}

rg_status rg_nccl_comm_destroy(void *comm) {
    if (!comm) return RG_OK;
    rg_status s = rg::need_nccl();
    if (s != RG_OK) return s;
    RG_NCCL_OK(rg::nccl().CommDestroy(static_cast<ncclComm_t>(comm)));
    return RG_OK;
}

void rg_knn_sharded_slice(uint64_t nq, int rank, int world, uint64_t *lo, uint64_t *hi) {
    if (world <= 0 || rank < 0 || rank >= world) {
        if (lo) *lo = 0;
        if (hi) *hi = 0;
        return;
    }
    const std::vector<uint64_t> b = rg::split_bounds(nq, world);
    if (lo) *lo = b[size_t(rank)];
    if (hi) *hi = b[size_t(rank) + 1];
}

// `rank` / `world` describe the group of ranks that share one query set (one rank per base shard); group member p is rank
// peer_base + p of the communicator.  rg_knn_exact_sharded: the group is the whole communicator; rg_knn_exact_grid: the
// communicator holds several such groups side by side, each with its own queries.
static rg_status sharded_impl(const float *d_base_shard, uint64_t n_shard, uint64_t id_base, const float *d_queries,
                              uint64_t nq, uint32_t dim, int metric, uint32_t K, uint32_t *d_ids, float *d_dists,
                              void *nccl_comm, int rank, int world, int peer_base, int device, void *cuda_stream) {
    if (world <= 0 || rank < 0 || rank >= world) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_sharded: bad rank/world");
    if (world > 1 && !nccl_comm) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_sharded: world > 1 needs an NCCL communicator");
    if (uint64_t(world) * K > 1024) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_sharded: need world * K <= 1024");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint64_t stats[3] = {0, 0, 0}, total[3] = {0, 0, 0};
    if (world == 1) {
        rg_status s = rg::knn::knn_device(d_base_shard, n_shard, id_base, d_queries, nq, dim, metric, K, d_ids, d_dists, st, stats, false);
        rg::knn::set_last_stats(stats);
        return s;
    }
    rg_status s = rg::need_nccl();
    if (s != RG_OK) return s;
    rg::NcclApi &api = rg::nccl();
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    const std::vector<uint64_t> qb = rg::split_bounds(nq, world);
    const uint64_t mine = qb[size_t(rank) + 1] - qb[size_t(rank)];
    const uint64_t longest = qb[1] - qb[0];
    // segment rows per rank and chunk: whole K2 query batches (131072), as many as keep the four exchange buffers
    // (world * seg * K entries each) around a gigabyte in total
    const uint64_t kBatch = 131072;
    uint64_t seg = std::max<uint64_t>(kBatch, ((uint64_t(1) << 27) / (uint64_t(world) * K)) / kBatch * kBatch);
    if (const char *e = getenv("RG_KNN_SHARD_SEG")) seg = uint64_t(std::max<long long>(1, atoll(e)));  // tests: force the chunked path
    static thread_local rg::ShardedScratch scratch;

    if (longest <= seg) {
        // Everything fits one chunk: ONE K2/K3 call over all queries (full 131072-query batches instead of `world` small
        // segments), the part lists laid out like the query array so that slice p starts at row qb[p].
        if ((s = scratch.ensure(std::max<uint64_t>(nq, uint64_t(world) * longest) * K)) != RG_OK) return s;
        s = rg::knn::knn_device(d_base_shard, n_shard, id_base, d_queries, nq, dim, metric, K, scratch.part_ids, scratch.part_d, st,
                                stats, false);
        if (s != RG_OK) return s;
        RG_NCCL_OK(api.GroupStart());
        for (int p = 0; p < world; ++p) {
            const uint64_t rows_p = qb[size_t(p) + 1] - qb[size_t(p)];
            if (rows_p) {
                RG_NCCL_OK(api.Send(scratch.part_ids + qb[size_t(p)] * K, rows_p * K, ncclUint32, peer_base + p, comm, st));
                RG_NCCL_OK(api.Send(scratch.part_d + qb[size_t(p)] * K, rows_p * K, ncclFloat32, peer_base + p, comm, st));
            }
            if (mine) {
                RG_NCCL_OK(api.Recv(scratch.recv_ids + uint64_t(p) * mine * K, mine * K, ncclUint32, peer_base + p, comm, st));
                RG_NCCL_OK(api.Recv(scratch.recv_d + uint64_t(p) * mine * K, mine * K, ncclFloat32, peer_base + p, comm, st));
            }
        }
        RG_NCCL_OK(api.GroupEnd());
        if (mine) {
            s = rg_knn_merge_device(scratch.recv_ids, scratch.recv_d, uint32_t(world), mine, K, metric, d_ids, d_dists, device, st);
            if (s != RG_OK) return s;
            stats[0] += 1;
        }
        RG_CUDA_OK(cudaStreamSynchronize(st));
        rg::knn::set_last_stats(stats);
        return RG_OK;
    }
    if ((s = scratch.ensure(uint64_t(world) * seg * K)) != RG_OK) return s;

    bool first = true;
    for (uint64_t c0 = 0; c0 < longest; c0 += seg) {
        // 1. local lists of segment c of every rank's slice -> part[g] (ids already global: + id_base)
        std::vector<uint64_t> rows(size_t(world), 0);
        for (int g = 0; g < world; ++g) {
            const uint64_t len = qb[size_t(g) + 1] - qb[size_t(g)];
            rows[size_t(g)] = c0 < len ? std::min<uint64_t>(seg, len - c0) : 0;
            if (!rows[size_t(g)]) continue;
            s = rg::knn::knn_device(d_base_shard, n_shard, id_base, d_queries + (qb[size_t(g)] + c0) * dim, rows[size_t(g)], dim,
                                    metric, K, scratch.part_ids + uint64_t(g) * seg * K, scratch.part_d + uint64_t(g) * seg * K,
                                    st, stats, !first);
            if (s != RG_OK) return s;
            first = false;
            for (int i = 0; i < 3; ++i) total[i] += stats[i];
        }
        // 2. exchange: part[g] goes to rank g, the lists of MY segment come from every rank (recv[p] = shard p's lists)
        const uint64_t my_rows = rows[size_t(rank)];
        RG_NCCL_OK(api.GroupStart());
        for (int p = 0; p < world; ++p) {
            if (rows[size_t(p)]) {
                RG_NCCL_OK(api.Send(scratch.part_ids + uint64_t(p) * seg * K, rows[size_t(p)] * K, ncclUint32, peer_base + p, comm, st));
                RG_NCCL_OK(api.Send(scratch.part_d + uint64_t(p) * seg * K, rows[size_t(p)] * K, ncclFloat32, peer_base + p, comm, st));
            }
            if (my_rows) {
                RG_NCCL_OK(api.Recv(scratch.recv_ids + uint64_t(p) * my_rows * K, my_rows * K, ncclUint32, peer_base + p, comm, st));
                RG_NCCL_OK(api.Recv(scratch.recv_d + uint64_t(p) * my_rows * K, my_rows * K, ncclFloat32, peer_base + p, comm, st));
            }
        }
        RG_NCCL_OK(api.GroupEnd());
        // 3. K4: merge the `world` sorted lists of each of my queries into the output rows of this segment
        if (my_rows) {
            s = rg_knn_merge_device(scratch.recv_ids, scratch.recv_d, uint32_t(world), my_rows, K, metric, d_ids + c0 * K,
                                    d_dists + c0 * K, device, st);
            if (s != RG_OK) return s;
            total[0] += 1;
        }
    }
    RG_CUDA_OK(cudaStreamSynchronize(st));
    rg::knn::set_last_stats(total);
    return RG_OK;
}

rg_status rg_knn_exact_sharded(const float *d_base_shard, uint64_t n_shard, uint64_t id_base, const float *d_queries,
                               uint64_t nq, uint32_t dim, int metric, uint32_t K, uint32_t *d_ids, float *d_dists,
                               void *nccl_comm, int rank, int world, int device, void *cuda_stream) {
    return sharded_impl(d_base_shard, n_shard, id_base, d_queries, nq, dim, metric, K, d_ids, d_dists, nccl_comm, rank, world, 0,
                        device, cuda_stream);
}

// Grid layout: the `world` ranks of the communicator form world / base_shards query groups of base_shards ranks each.  Rank r
// holds base shard r % base_shards (of base_shards shards) and the queries of group r / base_shards - ONLY those - and ends
// with the merged lists of slice r % base_shards of its group's queries.  base_shards = world is rg_knn_exact_sharded;
// base_shards = 1 is plain query sharding (every rank holds the whole base, no exchange).  Fewer, larger shards keep K2
// in its efficient regime (967 TFLOP/s per GPU on 1.25M-row shards, 1155 on 5M-row shards) and shrink the exchange.
rg_status rg_knn_exact_grid(const float *d_base_shard, uint64_t n_shard, uint64_t id_base, const float *d_group_queries,
                            uint64_t nq_group, uint32_t dim, int metric, uint32_t K, uint32_t *d_ids, float *d_dists,
                            void *nccl_comm, int rank, int world, int base_shards, int device, void *cuda_stream) {
    if (world <= 0 || rank < 0 || rank >= world || base_shards <= 0 || world % base_shards != 0)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_grid: base_shards (%d) must divide world (%d)", base_shards, world);
    const int sub_rank = rank % base_shards;
    return sharded_impl(d_base_shard, n_shard, id_base, d_group_queries, nq_group, dim, metric, K, d_ids, d_dists, nccl_comm,
                        sub_rank, base_shards, rank - sub_rank, device, cuda_stream);
}

// Host-buffer variant (what the compute_groundtruth driver calls from one thread per GPU): uploads the shard and the
// queries, runs the sharded / grid call on a private stream, downloads this rank's slice of merged lists.  `queries` are the
// rank's group's queries (all queries when base_shards == world).
rg_status rg_knn_exact_grid_host(const float *base_shard, uint64_t n_shard, uint64_t id_base, const float *queries,
                                 uint64_t nq, uint32_t dim, int metric, uint32_t K, uint32_t *ids, float *dists,
                                 void *nccl_comm, int rank, int world, int base_shards, int device) {
    if (!base_shard || !queries || !ids || !dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_sharded_host: null argument");
    if (world <= 0 || rank < 0 || rank >= world || base_shards <= 0 || world % base_shards != 0)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact_sharded_host: bad rank/world/base_shards");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    uint64_t lo = 0, hi = 0;
    rg_knn_sharded_slice(nq, rank % base_shards, base_shards, &lo, &hi);
    struct Bufs {
        float *base = nullptr, *q = nullptr, *d = nullptr;
        uint32_t *i = nullptr;
        cudaStream_t st = nullptr;
        ~Bufs() {
            cudaFree(base);
            cudaFree(q);
            cudaFree(d);
            cudaFree(i);
            if (st) cudaStreamDestroy(st);
        }
    } b;
    RG_CUDA_OK(cudaStreamCreateWithFlags(&b.st, cudaStreamNonBlocking));
    RG_CUDA_OK(cudaMalloc(&b.base, std::max<uint64_t>(n_shard, 1) * dim * sizeof(float)));
    RG_CUDA_OK(cudaMalloc(&b.q, std::max<uint64_t>(nq, 1) * dim * sizeof(float)));
    RG_CUDA_OK(cudaMalloc(&b.i, std::max<uint64_t>(hi - lo, 1) * K * sizeof(uint32_t)));
    RG_CUDA_OK(cudaMalloc(&b.d, std::max<uint64_t>(hi - lo, 1) * K * sizeof(float)));
    RG_CUDA_OK(cudaMemcpyAsync(b.base, base_shard, n_shard * dim * sizeof(float), cudaMemcpyHostToDevice, b.st));
    RG_CUDA_OK(cudaMemcpyAsync(b.q, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice, b.st));
    rg_status s = rg_knn_exact_grid(b.base, n_shard, id_base, b.q, nq, dim, metric, K, b.i, b.d, nccl_comm, rank, world, base_shards,
                                    device, b.st);
    if (s != RG_OK) return s;
    RG_CUDA_OK(cudaMemcpyAsync(ids, b.i, (hi - lo) * K * sizeof(uint32_t), cudaMemcpyDeviceToHost, b.st));
    RG_CUDA_OK(cudaMemcpyAsync(dists, b.d, (hi - lo) * K * sizeof(float), cudaMemcpyDeviceToHost, b.st));
    RG_CUDA_OK(cudaStreamSynchronize(b.st));
    return RG_OK;
}

rg_status rg_knn_exact_sharded_host(const float *base_shard, uint64_t n_shard, uint64_t id_base, const float *queries,
                                    uint64_t nq, uint32_t dim, int metric, uint32_t K, uint32_t *ids, float *dists,
                                    void *nccl_comm, int rank, int world, int device) {
    return rg_knn_exact_grid_host(base_shard, n_shard, id_base, queries, nq, dim, metric, K, ids, dists, nccl_comm, rank, world, world,
                                  device);
}

}  // extern "C"
