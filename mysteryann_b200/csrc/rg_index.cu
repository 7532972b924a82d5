// Index upload / teardown (replaces IndexBipartite::LoadVectorData + LoadProjectionGraph on the
// device side; the file parsing itself lives in the host C++ layer, mysteryann_b200/host/).
#include <algorithm>
#include <cstring>
#include <vector>

#include "rg_index.cuh"

namespace rg {
std::string &last_error() {
    thread_local std::string e;
    return e;
}
rg_status fail(rg_status code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

// CSR -> fixed-stride rows.  One warp per node: lanes copy the list, lane 0 writes the degree.
__global__ void expand_adjacency_kernel(const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ adj,
                                        uint32_t *__restrict__ out, uint64_t n, uint32_t stride) {
    uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t node = warp; node < n; node += nwarps) {
        uint64_t b = offsets[node], e = offsets[node + 1];
        uint32_t deg = uint32_t(e - b);
        uint32_t *row = out + node * stride;
        for (uint32_t i = lane; i < stride; i += 32) {
            uint32_t v = 0xFFFFFFFFu;
            if (i == 0) v = deg;
            else if (i <= deg) v = adj[b + i - 1];
            row[i] = v;
        }
    }
}
}  // namespace rg

extern "C" {

const char *rg_last_error_string(void) { return rg::last_error().c_str(); }
const char *rg_version_string(void) { return "roargraph_b200 0.1 (sm_100a)"; }

int rg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

rg_status rg_index_create(rg_index **out, const float *base, uint64_t n, uint32_t dim, int metric,
                          const uint64_t *adj_offsets, const uint32_t *adj, uint32_t ep, int device,
                          int base_on_device) {
    if (!out || !base || !adj_offsets || (!adj && adj_offsets[n] != 0))
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: null argument");
    if (n == 0 || n >= (1ull << 31)) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: n must be in [1, 2^31)");
    if (dim == 0 || dim % 8 != 0)
        return rg::fail(RG_ERR_INVALID_ARGUMENT,
                        "rg_index_create: dim must be a non-zero multiple of 8 (pad rows like data_align, util.h:37-75)");
    if (metric != RG_METRIC_L2 && metric != RG_METRIC_INNER_PRODUCT && metric != RG_METRIC_COSINE)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: unsupported metric %d", metric);
    if (ep >= n) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: entry point %u out of range", ep);
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= rg_device_count())
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: device %d out of range", device);

    uint32_t max_deg = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (adj_offsets[i + 1] < adj_offsets[i]) return rg::fail(RG_ERR_INVALID_ARGUMENT, "adjacency offsets not monotone");
        uint64_t d = adj_offsets[i + 1] - adj_offsets[i];
        if (d > 4095) return rg::fail(RG_ERR_INVALID_ARGUMENT, "out-degree %llu of node %llu too large", (unsigned long long)d, (unsigned long long)i);
        if (d > max_deg) max_deg = uint32_t(d);
    }
    uint64_t nnz = adj_offsets[n];
    // neighbour ids index the base rows on the device (TMA gathers): a foreign or truncated index file must fail here,
    // not as an out-of-bounds copy inside K1
    for (uint64_t i = 0; i < nnz; ++i)
        if (adj[i] >= n)
            return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create: neighbour id %u (adjacency entry %llu) out of range [0, %llu)", adj[i],
                            (unsigned long long)i, (unsigned long long)n);

    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg_index *ix = new rg_index();
    ix->device = device;
    ix->n = n;
    ix->dim = dim;
    ix->metric = metric;
    ix->ep = ep;
    ix->max_degree = max_deg;
    ix->adj_stride = (max_deg + 1 + 7) / 8 * 8;
    cudaDeviceProp prop;
    auto cleanup_fail = [&](rg_status s) {
        rg_index_destroy(ix);
        return s;
    };
#define RG_TRY(expr)                                                                                               \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess)                                                                                     \
            return cleanup_fail(rg::fail(_e == cudaErrorMemoryAllocation ? RG_ERR_OUT_OF_MEMORY : RG_ERR_CUDA,     \
                                         "%s failed: %s", #expr, cudaGetErrorString(_e)));                        \
    } while (0)
    RG_TRY(cudaGetDeviceProperties(&prop, device));
    ix->sm_count = prop.multiProcessorCount;
    ix->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    RG_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));

    if (base_on_device) {
        ix->d_base = base;
        ix->owns_base = false;
    } else {
        float *d = nullptr;
        RG_TRY(cudaMalloc(&d, n * uint64_t(dim) * sizeof(float)));
        ix->d_base = d;
        ix->owns_base = true;
        RG_TRY(cudaMemcpy(d, base, n * uint64_t(dim) * sizeof(float), cudaMemcpyHostToDevice));
    }
    // adjacency: upload CSR, expand on device, free CSR
    uint64_t *d_off = nullptr;
    uint32_t *d_csr = nullptr;
    RG_TRY(cudaMalloc(&ix->d_adj, n * uint64_t(ix->adj_stride) * sizeof(uint32_t)));
    RG_TRY(cudaMalloc(&d_off, (n + 1) * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&d_csr, (nnz ? nnz : 1) * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaFree(d_off);
        return cleanup_fail(rg::fail(RG_ERR_OUT_OF_MEMORY, "cudaMalloc(adjacency) failed: %s", cudaGetErrorString(e)));
    }
    e = cudaMemcpy(d_off, adj_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(d_csr, adj, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        int blocks = (int)std::min<uint64_t>((n + 7) / 8, uint64_t(ix->sm_count) * 32);
        rg::expand_adjacency_kernel<<<blocks, 256>>>(d_off, d_csr, ix->d_adj, n, ix->adj_stride);
        ix->launches++;
        e = cudaDeviceSynchronize();
    }
    cudaFree(d_off);
    cudaFree(d_csr);
    if (e != cudaSuccess) return cleanup_fail(rg::fail(RG_ERR_CUDA, "adjacency upload failed: %s", cudaGetErrorString(e)));
    RG_TRY(cudaMalloc(&ix->d_counters, 64 * sizeof(uint32_t)));
    RG_TRY(cudaMemset(ix->d_counters, 0, 64 * sizeof(uint32_t)));
    ix->reserve_search_scratch();
#undef RG_TRY
    *out = ix;
    return RG_OK;
}

rg_status rg_index_destroy(rg_index *ix) {
    if (!ix) return RG_OK;
    rg::DeviceGuard guard(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->persist_bytes) cudaCtxResetPersistingL2Cache();  // the visited-hash slabs go away: demote their L2 lines
    if (ix->owns_base && ix->d_base) cudaFree(const_cast<float *>(ix->d_base));
    cudaFree(ix->d_adj);
    cudaFree(ix->d_counters);
    cudaFree(ix->d_overflow_list);
    cudaFree(ix->d_ghash);
    cudaFree(ix->d_queries);
    cudaFree(ix->d_ids);
    cudaFree(ix->d_dists);
    cudaFree(ix->d_cmps);
    cudaFree(ix->d_hops);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    cudaGetLastError();
    delete ix;
    return RG_OK;
}

rg_status rg_index_info(const rg_index *ix, uint64_t *n, uint32_t *dim, int *metric, uint32_t *ep,
                        uint32_t *max_degree, int *device) {
    if (!ix) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_info: null index");
    if (n) *n = ix->n;
    if (dim) *dim = ix->dim;
    if (metric) *metric = ix->metric;
    if (ep) *ep = ix->ep;
    if (max_degree) *max_degree = ix->max_degree;
    if (device) *device = ix->device;
    return RG_OK;
}

rg_status rg_host_register(void *ptr, uint64_t bytes) {
    if (!ptr || !bytes) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_host_register: null argument");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        // fine only when the WHOLE range is already page-locked and mapped contiguously (a buffer that merely shares its
        // first page with a registered neighbour must not be taken for pinned memory: rg_search_batch would write past
        // the registered page); otherwise report it so that the caller keeps the staged path
        cudaGetLastError();
        cudaPointerAttributes a0, a1;
        if (cudaPointerGetAttributes(&a0, ptr) == cudaSuccess &&
            cudaPointerGetAttributes(&a1, static_cast<char *>(ptr) + bytes - 1) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
            a1.type == cudaMemoryTypeHost &&
            static_cast<char *>(a1.devicePointer) - static_cast<char *>(a0.devicePointer) == ptrdiff_t(bytes - 1))
            return RG_OK;
        cudaGetLastError();
        return rg::fail(RG_ERR_CUDA, "cudaHostRegister(%llu bytes): the range overlaps another registration and is not fully page-locked",
                        (unsigned long long)bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return rg::fail(RG_ERR_CUDA, "cudaHostRegister(%llu bytes) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return RG_OK;
}

rg_status rg_host_unregister(void *ptr) {
    if (!ptr) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_host_unregister: null argument");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return rg::fail(RG_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
    }
    return RG_OK;
}

uint64_t rg_index_launch_count(const rg_index *ix) { return ix ? ix->launches : 0; }

rg_status rg_search_configure(rg_index *ix, int gather, int warps_per_query, int ctas_per_sm, int stage_rows,
                              int hash_log2) {
    if (!ix) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_configure: null index");
    if (gather < 0 || gather > 2) return rg::fail(RG_ERR_INVALID_ARGUMENT, "gather must be 0 (auto), 1 (cp.async) or 2 (TMA bulk)");
    if (warps_per_query < 0 || warps_per_query > 8 || ctas_per_sm < 0 || ctas_per_sm > 32)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "warps_per_query in [0,8], ctas_per_sm in [0,32]");
    if (stage_rows < 0 || (stage_rows % 4) != 0 || stage_rows > 32)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "stage_rows must be a multiple of 4 in [0,32]");
    if (hash_log2 != 0 && (hash_log2 < 8 || hash_log2 > 22))
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "hash_log2 must be 0 (auto) or in [8,22]");
    ix->cfg_gather = gather;
    ix->cfg_warps = warps_per_query;
    ix->cfg_ctas = ctas_per_sm;
    ix->cfg_stage_rows = stage_rows;
    ix->cfg_hash_log2 = hash_log2;
    return RG_OK;
}

rg_status rg_search_set_option(rg_index *ix, const char *name, int value) {
    if (!ix || !name) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_set_option: null argument");
    if (!strcmp(name, "hash_space")) {
        if (value < 0 || value > 5)
            return rg::fail(RG_ERR_INVALID_ARGUMENT, "hash_space must be 0 (auto), 1 (shared memory), 2 / 3 (global memory, atomicCAS on 32-bit keys / 16-bit quotient entries), 4 (global memory, buckets without atomics; 16-bit entries when the id range allows) or 5 (buckets of 32-bit ids)");
        ix->cfg_hash_space = value;
        return RG_OK;
    }
    if (!strcmp(name, "batch_mode")) {
        if (value < 0 || value > 3) return rg::fail(RG_ERR_INVALID_ARGUMENT, "batch_mode must be 0 (auto), 1 (per-warp gather lists), 2 (one list per query, dynamic batches) or 3 (per-warp lists, idle warps take batches of the others)");
        ix->cfg_batch_mode = value;
        return RG_OK;
    }
    if (!strcmp(name, "l2_hint")) {
        if (value < 0 || value > 3) return rg::fail(RG_ERR_INVALID_ARGUMENT, "l2_hint is a bit mask 0..3");
        ix->cfg_l2_hint = value;
        return RG_OK;
    }
    if (!strcmp(name, "adj_prefetch")) {
        if (value < 0 || value > 7) return rg::fail(RG_ERR_INVALID_ARGUMENT, "adj_prefetch is a bit mask 0..7");
        ix->cfg_adj_prefetch = value;
        return RG_OK;
    }
    if (!strcmp(name, "zero_copy")) {
        if (value < 0 || value > 1) return rg::fail(RG_ERR_INVALID_ARGUMENT, "zero_copy must be 0 (always stage) or 1 (direct when the caller buffers are page-locked)");
        ix->cfg_zero_copy = value;
        return RG_OK;
    }
    return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_set_option: unknown option '%s'", name);
}


}  // extern "C"
