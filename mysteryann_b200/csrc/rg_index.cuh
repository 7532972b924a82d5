// Device-resident index: base rows + fixed-stride adjacency + entry point, plus search scratch.
#pragma once
#include "rg_common.cuh"

// Layout in HBM (see DESIGN.md "Data layout"):
//   base : float[n][dim]                 row-major, dim % 8 == 0, rows 32-byte aligned (800 B @ D=200)
//   adj  : uint32[n][adj_stride]         word 0 = out-degree, words 1..deg = neighbour ids, rest undefined;
//                                        adj_stride = round_up(max_degree + 1, 8) so every row is a whole
//                                        number of 32-byte sectors and one dependent read fetches a list
struct rg_index {
    int device = 0;
    uint64_t n = 0;
    uint32_t dim = 0;
    int metric = RG_METRIC_INNER_PRODUCT;
    uint32_t ep = 0;
    uint32_t max_degree = 0;
    uint32_t adj_stride = 0;
    const float *d_base = nullptr;
    bool owns_base = false;
    uint32_t *d_adj = nullptr;
    int sm_count = 0;
    int max_smem_optin = 0;

    // search scratch (device)
    uint32_t *d_counters = nullptr;       // see rg_search.cu: kCounter*
    uint32_t *d_overflow_list = nullptr;  // query ids that overflowed the shared-memory visited set
    uint64_t overflow_cap = 0;
    uint32_t *d_ghash = nullptr;  // global-memory visited-hash slabs, one per resident CTA (L2-resident at the usual L)
    uint64_t ghash_words = 0;

    // staging for the host-buffer API
    float *d_queries = nullptr;
    uint64_t queries_cap = 0;  // floats
    uint32_t *d_ids = nullptr;
    float *d_dists = nullptr;
    uint64_t res_cap = 0;  // nq*k entries
    uint32_t *d_cmps = nullptr, *d_hops = nullptr;
    uint64_t stat_cap = 0;  // nq entries
    cudaStream_t stream = nullptr;  // private stream of the host-buffer API

    // tuning (0 = auto)
    int cfg_gather = 0, cfg_warps = 0, cfg_ctas = 0, cfg_stage_rows = 0, cfg_hash_log2 = 0, cfg_hash_space = 0;
    int cfg_l2_hint = 3, cfg_adj_prefetch = 3;  // see SearchParams::l2_hint / adj_prefetch; measured best on B200
                                                // (profiles/r01_k1_variants_10m.txt)

    int cfg_batch_mode = 0;  // 0 auto, 1 every warp gathers the unvisited neighbours it filtered, 2 one CTA-wide list, batches pulled dynamically
    int cfg_zero_copy = 1;  // rg_search_batch: search straight out of / into page-locked caller buffers

    // Visited-set scratch is taken once, large enough for every beam width of the reference's sweep (210 MB at L_pq = 500 on
    // 148 SMs): growing it inside a search call costs a cudaFree + cudaMalloc in the middle of an L_pq sweep (measured: the
    // single timed call per L_pq of the drop-in driver lost 2 ms at every growth step).  Best effort - K1 sizes it exactly
    // if this allocation is not there.
    void reserve_search_scratch() {
        const uint64_t bytes = 256ull << 20;
        if (!d_ghash && cudaMalloc(&d_ghash, bytes) == cudaSuccess) ghash_words = bytes / 4;
        else (void)cudaGetLastError();
    }

    uint64_t persist_bytes = 0;  // persisting-L2 set-aside requested so far (l2_hint bit 1)
    uint64_t launches = 0;
};
