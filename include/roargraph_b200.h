/*
 * roargraph_b200.h - C ABI of the B200-native RoarGraph hot path (libroargraph_b200.so).
 *
 * The reference (matchyc/mysteryann) has no FFI: its boundary is the C++ class
 * efanna2e::IndexBipartite (include/index_bipartite.h:23-171) called from the CLI drivers
 * tests/test_search_roargraph.cpp and tests/test_build_roargraph.cpp, plus the DiskANN tool
 * thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp that produces the learn->base kNN file.
 * This header is the thin C layer the host C++ (mysteryann_b200/host/, same class and method
 * names as the reference) calls into; every entry point cites the reference code it replaces.
 * Paths below are relative to the reference repository root.
 *
 * Conventions: plain C types only; all sizes in elements unless stated; "host" pointers are
 * ordinary (pageable or pinned) CPU memory, "device" pointers are CUDA device memory on the
 * index's device; every function returns an rg_status and records a message retrievable with
 * rg_last_error_string() (thread local).  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with RG_ERR_NO_DEVICE.
 */
#ifndef ROARGRAPH_B200_H
#define ROARGRAPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_API __attribute__((visibility("default")))

typedef int rg_status;
enum {
    RG_OK = 0,
    RG_ERR_INVALID_ARGUMENT = 1,
    RG_ERR_CUDA = 2,
    RG_ERR_NOT_ENOUGH_RESULTS = 3, /* src/index_bipartite.cpp:2408-2412 "not enough results: .., expected: .." */
    RG_ERR_OUT_OF_MEMORY = 4,
    RG_ERR_NO_DEVICE = 5,
    RG_ERR_INTERNAL = 6,
    RG_ERR_IO = 7
};

/* efanna2e::Metric, include/efanna2e/distance.h:15.  COSINE searches with the inner-product
 * distance on L2-normalised rows exactly like the reference (src/index.cpp:14-17;
 * normalisation is done by the caller: src/index_bipartite.cpp:2675-2680, tests/test_search_roargraph.cpp:167-172). */
enum { RG_METRIC_L2 = 0, RG_METRIC_INNER_PRODUCT = 1, RG_METRIC_COSINE = 4 };

typedef struct rg_index rg_index; /* device-resident base vectors + projection graph + entry point */

RG_API const char *rg_last_error_string(void);
RG_API const char *rg_version_string(void);
RG_API int rg_device_count(void);

/* ---- index ------------------------------------------------------------------------------------
 * Replaces IndexBipartite::LoadVectorData + LoadProjectionGraph (src/index_bipartite.cpp:2661-2692,
 * 2097-2117): uploads the base rows and the adjacency (CSR: adj_offsets[n+1], adj[adj_offsets[n]])
 * to `device`.  `dim` is the padded row length (multiple of 8, include/efanna2e/util.h:37-75).
 * base may be a host pointer (copied) or, with base_on_device != 0, a device pointer that is
 * adopted WITHOUT copying and must outlive the index.  ep = projection_ep_. */
RG_API rg_status rg_index_create(rg_index **out, const float *base, uint64_t n, uint32_t dim, int metric,
                                 const uint64_t *adj_offsets, const uint32_t *adj, uint32_t ep, int device,
                                 int base_on_device);
RG_API rg_status rg_index_destroy(rg_index *index);
RG_API rg_status rg_index_info(const rg_index *index, uint64_t *n, uint32_t *dim, int *metric, uint32_t *ep,
                               uint32_t *max_degree, int *device);

/* ---- search -----------------------------------------------------------------------------------
 * Replaces the timed loop of tests/test_search_roargraph.cpp:203-209, i.e. nq calls of
 * IndexBipartite::SearchRoarGraph (src/index_bipartite.cpp:2311-2420) with L_pq = L:
 * ids[nq*k], dists[nq*k] (IP: negated dot; L2: squared distance), cmps[nq], hops[nq]
 * (cmps/hops may be NULL).  Results are bit-identical to the reference's.
 * Host variant: queries/results are host buffers; H2D, kernels and D2H happen inside the call.  When every buffer
 * passed is page-locked host memory (cudaHostAlloc / cudaHostRegister) the kernel reads the queries and writes
 * the results through the mapped pointers (no staging copies); pageable buffers are staged ("zero_copy" option).
 * Returns RG_ERR_NOT_ENOUGH_RESULTS if any query ends with fewer than k pool entries (its ids
 * are filled with 0xFFFFFFFF), like the reference's std::runtime_error.
 * Threading: an rg_index owns one stream and one set of scratch buffers, so calls on the SAME index must not overlap
 * (the host class serialises them); different indices - e.g. one replica per GPU - may be searched concurrently. */
RG_API rg_status rg_search_batch(rg_index *index, const float *queries, uint64_t nq, uint32_t k, uint32_t L,
                                 uint32_t *ids, float *dists, uint32_t *cmps, uint32_t *hops);
/* Device variant: all buffers are device memory, work is enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = default stream) and the call returns without synchronising.  d_status (may be NULL)
 * receives 2 x u32: {#queries with fewer than k results, #queries whose visited set overflowed
 * every fallback (must be 0)}. */
RG_API rg_status rg_search_batch_device(rg_index *index, const float *d_queries, uint64_t nq, uint32_t k,
                                        uint32_t L, uint32_t *d_ids, float *d_dists, uint32_t *d_cmps,
                                        uint32_t *d_hops, uint32_t *d_status, void *cuda_stream);
/* Build-time beam searches (replaces IndexBipartite::SearchProjectionGraphInternal, src/index_bipartite.cpp:1279-1350, the
 * inner loop of the connectivity enhancement :1192-1220): for every base row t in [node_lo, node_lo + count) a beam search of
 * width L over the index's graph with row t as the query - the entry point is scored and marked visited, neighbour t itself is
 * never scored - recording the EXPANDED nodes in expansion order (the reference's full_retset).  d_exp_keys[count][exp_cap]
 * receives (monotone image of the FP32 distance) << 32 | id << 1, d_exp_cnt[count] min(#expanded, exp_cap).  Device buffers,
 * enqueued on cuda_stream.  rg_build_roargraph_device runs these in waves; the entry is exported for parity tests. */
RG_API rg_status rg_search_expanded_device(rg_index *index, uint32_t node_lo, uint64_t count, uint32_t L,
                                           uint64_t *d_exp_keys, uint32_t *d_exp_cnt, uint32_t exp_cap, void *cuda_stream);
/* Tuning knobs (0 = automatic): see DESIGN.md "K1".  gather: 1 cp.async (LDGSTS), 2 TMA bulk copy (cp.async.bulk +
 * mbarrier); warps_per_query: warps of the CTA that owns a query (1..8); stage_rows: rows per warp staging buffer. */
RG_API rg_status rg_search_configure(rg_index *index, int gather, int warps_per_query, int ctas_per_sm, int stage_rows,
                                     int hash_log2);
/* Named options: "hash_space" = visited set of a query: 0 auto (= 4), 1 hash table in shared memory, 2 / 3 a slab per CTA in
 * global memory probed with atomicCAS (32-bit keys / 16-bit quotient entries where the id range allows), 4 a slab of
 * buckets per CTA in global memory without atomics (each warp owns a bucket range; 16-bit quotient entries where the id
 * range allows, else 32-bit ids), 5 the same with 32-bit ids always;
 * "l2_hint" bit mask (default 3): 1 = gathered base rows are loaded evict_first, 2 = the visited-hash slabs are pinned in
 * the persisting part of L2 (access-policy window; raises the device's persisting-L2 limit); "adj_prefetch" bit mask
 * (default 3): 1 = read the adjacency row of the next unexpanded pool entry ahead and prefetch the visited-hash slots of its
 * neighbours into L2, 2 = L2-prefetch the adjacency rows of scored candidates that beat it, 4 (bucketed visited set, off by
 * default: measured no gain) = when that prediction holds, issue the next hop's visited filter and first gather before the merge;
 * "batch_mode": 0 auto, 1 every warp gathers the unvisited neighbours it filtered itself, 2 the hop's unvisited neighbours go to
 * one list per query and the warps pull batches of stage_rows rows from it, 3 per-warp lists handed out in batches, a warp
 * that has emptied its own takes batches of the others';
 * "zero_copy" (default 1): rg_search_batch works straight on page-locked caller buffers, 0 = always stage through HBM. */
RG_API rg_status rg_search_set_option(rg_index *index, const char *name, int value);
/* Page-lock (and map) a caller-owned host buffer so that rg_search_batch can work on it without staging copies - what
 * the drop-in driver does with the query array and result vectors it allocates (tests/test_search_roargraph.cpp:
 * 166-179 in the reference).  Thin wrappers over cudaHostRegister / cudaHostUnregister. */
RG_API rg_status rg_host_register(void *ptr, uint64_t bytes);
RG_API rg_status rg_host_unregister(void *ptr);
/* Diagnostics (synchronises the device): queries of the last batch that were redone by the big-table visited-set pass. */
RG_API uint32_t rg_search_last_overflow_count(rg_index *index);
/* Diagnostics (synchronises the device): ids of the last batch that went to a warp's shared-memory exception list because
 * their probe window in the bucketed visited set was full (exact either way; see DESIGN.md "K1"). */
RG_API uint32_t rg_search_last_exception_count(rg_index *index);
/* Number of kernel launches issued by this library on behalf of `index` so far. */
RG_API uint64_t rg_index_launch_count(const rg_index *index);

/* ---- exact kNN (build time) -------------------------------------------------------------------
 * Replaces exact_knn + the per-part merge of compute_groundtruth.cpp:126-248, 396-448 for one base
 * shard: for each of nq queries the K base rows with the smallest score (-<p,q> for
 * RG_METRIC_INNER_PRODUCT, squared L2 otherwise), ascending by (score, id); ids are
 * shard-local row + id_base; dists are written as +<p,q> for inner product
 * (compute_groundtruth.cpp:438-441).  Host buffers. */
RG_API rg_status rg_knn_exact(const float *base, uint64_t n, uint64_t id_base, const float *queries, uint64_t nq,
                              uint32_t dim, int metric, uint32_t K, uint32_t *ids, float *dists, int device);
/* Device variant: all buffers are device memory on `device`; the kernels are enqueued on cuda_stream and the call returns
 * after synchronising that stream (the operand scale factors are read back at the start and the number of queries
 * without a completeness certificate at the end; nothing synchronises per query batch).  Scratch (FP16 operand copies,
 * candidate lists: ~5 GB for a 10M x 200 shard) is cached per device between calls: rg_knn_release_scratch(). */
RG_API rg_status rg_knn_exact_device(const float *d_base, uint64_t n, uint64_t id_base, const float *d_queries,
                                     uint64_t nq, uint32_t dim, int metric, uint32_t K, uint32_t *d_ids,
                                     float *d_dists, int device, void *cuda_stream);
/* K4: merges G per-shard lists (each nq x K, ascending by (score,id), in the OUTPUT convention
 * above) into the global top-K; replaces the concat + std::sort of compute_groundtruth.cpp:424-448.
 * Device buffers: d_part_ids/d_part_dists are [G][nq][K]. */
RG_API rg_status rg_knn_merge_device(const uint32_t *d_part_ids, const float *d_part_dists, uint32_t G,
                                     uint64_t nq, uint32_t K, int metric, uint32_t *d_ids, float *d_dists,
                                     int device, void *cuda_stream);

/* Host-buffer variant of the K4 merge (H2D, kernel, D2H inside the call); part_ids/part_dists are [G][nq][K].  An id of
 * 0xFFFFFFFF marks an empty slot (a part with fewer than K rows, or an entry dropped by the caller) and is skipped. */
RG_API rg_status rg_knn_merge(const uint32_t *part_ids, const float *part_dists, uint32_t G, uint64_t nq, uint32_t K,
                              int metric, uint32_t *ids, float *dists, int device);

/* Diagnostics of the last rg_knn_exact* call of the calling thread: kernel launches issued, and how many queries
 * failed the completeness certificate twice and were redone by the exact FP32 scan. */
RG_API void rg_knn_last_stats(uint64_t *launches, uint64_t *exact_scans);
/* ... and how many queries failed it under the optimistic threshold schedule and were re-run with the conservative one
 * (DESIGN.md "K2": expected to be a handful per million on rows in arbitrary order). */
RG_API uint64_t rg_knn_last_second_pass_count(void);
/* Frees the per-device scratch kept by rg_knn_exact_device (device < 0: every device). */
RG_API rg_status rg_knn_release_scratch(int device);

/* ---- exact kNN, base sharded over the GPUs of one box (build time) ------------------------------
 * Replaces the part loop + merge of compute_groundtruth.cpp:396-448 (there: 20M-point parts walked sequentially on one
 * host, per-part top-k concatenated and re-sorted).  `world` ranks - one per GPU; processes under torchrun or threads of
 * one process - each hold one base shard (d_base_shard, n_shard rows, global id of row 0 = id_base) and ALL nq queries
 * in device memory.  Rank r ends up with the merged global top-K of the contiguous query slice rg_knn_sharded_slice(nq,
 * r, world) in d_ids / d_dists ([slice rows][K], same conventions as rg_knn_exact).  Per chunk of queries: K2/K3 on the
 * local shard, ONE grouped ncclSend/ncclRecv exchange of the per-shard lists over NVLink (rank r receives the lists of
 * its slice), K4 merge on the device.  nccl_comm is an ncclComm_t of exactly `world` ranks in which this caller is
 * `rank` (NULL only when world == 1); every rank of the communicator must make the call with the same nq, dim, K and
 * metric.  NCCL is resolved at run time from the libnccl.so.2 already loaded in the process (e.g. torch's) or the system
 * one; the helpers below create a communicator without the caller linking NCCL itself. */
RG_API rg_status rg_knn_exact_sharded(const float *d_base_shard, uint64_t n_shard, uint64_t id_base,
                                      const float *d_queries, uint64_t nq, uint32_t dim, int metric, uint32_t K,
                                      uint32_t *d_ids, float *d_dists, void *nccl_comm, int rank, int world,
                                      int device, void *cuda_stream);
/* Grid layout over the same communicator: the `world` ranks form world / base_shards query groups of base_shards ranks each.
 * Rank r holds base shard r % base_shards (of base_shards shards of the base) and the nq_group queries of group
 * r / base_shards - only those - and ends up with the merged top-K of slice rg_knn_sharded_slice(nq_group, r % base_shards,
 * base_shards) of ITS GROUP's queries; the exchange runs between the ranks of one group only.  base_shards == world is the
 * call above, base_shards == 1 is plain query sharding (whole base on every GPU, no exchange).  The base needs as many
 * shards as it takes to fit HBM; beyond that, fewer and larger shards keep K2 in its efficient regime (DESIGN.md "K2"). */
RG_API rg_status rg_knn_exact_grid(const float *d_base_shard, uint64_t n_shard, uint64_t id_base,
                                   const float *d_group_queries, uint64_t nq_group, uint32_t dim, int metric, uint32_t K,
                                   uint32_t *d_ids, float *d_dists, void *nccl_comm, int rank, int world, int base_shards,
                                   int device, void *cuda_stream);
/* Host-buffer variant (H2D of the shard and the queries, the call above on a private stream, D2H of the slice): ids / dists
 * are host arrays of (hi - lo) x K entries.  Used by the compute_groundtruth driver, one host thread per GPU. */
RG_API rg_status rg_knn_exact_sharded_host(const float *base_shard, uint64_t n_shard, uint64_t id_base,
                                           const float *queries, uint64_t nq, uint32_t dim, int metric, uint32_t K,
                                           uint32_t *ids, float *dists, void *nccl_comm, int rank, int world,
                                           int device);
/* ... and of rg_knn_exact_grid: `queries` are the nq queries of the caller's group, ids / dists receive slice
 * rg_knn_sharded_slice(nq, rank % base_shards, base_shards) of them (compute_groundtruth --devices N --base_shards B). */
RG_API rg_status rg_knn_exact_grid_host(const float *base_shard, uint64_t n_shard, uint64_t id_base, const float *group_queries,
                                        uint64_t nq_group, uint32_t dim, int metric, uint32_t K, uint32_t *ids, float *dists,
                                        void *nccl_comm, int rank, int world, int base_shards, int device);
/* Query rows [*lo, *hi) whose merged lists rank `rank` of `world` receives (contiguous, ascending with the rank; the first
 * nq % world slices hold one extra row). */
RG_API void rg_knn_sharded_slice(uint64_t nq, int rank, int world, uint64_t *lo, uint64_t *hi);
/* ncclGetUniqueId / ncclCommInitRank / ncclCommInitAll / ncclCommDestroy through the run-time binding: id128 is the
 * 128-byte ncclUniqueId (created on one rank, handed to the others by whatever side channel the host program has);
 * rg_nccl_comm_init_all builds the ndev communicators of a one-process, thread-per-GPU program (devices may be NULL =
 * 0..ndev-1).  rg_nccl_version() returns the NCCL version code in use (0 when NCCL cannot be loaded). */
RG_API rg_status rg_nccl_get_unique_id(void *id128);
RG_API rg_status rg_nccl_comm_init_rank(void **comm, int world, int rank, const void *id128, int device);
RG_API rg_status rg_nccl_comm_init_all(void **comms, int ndev, const int *devices);
RG_API rg_status rg_nccl_comm_destroy(void *comm);
RG_API int rg_nccl_version(void);

/* ---- graph construction on the GPU (build time) ---------------------------------------------------
 * Replaces the CPU phases of IndexBipartite::BuildRoarGraph (src/index_bipartite.cpp:143-233) and LinkProjection
 * (:1043-1277): entry point (:2004-2041), pivot projection + PruneBiSearchBaseGetBase (:1059-1097, 1612-1694),
 * reverse edges (:1391-1432, 1527-1610), connectivity enhancement = L_pjpq beam searches from every base point
 * (:1192-1220, 1279-1350) + PruneProjectionBaseSearchCandidates (:1846-1940) + supply reverse edges (:1352-1389),
 * degree check (:1224-1248) and the final merge (:1251-1269).  Same pruning rules, applied per phase to all
 * nodes at once (waves for the searches); like a multi-threaded reference build the adjacency is not
 * edge-identical to the one-thread build.  d_knn_ids: the learn->base kNN ids [n_train][knn_k]
 * (LoadLearnBaseKNN, :2622-2639; only the first M_sq per row are used).  All pointers are device memory. */
typedef struct rg_graph rg_graph; /* device-resident fixed-stride adjacency + entry point */
RG_API rg_status rg_build_roargraph_device(const float *d_base, uint64_t n, uint32_t dim, int metric,
                                           const uint32_t *d_knn_ids, uint64_t n_train, uint32_t knn_k, uint32_t M_sq,
                                           uint32_t M_pjbp, uint32_t L_pjpq, rg_graph **out, int device,
                                           void *cuda_stream);
/* Diagnostic: only the pivot projection prune of every training query (PruneBiSearchBaseGetBase, :1059-1097, 1612-1694),
 * the deterministic first phase of the build: d_lists [n_train][M_pjbp + 1], word 0 = list length. */
RG_API rg_status rg_build_projection_lists_device(const float *d_base, uint64_t n, uint32_t dim, int metric,
                                                  const uint32_t *d_knn_ids, uint64_t n_train, uint32_t knn_k,
                                                  uint32_t M_sq, uint32_t M_pjbp, uint32_t *d_lists, int device,
                                                  void *cuda_stream);
/* Host-buffer variant (what the drop-in IndexBipartite::BuildRoarGraph calls when Parameters holds "gpu_build" != 0):
 * base and knn_ids are host memory; they are uploaded for the duration of the build. */
RG_API rg_status rg_build_roargraph(const float *base, uint64_t n, uint32_t dim, int metric, const uint32_t *knn_ids,
                                    uint64_t n_train, uint32_t knn_k, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq,
                                    rg_graph **out, int device);
/* phase_seconds (may be NULL): 6 doubles = entry point, projection, reverse edges, enhancement searches,
 * enhancement prune + reverse edges, degree check + merge. */
RG_API rg_status rg_graph_info(const rg_graph *graph, uint64_t *n, uint32_t *max_degree, uint64_t *nnz, uint32_t *ep,
                               double *phase_seconds);
/* CSR copy on the host for SaveProjectionGraph (:2606-2619): offsets[n+1], adj[nnz]. */
RG_API rg_status rg_graph_download(const rg_graph *graph, uint64_t *offsets, uint32_t *adj);
RG_API rg_status rg_graph_destroy(rg_graph *graph);
/* Search index over a graph built on the device; d_base is adopted without copying (must outlive the index). */
RG_API rg_status rg_index_create_from_graph(rg_index **out, const float *d_base, uint64_t n, uint32_t dim, int metric,
                                            const rg_graph *graph);

#ifdef __cplusplus
}
#endif
#endif /* ROARGRAPH_B200_H */
