/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's hot path, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
 * library (libroargraph_b200.so) never links, loads or calls anything declared here.
 *
 * Parity status: PINNED for the search path - tests/test_oracle_vs_ref.py checks every function
 * below bit-for-bit against the reference's own translation units compiled into
 * oracle/_ref/libroargraph_ref.so, and against the golden vectors under tests/golden/ that were
 * generated from that library (tests/golden/make_golden.py).
 * UNPINNED for rgo_exact_knn at the GEMM-rounding level: the reference's kNN tool
 * (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp) does its arithmetic in Intel MKL
 * (cblas_sgemm, not vendored, version unpinned: CMake takes /opt/intel/oneapi/mkl/latest or apt
 * libmkl-full-dev), which cannot be built here.  rgo_exact_knn restates its published algorithm
 * (exact top-K under FP32 scores, per-part top-k then merge, +ip sign on output).
 */
#ifndef ROARGRAPH_ORACLE_H
#define ROARGRAPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* efanna2e::Metric, include/efanna2e/distance.h:15 */
enum { RGO_L2 = 0, RGO_INNER_PRODUCT = 1, RGO_COSINE = 4 };

/* DistanceL2::compare (distance.h:39-89) / DistanceInnerProduct::compare (distance.h:179-223),
 * in the operation order GCC 13.3 -Ofast emits for them (see oracle/README in DESIGN.md §oracle). */
float rgo_distance(int metric, const float *a, const float *b, unsigned dim);
void rgo_distance_batch(int metric, const float *a, const float *b, unsigned dim, uint64_t n, float *out);

/* NeighborPriorityQueue (neighbor.h:138-223) driven by a script; mirrors ref_pool_script(). */
uint32_t rgo_pool_script(uint32_t capacity, uint32_t nops, const uint8_t *kind, const uint32_t *ids,
                         const float *dists, uint32_t *out_ids, float *out_dists, uint8_t *out_flags,
                         uint32_t *out_pop, uint32_t *n_pop);

/* IndexBipartite::SearchRoarGraph (src/index_bipartite.cpp:2311-2420) over a CSR copy of
 * projection_graph_, one query per OpenMP iteration like tests/test_search_roargraph.cpp:203-209.
 * base: n rows of dim floats (dim already padded to a multiple of 8 like data_align does).
 * Returns 0, or 2 if some query ended with fewer than k pool entries ("not enough results"). */
int rgo_search_roargraph(const float *base, uint64_t n, uint32_t dim, int metric, const uint64_t *adj_offsets,
                         const uint32_t *adj, uint32_t ep, const float *queries, uint64_t nq, uint32_t k,
                         uint32_t L, int num_threads, uint32_t *ids, float *dists, uint32_t *cmps,
                         uint32_t *hops, double *seconds);

/* exact_knn + aux_main merge (compute_groundtruth.cpp:126-248, 396-448): ids ascending by
 * (score, id); dists written as +ip for RGO_INNER_PRODUCT, squared L2 otherwise. */
int rgo_exact_knn(const float *base, uint64_t n, const float *queries, uint64_t nq, uint32_t dim, int metric,
                  uint32_t K, uint64_t part_size, int num_threads, uint32_t *ids, float *dists,
                  double *seconds);

/* ComputeRecall, tests/test_search_roargraph.cpp:23-36 */
float rgo_compute_recall(uint32_t q_num, uint32_t k, uint32_t gt_dim, const uint32_t *res, const uint32_t *gt);

int rgo_omp_num_procs(void);

#ifdef __cplusplus
}
#endif
#endif
