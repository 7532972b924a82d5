/* TEST INFRASTRUCTURE ONLY - see roargraph_oracle.h.  Plain C11 restatement of the reference's
 * search hot path and of the build-time exact kNN.  Written from the reference's behaviour, not
 * copied; every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fno-fast-math): the compiler may neither fuse
 * nor reassociate, so every rounding below happens exactly where it is written; fused steps are
 * explicit fmaf() calls.
 */
#include "roargraph_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * Distances.  include/efanna2e/distance.h:39-89 (L2) and :179-223 (inner product), AS COMPILED by
 * g++ 13.3 -Ofast (objdump of oracle/_ref/libroargraph_ref.so):
 *   main loop, 16 lanes : IP  acc = acc + (a*b)          vmulps, vaddps   (NOT fused)
 *                         L2  d = a-b; acc = acc + (d*d)  vsubps, vmulps, vaddps
 *   fold 16 -> 8        : hi8 + lo8
 *   8-wide tail         : acc8 = fma(a, b, acc8)          vfmadd231ps      (fused)
 *   fold 8 -> 4         : hi4 + lo4
 *   4-wide tail, masked tail (<4 left, zero filled): fused likewise
 *   two hadd            : (x0+x1)+(x2+x3)
 *   IP returns -1.0 * sum (exact sign flip).
 * ---------------------------------------------------------------------------------------------- */
static inline float ip_sum(const float *a, const float *b, unsigned size) {
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    while (size >= 16) {
        for (int j = 0; j < 16; ++j) {
            float p = a[j] * b[j];
            acc[j] = acc[j] + p;
        }
        a += 16; b += 16; size -= 16;
    }
    float m1[8];
    for (int j = 0; j < 8; ++j) m1[j] = acc[j + 8] + acc[j];
    if (size >= 8) {
        for (int j = 0; j < 8; ++j) m1[j] = fmaf(a[j], b[j], m1[j]);
        a += 8; b += 8; size -= 8;
    }
    float m2[4];
    for (int j = 0; j < 4; ++j) m2[j] = m1[j + 4] + m1[j];
    if (size >= 4) {
        for (int j = 0; j < 4; ++j) m2[j] = fmaf(a[j], b[j], m2[j]);
        a += 4; b += 4; size -= 4;
    }
    if (size > 0) {
        float x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
        for (unsigned j = 0; j < size; ++j) { x[j] = a[j]; y[j] = b[j]; }
        for (int j = 0; j < 4; ++j) m2[j] = fmaf(x[j], y[j], m2[j]);
    }
    float t0 = m2[0] + m2[1];
    float t1 = m2[2] + m2[3];
    return t0 + t1;
}

static inline float l2_sum(const float *a, const float *b, unsigned size) {
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    while (size >= 16) {
        for (int j = 0; j < 16; ++j) {
            float d = a[j] - b[j];
            float p = d * d;
            acc[j] = acc[j] + p;
        }
        a += 16; b += 16; size -= 16;
    }
    float m1[8];
    for (int j = 0; j < 8; ++j) m1[j] = acc[j + 8] + acc[j];
    if (size >= 8) {
        for (int j = 0; j < 8; ++j) { float d = a[j] - b[j]; m1[j] = fmaf(d, d, m1[j]); }
        a += 8; b += 8; size -= 8;
    }
    float m2[4];
    for (int j = 0; j < 4; ++j) m2[j] = m1[j + 4] + m1[j];
    if (size >= 4) {
        for (int j = 0; j < 4; ++j) { float d = a[j] - b[j]; m2[j] = fmaf(d, d, m2[j]); }
        a += 4; b += 4; size -= 4;
    }
    if (size > 0) {
        float x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
        for (unsigned j = 0; j < size; ++j) { x[j] = a[j]; y[j] = b[j]; }
        for (int j = 0; j < 4; ++j) { float d = x[j] - y[j]; m2[j] = fmaf(d, d, m2[j]); }
    }
    float t0 = m2[0] + m2[1];
    float t1 = m2[2] + m2[3];
    return t0 + t1;
}

/* src/index.cpp:8-26: L2 -> DistanceL2; INNER_PRODUCT and COSINE -> DistanceInnerProduct;
 * anything else -> DistanceL2. */
float rgo_distance(int metric, const float *a, const float *b, unsigned dim) {
    if (metric == RGO_INNER_PRODUCT || metric == RGO_COSINE) return -ip_sum(a, b, dim);
    return l2_sum(a, b, dim);
}

void rgo_distance_batch(int metric, const float *a, const float *b, unsigned dim, uint64_t n, float *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = rgo_distance(metric, a + i * dim, b + i * dim, dim);
}

/* ------------------------------------------------------------------------------------------------
 * Candidate pool.  include/efanna2e/neighbor.h:21-34 (Neighbor, order (distance, id)) and
 * :138-223 (NeighborPriorityQueue).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t id;
    float distance;
    uint8_t flag;
} rgo_neighbor;

typedef struct {
    size_t size, capacity, cur;
    rgo_neighbor *data; /* capacity + 1 slots, neighbor.h:142 */
} rgo_pool;

static inline int nb_less(const rgo_neighbor *x, const rgo_neighbor *y) { /* neighbor.h:29-31 */
    return x->distance < y->distance || (x->distance == y->distance && x->id < y->id);
}

static void pool_init(rgo_pool *q, size_t capacity) {
    q->size = 0; q->capacity = capacity; q->cur = 0;
    q->data = (rgo_neighbor *)calloc(capacity + 1, sizeof(rgo_neighbor));
}
static void pool_free(rgo_pool *q) { free(q->data); }

/* neighbor.h:150-183 */
static void pool_insert(rgo_pool *q, uint32_t id, float distance) {
    rgo_neighbor nbr = {id, distance, 0};
    if (q->size == q->capacity && nb_less(&q->data[q->size - 1], &nbr)) return;
    size_t lo = 0, hi = q->size;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (nb_less(&nbr, &q->data[mid])) {
            hi = mid;
        } else if (q->data[mid].id == id) {
            return; /* same id already in the set */
        } else {
            lo = mid + 1;
        }
    }
    if (lo < q->capacity) memmove(&q->data[lo + 1], &q->data[lo], (q->size - lo) * sizeof(rgo_neighbor));
    q->data[lo] = nbr;
    if (q->size < q->capacity) q->size++;
    if (lo < q->cur) q->cur = lo;
}

/* neighbor.h:185-192 */
static rgo_neighbor pool_closest_unexpanded(rgo_pool *q) {
    q->data[q->cur].flag = 1;
    size_t pre = q->cur;
    while (q->cur < q->size && q->data[q->cur].flag) q->cur++;
    return q->data[pre];
}
static inline int pool_has_unexpanded(const rgo_pool *q) { return q->cur < q->size; } /* neighbor.h:194 */

uint32_t rgo_pool_script(uint32_t capacity, uint32_t nops, const uint8_t *kind, const uint32_t *ids,
                         const float *dists, uint32_t *out_ids, float *out_dists, uint8_t *out_flags,
                         uint32_t *out_pop, uint32_t *n_pop) {
    rgo_pool q;
    pool_init(&q, capacity);
    uint32_t np = 0;
    for (uint32_t i = 0; i < nops; ++i) {
        if (kind[i] == 0) pool_insert(&q, ids[i], dists[i]);
        else if (pool_has_unexpanded(&q)) out_pop[np++] = pool_closest_unexpanded(&q).id;
    }
    *n_pop = np;
    for (size_t i = 0; i < q.size; ++i) {
        out_ids[i] = q.data[i].id;
        out_dists[i] = q.data[i].distance;
        out_flags[i] = q.data[i].flag;
    }
    uint32_t s = (uint32_t)q.size;
    pool_free(&q);
    return s;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int rgo_omp_num_procs(void) { return omp_get_num_procs(); }

/* ------------------------------------------------------------------------------------------------
 * IndexBipartite::SearchRoarGraph, src/index_bipartite.cpp:2311-2420, with the visited tag array of
 * include/visited_list_pool.h:8-29 (one uint16 array per thread, tag bumped per query, cleared on
 * wrap).  The entry point is inserted but NOT marked visited (:2349 is commented out in the
 * reference), so it is scored a second time if it shows up as a neighbour.
 * ---------------------------------------------------------------------------------------------- */
int rgo_search_roargraph(const float *base, uint64_t n, uint32_t dim, int metric, const uint64_t *adj_offsets,
                         const uint32_t *adj, uint32_t ep, const float *queries, uint64_t nq, uint32_t k,
                         uint32_t L, int num_threads, uint32_t *ids, float *dists, uint32_t *cmps,
                         uint32_t *hops, double *seconds) {
    if (num_threads < 1) num_threads = 1;
    int not_enough = 0;
    double t0 = now_s();
#pragma omp parallel num_threads(num_threads)
    {
        uint16_t *visited = (uint16_t *)calloc(n, sizeof(uint16_t));
        uint16_t tag = 0;
        rgo_pool q;
        pool_init(&q, L);
#pragma omp for schedule(dynamic, 1)
        for (uint64_t qi = 0; qi < nq; ++qi) {
            const float *query = queries + qi * dim;
            q.size = 0; q.cur = 0;                       /* NeighborPriorityQueue search_queue(L_pq) :2314 */
            if (++tag == 0) { memset(visited, 0, n * sizeof(uint16_t)); ++tag; } /* VisitedList::reset */
            float d0 = rgo_distance(metric, base + (uint64_t)ep * dim, query, dim); /* :2338 */
            pool_insert(&q, ep, d0);                                                /* :2344-2345 */
            uint32_t c = 0, h = 0;
            while (pool_has_unexpanded(&q)) {                                       /* :2356 */
                rgo_neighbor cur = pool_closest_unexpanded(&q);                     /* :2358 */
                ++h;                                                                /* :2366 */
                for (uint64_t j = adj_offsets[cur.id]; j < adj_offsets[cur.id + 1]; ++j) { /* :2368 */
                    uint32_t nbr = adj[j];
                    if (visited[nbr] != tag) {                                      /* :2378 */
                        visited[nbr] = tag;                                         /* :2385 */
                        float d = rgo_distance(metric, base + (uint64_t)nbr * dim, query, dim); /* :2387 */
                        ++c;                                                        /* :2397 */
                        pool_insert(&q, nbr, d);                                    /* :2398 */
                    }
                }
            }
            if (q.size < k) {                                                       /* :2408-2412 */
#pragma omp atomic write
                not_enough = 1;
                for (uint32_t i = 0; i < k; ++i) { ids[qi * k + i] = 0xFFFFFFFFu; dists[qi * k + i] = 0.0f; }
            } else {
                for (uint32_t i = 0; i < k; ++i) {                                  /* :2414-2418 */
                    ids[qi * k + i] = q.data[i].id;
                    dists[qi * k + i] = q.data[i].distance;
                }
            }
            if (cmps) cmps[qi] = c;
            if (hops) hops[qi] = h;
        }
        pool_free(&q);
        free(visited);
    }
    if (seconds) *seconds = now_s() - t0;
    return not_enough ? 2 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * IndexBipartite::SearchProjectionGraphInternal, src/index_bipartite.cpp:1279-1350: the beam search of the connectivity
 * enhancement (:1192-1220).  The query is base row tgt; the entry point is scored, inserted and marked visited (:1302-1312);
 * every closest_unexpanded node is appended to full_retset (:1318) before its neighbours are visited; a neighbour equal to
 * tgt or already visited is skipped (:1326), the rest are marked, scored and inserted (:1332-1345).  Output: the first
 * `cap` expanded (id, distance) pairs per target and min(#expanded, cap).  (No compiled-reference pin for this function:
 * it is a private member working on the builder's internal supply_nbrs_; its parts - pool, distance - are pinned.)
 * ---------------------------------------------------------------------------------------------- */
int rgo_search_projection_internal(const float *base, uint64_t n, uint32_t dim, int metric, const uint64_t *adj_offsets,
                                   const uint32_t *adj, uint32_t ep, uint32_t node_lo, uint64_t count, uint32_t L,
                                   uint32_t cap, int num_threads, uint32_t *exp_ids, float *exp_dists, uint32_t *exp_cnt) {
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
    {
        uint8_t *visited = (uint8_t *)malloc(n);
        rgo_pool q;
        pool_init(&q, L);
#pragma omp for schedule(dynamic, 1)
        for (uint64_t t = 0; t < count; ++t) {
            const uint32_t tgt = node_lo + (uint32_t)t;
            const float *query = base + (uint64_t)tgt * dim;
            memset(visited, 0, n);
            q.size = 0; q.cur = 0;
            pool_insert(&q, ep, rgo_distance(metric, base + (uint64_t)ep * dim, query, dim)); /* :1305-1310 */
            visited[ep] = 1;                                                                  /* :1311 */
            uint32_t h = 0;
            while (pool_has_unexpanded(&q)) {                                                 /* :1315 */
                rgo_neighbor cur = pool_closest_unexpanded(&q);                               /* :1317 */
                if (h < cap) { exp_ids[t * cap + h] = cur.id; exp_dists[t * cap + h] = cur.distance; } /* :1319 */
                ++h;
                for (uint64_t j = adj_offsets[cur.id]; j < adj_offsets[cur.id + 1]; ++j) {    /* :1325 */
                    uint32_t nbr = adj[j];
                    if (visited[nbr] || nbr == tgt) continue;                                 /* :1328 */
                    visited[nbr] = 1;                                                         /* :1333 */
                    pool_insert(&q, nbr, rgo_distance(metric, base + (uint64_t)nbr * dim, query, dim)); /* :1335-1345 */
                }
            }
            exp_cnt[t] = h < cap ? h : cap;
        }
        pool_free(&q);
        free(visited);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Exact kNN.  thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:
 *   exact_knn :126-248  - score(q,p) = -<p,q> (mips, :118-120) or squared L2 (:92-100); per query a
 *                         bounded max-heap keeps the k smallest scores (:206-234);
 *   aux_main  :396-448  - the base is cut into parts (PARTSIZE :32), per-part top-k lists are
 *                         concatenated, sorted by score and cut to k (:424-448); mips distances are
 *                         written with the sign flipped back (+ip, :438-441).
 * The reference forms the scores with MKL sgemm (rounding unpinned, see header); here the score is
 * the FP32 lane-ordered distance above, and ties are broken by id (the reference's heap/sort leave
 * ties unspecified; (score, id) ascending is one valid outcome).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float s; uint32_t id; } knn_ent;

static inline int ent_less(knn_ent x, knn_ent y) { return x.s < y.s || (x.s == y.s && x.id < y.id); }

/* max-heap on (s,id): root = worst kept */
static void heap_sift_down(knn_ent *h, uint32_t n, uint32_t i) {
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && ent_less(h[m], h[l])) m = l;
        if (r < n && ent_less(h[m], h[r])) m = r;
        if (m == i) return;
        knn_ent t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}
static void heap_sift_up(knn_ent *h, uint32_t i) {
    while (i > 0) {
        uint32_t p = (i - 1) / 2;
        if (!ent_less(h[p], h[i])) return;
        knn_ent t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}
static int ent_cmp(const void *a, const void *b) {
    knn_ent x = *(const knn_ent *)a, y = *(const knn_ent *)b;
    return ent_less(x, y) ? -1 : (ent_less(y, x) ? 1 : 0);
}

int rgo_exact_knn(const float *base, uint64_t n, const float *queries, uint64_t nq, uint32_t dim, int metric,
                  uint32_t K, uint64_t part_size, int num_threads, uint32_t *ids, float *dists,
                  double *seconds) {
    if (num_threads < 1) num_threads = 1;
    if (part_size == 0) part_size = 20000000ull; /* PARTSIZE, compute_groundtruth.cpp:32 */
    uint64_t nparts = (n + part_size - 1) / part_size;
    int ip = (metric == RGO_INNER_PRODUCT);
    double t0 = now_s();
#pragma omp parallel num_threads(num_threads)
    {
        knn_ent *heap = (knn_ent *)malloc(sizeof(knn_ent) * K);
        knn_ent *all = (knn_ent *)malloc(sizeof(knn_ent) * K * nparts);
#pragma omp for schedule(dynamic, 16)
        for (uint64_t q = 0; q < nq; ++q) {
            const float *qv = queries + q * dim;
            uint64_t nall = 0;
            for (uint64_t part = 0; part < nparts; ++part) {
                uint64_t p0 = part * part_size, p1 = p0 + part_size < n ? p0 + part_size : n;
                uint32_t hn = 0;
                for (uint64_t p = p0; p < p1; ++p) {
                    knn_ent e;
                    e.s = ip ? -ip_sum(base + p * dim, qv, dim) : l2_sum(base + p * dim, qv, dim);
                    e.id = (uint32_t)p;
                    if (hn < K) {
                        heap[hn] = e;
                        heap_sift_up(heap, hn++);
                    } else if (ent_less(e, heap[0])) {
                        heap[0] = e;
                        heap_sift_down(heap, hn, 0);
                    }
                }
                for (uint32_t i = 0; i < hn; ++i) all[nall++] = heap[i];
            }
            qsort(all, nall, sizeof(knn_ent), ent_cmp);
            for (uint32_t j = 0; j < K; ++j) {
                if (j < nall) {
                    ids[q * K + j] = all[j].id;
                    dists[q * K + j] = ip ? -all[j].s : all[j].s;
                } else {
                    ids[q * K + j] = 0xFFFFFFFFu;
                    dists[q * K + j] = 0.0f;
                }
            }
        }
        free(heap);
        free(all);
    }
    if (seconds) *seconds = now_s() - t0;
    return 0;
}

/* tests/test_search_roargraph.cpp:23-36 */
float rgo_compute_recall(uint32_t q_num, uint32_t k, uint32_t gt_dim, const uint32_t *res, const uint32_t *gt) {
    uint32_t total = 0;
    for (uint32_t i = 0; i < q_num; ++i) {
        for (uint32_t a = 0; a < k; ++a) {
            uint32_t p = gt[(uint64_t)i * gt_dim + a];
            for (uint32_t b = 0; b < k; ++b) {
                if (res[(uint64_t)i * k + b] == p) { ++total; break; }
            }
        }
    }
    return (float)total / (float)(k * q_num);
}
