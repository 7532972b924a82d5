// GPU construction of the projected graph (replaces the CPU phases of
// IndexBipartite::BuildRoarGraph / LinkProjection, /root/reference src/index_bipartite.cpp:143-233, 1043-1277).
//
// The reference builds the graph with per-node locks and order-dependent updates (it is only deterministic at one
// thread).  Here every phase is a BATCH over all nodes (or over a wave of nodes) with the same pruning rules:
//   ep   CalculateProjectionep :2004-2041       centroid + argmin squared L2 (two small kernels)
//   P1   :1059-1097  pivot projection           last training query per pivot wins; PruneBiSearchBaseGetBase
//   P2   :1100-1104  reverse edges, cap M       all reverse candidates of a node collected, then ONE
//                                                PruneProjectionReverseCandidates per overflowing node
//   P4   :1192-1220  connectivity enhancement   waves of nodes: K1 beam search (L_pjpq, build variant) over the supply
//                                                graph as of the previous wave -> PruneProjectionBaseSearchCandidates
//                                                -> reverse edges with cap 2M (PruneProjectionInternalReverseCandidates
//                                                on overflow, without the reference's value-initialised phantom entries)
//   P5   :1224-1248  supply lists > M re-pruned
//   P6   :1251-1269  projection list ++ supply edges not yet present
// The result is a valid RoarGraph with the reference's degree bounds; it is NOT edge-identical to a one-thread CPU
// build (neither are two multi-threaded reference builds); tests gate it on recall parity.  The edge-exact CPU
// restatement stays in mysteryann_b200/host/src/index_bipartite.cpp.
//
// prune_kernel: one warp per list owner.  Candidates are scored against the owner (lane-exact FP32 distances), sorted
// by (distance, id), de-duplicated, then walked in order; a candidate is kept unless some kept member r has
// dist(candidate, r) < dist(candidate, owner) ("occluded").  Kept rows live in shared memory; candidate rows arrive in
// batches of 8 by TMA bulk copies.
#include <cstdlib>
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "rg_distance.cuh"
#include "rg_index.cuh"

struct rg_graph {
    int device = 0;
    uint64_t n = 0;
    uint32_t stride = 0, ep = 0, max_degree = 0;
    uint64_t nnz = 0;
    uint32_t *d_adj = nullptr;
    double seconds[6] = {0, 0, 0, 0, 0, 0};  // ep, P1, P2, P4 search, P4 prune+reverse, P5+P6
};

namespace rg {
rg_status search_expanded_device(rg_index *ix, uint32_t node_lo, uint64_t count, uint32_t L, uint64_t *d_exp_keys,
                                 uint32_t *d_exp_cnt, uint32_t exp_cap, cudaStream_t st);

namespace build {

constexpr uint32_t kCap = 1024;       // candidates per prune
constexpr uint32_t kPlistAt = 768;    // s_orig[kPlistAt..] holds the owner's projection list (base-search variants)
enum Variant { kProjection = 0, kReverse = 1, kInternal = 2, kBaseSearchKeys = 3, kBaseSearchIds = 4 };

struct PruneParams {
    const float *base;
    uint64_t n;
    uint32_t dim, row_stride, M;
    int variant;
    uint32_t *P, *S;          // fixed-stride graphs: word 0 = degree
    uint32_t gstride;
    // kProjection: item = training query; its pruned list goes to T[item] (stride M + 1, word 0 = count)
    const uint32_t *knn;
    uint32_t *T;
    uint32_t knn_k, M_sq;
    // kReverse / kInternal: sorted (destination << 32 | scramble(source)) pairs, heads[i] = first pair of a destination
    const uint64_t *pairs;
    const uint32_t *heads;
    uint64_t n_pairs;
    // kBaseSearch*: owner = node_lo + item
    uint32_t node_lo;
    const uint64_t *exp_keys;
    const uint32_t *exp_cnt;
    uint32_t exp_cap;
    uint32_t off_R, off_stage, off_keys, off_orig, off_kept, off_mbar;
    uint64_t neg_zero2;       // (-0.0f, -0.0f) for lane_exact_distance_x2 (rg_distance.cuh), opaque to the compiler
};

// bijection on 32-bit ids: pairs sort by scrambled source, so a truncated candidate list is a pseudo-random subset
__host__ __device__ __forceinline__ uint32_t scramble(uint32_t x) { return x * 0x9E3779B1u; }
__host__ __device__ __forceinline__ uint32_t unscramble(uint32_t x) { return x * 0x0E8B2F51u; }  // inverse mod 2^32

template <bool kIP>
__global__ void __launch_bounds__(32) prune_kernel(const PruneParams p) {
    extern __shared__ __align__(128) unsigned char sm[];
    const uint32_t lane = threadIdx.x, grp = lane >> 2, t = lane & 3;
    float *s_owner = reinterpret_cast<float *>(sm);
    float *s_R = reinterpret_cast<float *>(sm + p.off_R);
    float *s_stage = reinterpret_cast<float *>(sm + p.off_stage);
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(sm + p.off_keys);
    uint32_t *s_orig = reinterpret_cast<uint32_t *>(sm + p.off_orig);
    uint32_t *s_kept = reinterpret_cast<uint32_t *>(sm + p.off_kept);
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(sm + p.off_mbar);
    uint32_t *s_plist = s_orig + kPlistAt;
    const uint32_t dim = p.dim, n16 = dim >> 4, RS = p.row_stride, M = p.M;
    const bool tail8 = (dim & 15u) != 0;
    const int variant = p.variant;
    uint32_t phase = 0;
    if (lane == 0) {
        mbar_init(s_mbar, 1);
        fence_mbar_init();
    }
    __syncwarp();

    // ---- owner, input, output --------------------------------------------------------------------
    const uint32_t item = blockIdx.x;
    uint32_t owner = 0, n_in = 0, pdeg = 0;
    uint32_t *out_row = nullptr;
    if (variant == kProjection) {
        const uint32_t sq = item;
        const uint32_t *nn = p.knn + size_t(sq) * p.knn_k;
        owner = nn[0];
        const uint32_t cnt = min(min(p.knn_k, p.M_sq), kCap);
        for (uint32_t i0 = 1; i0 < cnt; i0 += 32) {  // nn[1..): everything but the pivot (:1063-1084)
            const uint32_t i = i0 + lane;
            const uint32_t id = (i < cnt) ? nn[i] : 0xFFFFFFFFu;
            const bool ok = id < p.n;
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (ok) s_orig[n_in + __popc(m & lanemask_lt())] = id;
            n_in += __popc(m);
        }
        out_row = p.T + size_t(sq) * (M + 1);
        if (owner >= p.n) {  // no valid pivot (K > n padding)
            if (lane == 0) out_row[0] = 0;
            return;
        }
    } else if (variant == kReverse || variant == kInternal) {
        // current list ++ the new sources of this destination (unique, in scrambled-id order)
        const uint64_t h = p.heads[item];
        owner = uint32_t(p.pairs[h] >> 32);
        const bool rev = variant == kReverse;
        out_row = (rev ? p.P : p.S) + size_t(owner) * p.gstride;
        const uint32_t cur = min(out_row[0], rev ? M : 2 * M);
        for (uint32_t i = lane; i < cur; i += 32) s_orig[i] = out_row[1 + i];
        __syncwarp();
        n_in = cur;
        for (uint64_t i0 = h; i0 < p.n_pairs && n_in < kPlistAt; i0 += 32) {
            const uint64_t i = i0 + lane;
            const uint64_t pr = (i < p.n_pairs) ? p.pairs[i] : ~0ull;
            const bool mine = i < p.n_pairs && uint32_t(pr >> 32) == owner;
            const uint32_t src = unscramble(uint32_t(pr));
            bool fresh = mine && src != owner;
            for (uint32_t j = 0; j < cur && fresh; ++j) fresh = s_orig[j] != src;
            const uint32_t m = __ballot_sync(0xffffffffu, fresh);
            const uint32_t pos = n_in + __popc(m & lanemask_lt());
            if (fresh && pos < kPlistAt) s_orig[pos] = src;
            n_in = min(n_in + uint32_t(__popc(m)), kPlistAt);
            if (__ballot_sync(0xffffffffu, mine) != 0xffffffffu) break;
        }
        __syncwarp();
        if (rev && n_in <= M) {  // room for everybody (:1398-1409)
            for (uint32_t i = cur + lane; i < n_in; i += 32) out_row[1 + i] = s_orig[i];
            if (lane == 0) out_row[0] = n_in;
            return;
        }
    } else {
        owner = p.node_lo + item;
        out_row = p.S + size_t(owner) * p.gstride;
        const uint32_t *prow = p.P + size_t(owner) * p.gstride;
        pdeg = min(prow[0], p.gstride - 1);
        for (uint32_t i = lane; i < pdeg; i += 32) s_plist[i] = prow[1 + i];
        if (variant == kBaseSearchIds) {
            const uint32_t sdeg = min(out_row[0], p.gstride - 1);
            if (sdeg <= M) return;  // :1226
            for (uint32_t i = lane; i < sdeg; i += 32) s_orig[i] = out_row[1 + i];
            n_in = sdeg;
        }
    }
    __syncwarp();

    auto stage = [&](float *dst, uint32_t rows, auto id_of) {
        if (lane == 0) mbar_arrive_expect_tx(s_mbar, rows * dim * 4u);
        __syncwarp();
        for (uint32_t r = lane; r < rows; r += 32)
            bulk_g2s(dst + size_t(r) * RS, p.base + size_t(id_of(r)) * dim, dim * 4u, s_mbar);
        mbar_wait(s_mbar, phase);
        phase ^= 1u;
    };

    // ---- keys: (distance to the owner, id) -------------------------------------------------------------
    uint32_t n_raw;
    if (variant == kBaseSearchKeys) {
        n_raw = min(p.exp_cnt[item], min(p.exp_cap, kCap));
        const uint64_t *src = p.exp_keys + size_t(item) * p.exp_cap;
        for (uint32_t i = lane; i < n_raw; i += 32) {
            const uint64_t k = src[i];
            s_keys[i] = (key_id(k) == owner) ? ~0ull : k;  // :1203-1208
        }
    } else {
        n_raw = n_in;
        stage(s_owner, 1, [&](uint32_t) { return owner; });
        for (uint32_t b0 = 0; b0 < n_in; b0 += 8) {
            const uint32_t rows = min(8u, n_in - b0);
            stage(s_stage, rows, [&](uint32_t r) { return s_orig[b0 + r]; });
            const bool valid = grp < rows;
            const float4 *rp = reinterpret_cast<const float4 *>(s_stage + size_t(valid ? grp : 0) * RS) + t;
            const float4 *qp = reinterpret_cast<const float4 *>(s_owner) + t;
            const float d = lane_exact_distance_x2<kIP>(rp, qp, n16, tail8, t, p.neg_zero2);
            if (valid && t == 0) {
                const uint32_t id = s_orig[b0 + grp];
                s_keys[b0 + grp] = (id == owner) ? ~0ull : make_key(d, id);
            }
            __syncwarp();
        }
    }
    const uint32_t P2 = max(32u, pow2_at_least(n_raw));
    for (uint32_t i = n_raw + lane; i < P2; i += 32) s_keys[i] = ~0ull;
    __syncwarp();
    warp_sort_u64(s_keys, P2, lane);
    // drop repeated ids (same id -> same key -> adjacent) and the ~0 sentinels
    uint32_t n_c = 0;
    {
        uint64_t prev_last = ~0ull;
        for (uint32_t i0 = 0; i0 < P2; i0 += 32) {
            const uint64_t k = s_keys[i0 + lane];
            uint64_t prev = __shfl_up_sync(0xffffffffu, k, 1);
            if (lane == 0) prev = prev_last;
            prev_last = __shfl_sync(0xffffffffu, k, 31);
            const bool keep = k != ~0ull && k != prev;
            __syncwarp();
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) s_keys[n_c + __popc(m & lanemask_lt())] = k;
            n_c += __popc(m);
            __syncwarp();
        }
    }

    auto contains = [&](const uint32_t *list, uint32_t len, uint32_t id) {
        bool hit = false;
        for (uint32_t i0 = 0; i0 < len; i0 += 32) hit |= (i0 + lane < len) && list[i0 + lane] == id;
        return __any_sync(0xffffffffu, hit);
    };

    // ---- occlusion walks ---------------------------------------------------------------------------------
    uint32_t nkept = 0;
    uint32_t first = 0;
    if (variant == kBaseSearchKeys || variant == kBaseSearchIds)  // leading projection neighbours are skipped (:1862-1864)
        while (first < n_c && contains(s_plist, pdeg, key_id(s_keys[first]))) ++first;
    auto walk = [&](uint32_t begin, uint32_t end) {
        for (uint32_t b0 = begin; b0 < end && nkept < M; b0 += 8) {
            const uint32_t rows = min(8u, end - b0);
            stage(s_stage, rows, [&](uint32_t r) { return key_id(s_keys[b0 + r]); });
            for (uint32_t c = 0; c < rows && nkept < M; ++c) {
                const uint64_t key = s_keys[b0 + c];
                const uint32_t id = key_id(key);
                const float pd = key_dist(key);
                bool occluded = contains(s_kept, nkept, id);
                for (uint32_t r0 = 0; r0 < nkept && !occluded; r0 += 8) {
                    const uint32_t r = r0 + grp;
                    const bool valid = r < nkept;
                    const float4 *ap = reinterpret_cast<const float4 *>(s_stage + size_t(c) * RS) + t;
                    const float4 *bp = reinterpret_cast<const float4 *>(s_R + size_t(valid ? r : 0) * RS) + t;
                    const float d = lane_exact_distance_x2<kIP>(ap, bp, n16, tail8, t, p.neg_zero2);
                    occluded = __any_sync(0xffffffffu, valid && t == 0 && d < pd);
                }
                if (!occluded && id != owner) {
                    const float4 *src = reinterpret_cast<const float4 *>(s_stage + size_t(c) * RS);
                    float4 *dst = reinterpret_cast<float4 *>(s_R + size_t(nkept) * RS);
                    for (uint32_t i = lane; i < (dim >> 2); i += 32) dst[i] = src[i];
                    if (lane == 0) s_kept[nkept] = id;
                    ++nkept;
                    __syncwarp();
                }
            }
            __syncwarp();
        }
    };
    if (first < n_c) {
        const uint32_t id0 = key_id(s_keys[first]);  // the nearest candidate is kept unconditionally
        stage(s_R, 1, [&](uint32_t) { return id0; });
        if (lane == 0) s_kept[0] = id0;
        nkept = 1;
        __syncwarp();
        walk(first + 1, n_c);
        if (variant == kBaseSearchKeys || variant == kBaseSearchIds) walk(1, first);  // second walk (:1896-1926)
    }
    // ---- fill passes ---------------------------------------------------------------------------------------
    if (variant == kProjection) {  // nearest first (:1685-1691)
        for (uint32_t i = 1; i < n_c && nkept < M; ++i) {
            const uint32_t id = key_id(s_keys[i]);
            if (!contains(s_kept, nkept, id)) {
                if (lane == 0) s_kept[nkept] = id;
                ++nkept;
                __syncwarp();
            }
        }
    } else if (variant == kReverse) {  // original order (:1596-1600)
        for (uint32_t i = 0; i < n_in && nkept < M; ++i) {
            const uint32_t id = s_orig[i];
            if (id != owner && !contains(s_kept, nkept, id)) {
                if (lane == 0) s_kept[nkept] = id;
                ++nkept;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < nkept; i += 32) out_row[1 + i] = s_kept[i];
    if (lane == 0) out_row[0] = nkept;
}

// ---- small kernels ---------------------------------------------------------------------------------------------------
__global__ void column_sum_kernel(const float *__restrict__ base, uint64_t n, uint32_t dim, double *__restrict__ sums) {
    const uint64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t r0 = uint64_t(blockIdx.x) * rows_per, r1 = min(n, r0 + rows_per);
    for (uint32_t c = threadIdx.x; c < dim; c += blockDim.x) {
        double acc = 0;
        for (uint64_t r = r0; r < r1; ++r) acc += double(base[r * dim + c]);
        atomicAdd(&sums[c], acc);
    }
}
__global__ void ep_argmin_kernel(const float *__restrict__ base, uint64_t n, uint32_t dim, const double *__restrict__ sums,
                                 unsigned long long *__restrict__ best) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t r = warp; r < n; r += nwarps) {
        float acc = 0.f;
        for (uint32_t c = lane; c < dim; c += 32) {
            const float d = float(sums[c] / double(n)) - base[r * dim + c];
            acc += d * d;
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicMin(best, (unsigned long long)((uint64_t(float_to_ordered(acc)) << 32) | uint32_t(r)));
    }
}
__global__ void zero_degrees_kernel(uint32_t *g, uint64_t n, uint32_t stride) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) g[i * stride] = 0;
}
// P1: the LAST training query with a given pivot owns it (sequential "last writer wins", :1090)
__global__ void pivot_owner_kernel(const uint32_t *__restrict__ knn, uint64_t n_train, uint32_t knn_k, uint64_t n, uint32_t *owner) {
    const uint64_t sq = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (sq >= n_train) return;
    const uint32_t pivot = knn[sq * knn_k];
    if (pivot < n) atomicMax(&owner[pivot], uint32_t(sq) + 1u);
}
// the owner's pruned list becomes the pivot's forward list
__global__ void copy_owner_lists_kernel(const uint32_t *__restrict__ knn, uint64_t n_train, uint32_t knn_k, uint64_t n,
                                        const uint32_t *__restrict__ owner, const uint32_t *__restrict__ T, uint32_t M,
                                        uint32_t *P, uint32_t stride) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t sq = warp; sq < n_train; sq += nwarps) {
        const uint32_t pivot = knn[sq * knn_k];
        if (pivot >= n || owner[pivot] != uint32_t(sq) + 1u) continue;
        const uint32_t *t = T + sq * (M + 1);
        const uint32_t cnt = min(t[0], M);
        uint32_t *row = P + size_t(pivot) * stride;
        for (uint32_t j = lane; j < cnt; j += 32) row[1 + j] = t[1 + j];
        if (lane == 0) row[0] = cnt;
    }
}
// P1 + P2 as one batch: every training query's pruned list T(sq) links its pivot p with each member d.  The
// reference inserts p into list(d) right away (ProjectionAddReverse, :1391-1432) and, in P2, d back into list(p) when p's
// own list was overwritten by a later query; here both directions become (destination, source) pairs.
__global__ void projection_pairs_kernel(const uint32_t *__restrict__ knn, uint64_t n_train, uint32_t knn_k, uint64_t n,
                                        const uint32_t *__restrict__ owner, const uint32_t *__restrict__ T, uint32_t M,
                                        uint64_t *pairs, unsigned long long *n_pairs) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t sq = warp; sq < n_train; sq += nwarps) {
        const uint32_t pivot = knn[sq * knn_k];
        if (pivot >= n) continue;
        const bool is_owner = owner[pivot] == uint32_t(sq) + 1u;
        const uint32_t *t = T + sq * (M + 1);
        const uint32_t cnt = min(t[0], M);
        const uint32_t per = is_owner ? 1u : 2u;
        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool valid = j < cnt;
            const uint32_t m = __ballot_sync(0xffffffffu, valid);
            unsigned long long pos0 = 0;
            if (lane == 0) pos0 = atomicAdd(n_pairs, (unsigned long long)(per * __popc(m)));
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
            if (valid) {
                const uint32_t d = t[1 + j];
                const unsigned long long pos = pos0 + per * __popc(m & lanemask_lt());
                pairs[pos] = (uint64_t(d) << 32) | scramble(pivot);
                if (!is_owner) pairs[pos + 1] = (uint64_t(pivot) << 32) | scramble(d);
            }
        }
    }
}
// P4: reverse edges of the wave's new supply lists, cap 2M; overflowing (destination, source) pairs are collected
__global__ void supply_append_kernel(uint32_t *S, uint32_t stride, uint32_t cap, uint32_t node_lo, uint32_t count,
                                     uint64_t *ovf, uint32_t *ovf_count, uint32_t ovf_cap) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;
    const uint32_t s = node_lo + warp;
    const uint32_t *row = S + size_t(s) * stride;
    const uint32_t deg = min(row[0], cap);
    for (uint32_t j = lane; j < deg; j += 32) {
        const uint32_t d = row[1 + j];
        uint32_t *drow = S + size_t(d) * stride;
        const uint32_t dd = min(*(volatile uint32_t *)drow, cap);
        bool present = (d == s);
        for (uint32_t i = 0; i < dd && !present; ++i) present = ((volatile uint32_t *)drow)[1 + i] == s;
        if (present) continue;
        const uint32_t pos = atomicAdd(drow, 1u);
        if (pos < cap) {
            drow[1 + pos] = s;
        } else {
            const uint32_t o = atomicAdd(ovf_count, 1u);
            if (o < ovf_cap) ovf[o] = (uint64_t(d) << 32) | scramble(s);
        }
    }
}
__global__ void segment_heads_kernel(const uint64_t *__restrict__ pairs, uint64_t n_pairs, uint32_t *heads, uint32_t *n_heads) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    if (i == 0 || (pairs[i] >> 32) != (pairs[i - 1] >> 32)) heads[atomicAdd(n_heads, 1u)] = uint32_t(i);
}
// P6: projection list ++ supply edges that are not projection edges yet (:1251-1269); also degree statistics
__global__ void merge_supply_kernel(uint32_t *P, const uint32_t *__restrict__ S, uint64_t n, uint32_t stride, uint32_t M,
                                    unsigned long long *nnz, uint32_t *max_deg) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t v = warp; v < n; v += nwarps) {
        uint32_t *prow = P + v * stride;
        const uint32_t *srow = S + v * stride;
        const uint32_t pdeg = min(prow[0], stride - 1), sdeg = min(srow[0], stride - 1);
        uint32_t added = 0;
        for (uint32_t j0 = 0; j0 < sdeg; j0 += 32) {
            const uint32_t j = j0 + lane;
            const uint32_t x = (j < sdeg) ? srow[1 + j] : 0xFFFFFFFFu;
            bool fresh = j < sdeg;
            for (uint32_t i = 0; i < pdeg && fresh; ++i) fresh = prow[1 + i] != x;
            const uint32_t m = __ballot_sync(0xffffffffu, fresh);
            const uint32_t pos = pdeg + added + __popc(m & lanemask_lt());
            if (fresh && added + __popc(m & lanemask_lt()) < 2 * M && pos + 1 < stride) prow[1 + pos] = x;
            added += __popc(m);
        }
        const uint32_t deg = min(min(pdeg + min(added, 2 * M), stride - 1), pdeg + added);
        if (lane == 0) {
            prow[0] = deg;
            atomicAdd(nnz, (unsigned long long)deg);
            atomicMax(max_deg, deg);
        }
    }
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t alloc(T **p, uint64_t count) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), std::max<uint64_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static uint32_t round_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// d_p1_lists != nullptr: stop after the pivot projection prune (P1) and hand out its per-training-query lists
// ([n_train][M + 1], word 0 = length) - the deterministic part of the build, compared list by list with the host's
// PruneBiSearchBaseGetBase restatement by tests/test_build_gpu.py
rg_status build_device(const float *d_base, uint64_t n, uint32_t dim, int metric, const uint32_t *d_knn, uint64_t n_train,
                       uint32_t knn_k, uint32_t M_sq, uint32_t M, uint32_t L_pjpq, rg_graph *g, cudaStream_t st,
                       uint32_t *d_p1_lists = nullptr) {
    const bool ip = metric != RG_METRIC_L2;
    int dev = 0, sms = 0, smem_max = 0;
    RG_CUDA_OK(cudaGetDevice(&dev));
    RG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RG_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const uint32_t stride = round_up(2 * M + 1, 8);
    g->device = dev;
    g->n = n;
    g->stride = stride;
    RG_CUDA_OK(cudaMalloc(&g->d_adj, n * uint64_t(stride) * sizeof(uint32_t)));
    uint32_t *P = g->d_adj;

    Scratch sc;
    uint32_t *S = nullptr, *owner = nullptr, *cnt = nullptr, *T = nullptr, *heads = nullptr;
    uint64_t *pairs = nullptr, *pairs_sorted = nullptr;
    double *sums = nullptr;
    unsigned long long *best = nullptr;
    // pair buffers serve P2 (two directions of every pruned list) and, per wave, the supply overflow of P4
    // Nodes per connectivity-enhancement wave.  The searches of a wave see the supply graph as of the previous wave, while in
    // the reference a node sees the lists and reverse edges of (nearly) every node processed before it (:1192-1220), and
    // that staleness costs graph quality: on C1 (100K nodes) waves of 16 % / 4 % / 1 % / 0.5 % of the nodes end 0.050 /
    // 0.007 / 0.0016 / 0.001 below the reference's recall@10 curve, the last one inside the spread of two reference builds
    // (profiles/r02_build_quality_c1_waves.txt).  So a wave is n/256 of the nodes (>= 128, <= 131072); RG_BUILD_WAVE overrides.
    uint32_t wave = uint32_t(std::min<uint64_t>(131072, std::max<uint64_t>(128, round_up(uint32_t(n / 256 + 1), 128))));
    if (const char *e = std::getenv("RG_BUILD_WAVE")) wave = uint32_t(std::max<long long>(32, std::min<long long>(1 << 20, atoll(e))));
    const uint64_t pair_cap = std::max<uint64_t>(n_train * 2ull * M, uint64_t(wave) * M);
    RG_CUDA_OK(sc.alloc(&S, n * uint64_t(stride)));
    RG_CUDA_OK(sc.alloc(&owner, n));
    RG_CUDA_OK(sc.alloc(&T, n_train * uint64_t(M + 1)));
    RG_CUDA_OK(sc.alloc(&cnt, 16));
    RG_CUDA_OK(sc.alloc(&pairs, pair_cap));
    RG_CUDA_OK(sc.alloc(&pairs_sorted, pair_cap));
    RG_CUDA_OK(sc.alloc(&heads, std::min<uint64_t>(pair_cap, n)));
    RG_CUDA_OK(sc.alloc(&sums, dim));
    RG_CUDA_OK(sc.alloc(&best, 2));
    size_t sort_bytes = 0, uniq_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, pairs, pairs_sorted, pair_cap, 0, 64, st);
    cub::DeviceSelect::Unique(nullptr, uniq_bytes, pairs_sorted, pairs, best, pair_cap, st);
    unsigned char *cub_tmp = nullptr;
    const size_t cub_bytes = std::max(sort_bytes, uniq_bytes);
    RG_CUDA_OK(sc.alloc(&cub_tmp, cub_bytes));
    const int wide = sms * 16;

    // prune kernel geometry
    PruneParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.neg_zero2 = kNegZero2;
    pp.base = d_base;
    pp.n = n;
    pp.dim = dim;
    pp.row_stride = (dim % 32 <= 16) ? dim - dim % 32 + 16 : dim - dim % 32 + 48;
    pp.M = M;
    pp.P = P;
    pp.S = S;
    pp.gstride = stride;
    pp.knn = d_knn;
    pp.T = T;
    pp.knn_k = knn_k;
    pp.M_sq = M_sq;
    const uint32_t row_bytes = pp.row_stride * 4;
    uint32_t off = round_up(row_bytes, 128);
    pp.off_R = off;
    off += round_up(M * row_bytes, 128);
    pp.off_stage = off;
    off += round_up(8 * row_bytes, 128);
    pp.off_keys = off;
    off += kCap * 8;
    pp.off_orig = off;
    off += kCap * 4;
    pp.off_kept = off;
    off += round_up(M * 4, 128);
    pp.off_mbar = off;
    off += 128;
    const size_t prune_smem = off;
    if (prune_smem > size_t(smem_max)) return fail(RG_ERR_INVALID_ARGUMENT, "M_pjbp=%u x dim=%u does not fit the prune kernel's shared memory", M, dim);
    if (stride - 1 > kPlistAt / 4 || M == 0) return fail(RG_ERR_INVALID_ARGUMENT, "M_pjbp=%u out of range", M);
    auto prune_fn = ip ? prune_kernel<true> : prune_kernel<false>;
    RG_CUDA_OK(cudaFuncSetAttribute(prune_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prune_smem)));
    auto run_prune = [&](int variant, uint32_t n_items) -> cudaError_t {
        if (n_items == 0) return cudaSuccess;
        pp.variant = variant;
        prune_fn<<<n_items, 32, prune_smem, st>>>(pp);
        return cudaGetLastError();
    };

    // ---- entry point ---------------------------------------------------------------------------------------------
    double t0 = now_s();
    RG_CUDA_OK(cudaMemsetAsync(sums, 0, dim * sizeof(double), st));
    RG_CUDA_OK(cudaMemsetAsync(best, 0xff, 2 * sizeof(unsigned long long), st));
    column_sum_kernel<<<std::min<uint64_t>(n, uint64_t(sms) * 8), 256, 0, st>>>(d_base, n, dim, sums);
    ep_argmin_kernel<<<wide, 256, 0, st>>>(d_base, n, dim, sums, best);
    unsigned long long h_best = 0;
    RG_CUDA_OK(cudaMemcpyAsync(&h_best, best, sizeof(h_best), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    g->ep = uint32_t(h_best & 0xffffffffull);
    g->seconds[0] = now_s() - t0;

    // ---- P1: every training query prunes its neighbour list around its pivot --------------------------------------
    t0 = now_s();
    zero_degrees_kernel<<<wide, 256, 0, st>>>(P, n, stride);
    RG_CUDA_OK(cudaMemsetAsync(owner, 0, n * sizeof(uint32_t), st));
    const unsigned tb = unsigned((n_train + 255) / 256);
    pivot_owner_kernel<<<tb, 256, 0, st>>>(d_knn, n_train, knn_k, n, owner);
    RG_CUDA_OK(run_prune(kProjection, uint32_t(n_train)));
    if (d_p1_lists) {
        RG_CUDA_OK(cudaMemcpyAsync(d_p1_lists, T, n_train * uint64_t(M + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        return RG_OK;
    }
    copy_owner_lists_kernel<<<wide, 256, 0, st>>>(d_knn, n_train, knn_k, n, owner, T, M, P, stride);
    RG_CUDA_OK(cudaStreamSynchronize(st));
    g->seconds[1] = now_s() - t0;

    // ---- P2: reverse edges -----------------------------------------------------------------------------------------
    t0 = now_s();
    uint32_t h_cnt[16];
    RG_CUDA_OK(cudaMemsetAsync(best, 0, 2 * sizeof(unsigned long long), st));
    RG_CUDA_OK(cudaMemsetAsync(cnt, 0, 16 * sizeof(uint32_t), st));
    projection_pairs_kernel<<<wide, 256, 0, st>>>(d_knn, n_train, knn_k, n, owner, T, M, pairs, best);
    RG_CUDA_OK(cudaMemcpyAsync(&h_best, best, sizeof(h_best), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    if (h_best) {
        const uint64_t n_raw = h_best;
        RG_CUDA_OK(cub::DeviceRadixSort::SortKeys(cub_tmp, sort_bytes, pairs, pairs_sorted, n_raw, 0, 64, st));
        RG_CUDA_OK(cub::DeviceSelect::Unique(cub_tmp, uniq_bytes, pairs_sorted, pairs, best + 1, n_raw, st));
        RG_CUDA_OK(cudaMemcpyAsync(&h_best, best + 1, sizeof(h_best), cudaMemcpyDeviceToHost, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        const uint64_t n_uniq = h_best;
        segment_heads_kernel<<<unsigned((n_uniq + 255) / 256), 256, 0, st>>>(pairs, n_uniq, heads, cnt + 1);
        RG_CUDA_OK(cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        pp.pairs = pairs;
        pp.heads = heads;
        pp.n_pairs = n_uniq;
        RG_CUDA_OK(run_prune(kReverse, h_cnt[1]));
    }
    RG_CUDA_OK(cudaStreamSynchronize(st));
    g->seconds[2] = now_s() - t0;

    // ---- P4: connectivity enhancement in waves --------------------------------------------------------------------
    RG_CUDA_OK(cudaMemcpyAsync(S, P, n * uint64_t(stride) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));  // :1183-1188
    const uint32_t exp_cap = std::min<uint32_t>(kCap, L_pjpq + L_pjpq / 2 + 16);
    uint64_t *exp_keys = nullptr;
    uint32_t *exp_cnt = nullptr;
    uint64_t *ovf = pairs, *ovf_sorted = pairs_sorted;
    const uint32_t ovf_cap = wave * M;
    RG_CUDA_OK(sc.alloc(&exp_keys, uint64_t(wave) * exp_cap));
    RG_CUDA_OK(sc.alloc(&exp_cnt, wave));

    rg_index view;  // search view of the supply graph
    view.device = dev;
    view.n = n;
    view.dim = dim;
    view.metric = metric;
    view.ep = g->ep;
    view.max_degree = stride - 1;
    view.adj_stride = stride;
    view.d_base = d_base;
    view.d_adj = S;
    view.sm_count = sms;
    view.max_smem_optin = smem_max;
    // K1 cache hints for the build searches (tools/microbench_build.py measures the combinations)
    if (const char *e = std::getenv("RG_BUILD_L2_HINT")) view.cfg_l2_hint = atoi(e) & 3;
    if (const char *e = std::getenv("RG_BUILD_ADJ_PREFETCH")) view.cfg_adj_prefetch = atoi(e) & 3;
    if (const char *e = std::getenv("RG_BUILD_WARPS")) view.cfg_warps = std::max(1, std::min(8, atoi(e)));
    RG_CUDA_OK(sc.alloc(&view.d_counters, 64));
    RG_CUDA_OK(cudaMemsetAsync(view.d_counters, 0, 64 * sizeof(uint32_t), st));
    pp.exp_keys = exp_keys;
    pp.exp_cnt = exp_cnt;
    pp.exp_cap = exp_cap;
    pp.pairs = ovf_sorted;
    pp.heads = heads;
    double t_search = 0, t_prune = 0;
    rg_status status = RG_OK;
    // every failure inside the loop becomes a status (the scratch of `view` is released on all paths below): a wave that
    // dropped out silently would leave part of the nodes without their connectivity edges and still return RG_OK
    auto cuda_fail = [&](cudaError_t e, const char *what) {
        status = rg::fail(e == cudaErrorMemoryAllocation ? RG_ERR_OUT_OF_MEMORY : RG_ERR_CUDA, "connectivity enhancement: %s failed: %s", what,
                          cudaGetErrorString(e));
        (void)cudaGetLastError();
        return false;
    };
#define RG_WAVE_OK(expr)                                   \
    {                                                      \
        const cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess && !cuda_fail(_e, #expr)) break; \
    }
    for (uint64_t lo = 0; lo < n && status == RG_OK; lo += wave) {
        const uint32_t count = uint32_t(std::min<uint64_t>(wave, n - lo));
        t0 = now_s();
        status = search_expanded_device(&view, uint32_t(lo), count, L_pjpq, exp_keys, exp_cnt, exp_cap, st);
        if (status != RG_OK) break;
        RG_WAVE_OK(cudaStreamSynchronize(st));
        t_search += now_s() - t0;
        t0 = now_s();
        pp.node_lo = uint32_t(lo);
        RG_WAVE_OK(run_prune(kBaseSearchKeys, count));
        RG_WAVE_OK(cudaMemsetAsync(cnt, 0, 16 * sizeof(uint32_t), st));
        supply_append_kernel<<<(count + 7) / 8, 256, 0, st>>>(S, stride, 2 * M, uint32_t(lo), count, ovf, cnt, ovf_cap);
        RG_WAVE_OK(cudaGetLastError());
        RG_WAVE_OK(cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        RG_WAVE_OK(cudaStreamSynchronize(st));
        const uint32_t n_ovf = std::min(h_cnt[0], ovf_cap);
        if (n_ovf) {
            RG_WAVE_OK(cub::DeviceRadixSort::SortKeys(cub_tmp, sort_bytes, ovf, ovf_sorted, uint64_t(n_ovf), 0, 64, st));
            segment_heads_kernel<<<(n_ovf + 255) / 256, 256, 0, st>>>(ovf_sorted, n_ovf, heads, cnt + 1);
            RG_WAVE_OK(cudaGetLastError());
            RG_WAVE_OK(cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
            RG_WAVE_OK(cudaStreamSynchronize(st));
            pp.n_pairs = n_ovf;
            RG_WAVE_OK(run_prune(kInternal, h_cnt[1]));
        }
        RG_WAVE_OK(cudaStreamSynchronize(st));
        t_prune += now_s() - t0;
    }
#undef RG_WAVE_OK
    cudaFree(view.d_overflow_list);
    cudaFree(view.d_ghash);
    view.d_overflow_list = nullptr;
    view.d_ghash = nullptr;
    if (status != RG_OK) return status;
    RG_CUDA_OK(cudaGetLastError());
    RG_CUDA_OK(cudaStreamSynchronize(st));
    g->seconds[3] = t_search;
    g->seconds[4] = t_prune;

    // ---- P5 + P6 -------------------------------------------------------------------------------------------------
    t0 = now_s();
    pp.node_lo = 0;
    RG_CUDA_OK(run_prune(kBaseSearchIds, uint32_t(n)));
    RG_CUDA_OK(cudaMemsetAsync(best, 0, 2 * sizeof(unsigned long long), st));
    RG_CUDA_OK(cudaMemsetAsync(cnt, 0, 16 * sizeof(uint32_t), st));
    merge_supply_kernel<<<wide, 256, 0, st>>>(P, S, n, stride, M, best, cnt);
    RG_CUDA_OK(cudaMemcpyAsync(&h_best, best, sizeof(h_best), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    g->nnz = h_best;
    g->max_degree = h_cnt[0];
    g->seconds[5] = now_s() - t0;
    return RG_OK;
}

__global__ void degrees_to_u64_kernel(const uint32_t *__restrict__ adj, uint64_t n, uint32_t stride, uint64_t *deg) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
        deg[i] = adj[i * stride];
}
__global__ void compact_rows_kernel(const uint32_t *__restrict__ adj, uint64_t n, uint32_t stride, const uint64_t *__restrict__ off,
                                    uint32_t *out) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t v = warp; v < n; v += nwarps) {
        const uint32_t deg = adj[v * stride];
        for (uint32_t j = lane; j < deg; j += 32) out[off[v] + j] = adj[v * stride + 1 + j];
    }
}

}  // namespace build
}  // namespace rg

extern "C" {

rg_status rg_build_roargraph_device(const float *d_base, uint64_t n, uint32_t dim, int metric, const uint32_t *d_knn_ids,
                                    uint64_t n_train, uint32_t knn_k, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq,
                                    rg_graph **out, int device, void *cuda_stream) {
    if (!d_base || !d_knn_ids || !out) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph_device: null argument");
    if (n == 0 || n >= (1ull << 31) || n_train == 0 || n_train >= (1ull << 32) - 1)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph_device: sizes out of range");
    if (dim == 0 || dim % 8 != 0) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph_device: dim must be a multiple of 8");
    if (n_train * 2ull * M_pjbp >= (1ull << 32))  // segment_heads_kernel stores pair offsets as 32-bit words
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph_device: n_train * 2 * M_pjbp must stay below 2^32 (got %llu x 2 x %u)",
                        (unsigned long long)n_train, M_pjbp);
    if (knn_k == 0 || M_sq == 0 || M_pjbp == 0 || M_pjbp > 64 || L_pjpq == 0 || L_pjpq > 8192)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph_device: need knn_k, M_sq >= 1, 1 <= M_pjbp <= 64, 1 <= L_pjpq <= 8192");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg_graph *g = new rg_graph();
    const rg_status s = rg::build::build_device(d_base, n, dim, metric, d_knn_ids, n_train, knn_k, M_sq, M_pjbp, L_pjpq, g,
                                                static_cast<cudaStream_t>(cuda_stream));
    if (s != RG_OK) {
        cudaFree(g->d_adj);
        delete g;
        cudaGetLastError();
        return s;
    }
    *out = g;
    return RG_OK;
}

// Host-buffer variant: uploads the base rows and the kNN ids, builds on the device, frees the uploads.
rg_status rg_build_roargraph(const float *base, uint64_t n, uint32_t dim, int metric, const uint32_t *knn_ids, uint64_t n_train,
                             uint32_t knn_k, uint32_t M_sq, uint32_t M_pjbp, uint32_t L_pjpq, rg_graph **out, int device) {
    if (!base || !knn_ids || !out) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_roargraph: null argument");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg::build::Scratch sc;
    float *d_base = nullptr;
    uint32_t *d_knn = nullptr;
    RG_CUDA_OK(sc.alloc(&d_base, n * uint64_t(dim)));
    RG_CUDA_OK(sc.alloc(&d_knn, n_train * uint64_t(knn_k)));
    RG_CUDA_OK(cudaMemcpy(d_base, base, n * uint64_t(dim) * sizeof(float), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMemcpy(d_knn, knn_ids, n_train * uint64_t(knn_k) * sizeof(uint32_t), cudaMemcpyHostToDevice));
    rg_status s = rg_build_roargraph_device(d_base, n, dim, metric, d_knn, n_train, knn_k, M_sq, M_pjbp, L_pjpq, out, device, nullptr);
    if (s != RG_OK) return s;
    RG_CUDA_OK(cudaDeviceSynchronize());
    return RG_OK;
}

rg_status rg_graph_info(const rg_graph *g, uint64_t *n, uint32_t *max_degree, uint64_t *nnz, uint32_t *ep, double *phase_seconds) {
    if (!g) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_graph_info: null graph");
    if (n) *n = g->n;
    if (max_degree) *max_degree = g->max_degree;
    if (nnz) *nnz = g->nnz;
    if (ep) *ep = g->ep;
    if (phase_seconds) memcpy(phase_seconds, g->seconds, sizeof(g->seconds));
    return RG_OK;
}

rg_status rg_graph_download(const rg_graph *g, uint64_t *offsets, uint32_t *adj) {
    if (!g || !offsets || (!adj && g->nnz)) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_graph_download: null argument");
    rg::DeviceGuard guard(g->device);
    rg::build::Scratch sc;
    uint64_t *d_deg = nullptr, *d_off = nullptr;
    uint32_t *d_out = nullptr;
    RG_CUDA_OK(sc.alloc(&d_deg, g->n + 1));
    RG_CUDA_OK(sc.alloc(&d_off, g->n + 1));
    RG_CUDA_OK(sc.alloc(&d_out, g->nnz));
    RG_CUDA_OK(cudaMemset(d_deg, 0, (g->n + 1) * sizeof(uint64_t)));
    rg::build::degrees_to_u64_kernel<<<1024, 256>>>(g->d_adj, g->n, g->stride, d_deg);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_deg, d_off, g->n + 1);
    unsigned char *tmp = nullptr;
    RG_CUDA_OK(sc.alloc(&tmp, tmp_bytes));
    RG_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_deg, d_off, g->n + 1));
    rg::build::compact_rows_kernel<<<1024, 256>>>(g->d_adj, g->n, g->stride, d_off, d_out);
    RG_CUDA_OK(cudaMemcpy(offsets, d_off, (g->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (g->nnz) RG_CUDA_OK(cudaMemcpy(adj, d_out, g->nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RG_OK;
}

rg_status rg_graph_destroy(rg_graph *g) {
    if (!g) return RG_OK;
    rg::DeviceGuard guard(g->device);
    cudaFree(g->d_adj);
    cudaGetLastError();
    delete g;
    return RG_OK;
}

rg_status rg_index_create_from_graph(rg_index **out, const float *d_base, uint64_t n, uint32_t dim, int metric,
                                     const rg_graph *g) {
    if (!out || !d_base || !g) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create_from_graph: null argument");
    if (n != g->n) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create_from_graph: graph has %llu nodes, base %llu",
                                   (unsigned long long)g->n, (unsigned long long)n);
    if (dim == 0 || dim % 8 != 0) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_index_create_from_graph: dim must be a multiple of 8");
    rg::DeviceGuard guard(g->device);
    rg_index *ix = new rg_index();
    ix->device = g->device;
    ix->n = n;
    ix->dim = dim;
    ix->metric = metric;
    ix->ep = g->ep;
    ix->max_degree = g->max_degree;
    ix->adj_stride = g->stride;
    ix->d_base = d_base;
    ix->owns_base = false;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, g->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_adj, n * uint64_t(g->stride) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(ix->d_adj, g->d_adj, n * uint64_t(g->stride) * sizeof(uint32_t), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_counters, 64 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ix->d_counters, 0, 64 * sizeof(uint32_t));
    if (e == cudaSuccess) ix->reserve_search_scratch();
    if (e != cudaSuccess) {
        rg_index_destroy(ix);
        return rg::fail(e == cudaErrorMemoryAllocation ? RG_ERR_OUT_OF_MEMORY : RG_ERR_CUDA, "rg_index_create_from_graph: %s",
                        cudaGetErrorString(e));
    }
    ix->sm_count = prop.multiProcessorCount;
    ix->max_smem_optin = int(prop.sharedMemPerBlockOptin);
    *out = ix;
    return RG_OK;
}


// Diagnostic: only the pivot projection prune (P1, PruneBiSearchBaseGetBase src/index_bipartite.cpp:1059-1097, 1612-1694) of
// every training query; d_lists [n_train][M_pjbp + 1] (word 0 = list length).  Deterministic given the kNN ids.
rg_status rg_build_projection_lists_device(const float *d_base, uint64_t n, uint32_t dim, int metric, const uint32_t *d_knn_ids,
                                           uint64_t n_train, uint32_t knn_k, uint32_t M_sq, uint32_t M_pjbp, uint32_t *d_lists,
                                           int device, void *cuda_stream) {
    if (!d_base || !d_knn_ids || !d_lists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_projection_lists_device: null argument");
    if (n == 0 || n >= (1ull << 31) || n_train == 0 || n_train * 2ull * M_pjbp >= (1ull << 32) || dim == 0 || dim % 8 != 0 ||
        knn_k == 0 || M_sq == 0 || M_pjbp == 0 || M_pjbp > 64)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_build_projection_lists_device: argument out of range");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg_graph g;
    const rg_status s = rg::build::build_device(d_base, n, dim, metric, d_knn_ids, n_train, knn_k, M_sq, M_pjbp, 16, &g,
                                                static_cast<cudaStream_t>(cuda_stream), d_lists);
    cudaFree(g.d_adj);
    return s;
}
}  // extern "C"
