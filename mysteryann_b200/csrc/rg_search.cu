// K1 - batched beam search over the projected graph (replaces IndexBipartite::SearchRoarGraph,
// /root/reference src/index_bipartite.cpp:2311-2420, and the OpenMP query loop of
// tests/test_search_roargraph.cpp:203-209).
//
// One warp owns one query at a time; a persistent grid pulls query indices from an atomic counter
// (the reference's schedule(dynamic,1)).  Per hop the warp
//   1. pops the closest unexpanded pool entry              (NeighborPriorityQueue::closest_unexpanded)
//   2. reads that node's fixed-stride adjacency row        (one dependent HBM read)
//   3. filters the neighbours through an exact visited set (32-bit open-addressing hash in shared memory;
//                                                           replaces VisitedList's uint16 tag array)
//   4. gathers the surviving rows HBM -> shared memory     (TMA bulk copies on an mbarrier, or cp.async),
//      double buffered in batches of `stage_rows`
//   5. scores 8 rows at a time, 4 lanes per row, in the exact FP32 operation order of the compiled
//      reference distance (16 lane accumulators, unfused main loop, fused tails; distance.h:39-89,179-223)
//   6. merges the scored candidates into the sorted pool   (NeighborPriorityQueue::insert semantics)
// Within a hop the order of insertion does not change the final pool (bounded sorted set under the
// strict order (distance,id)), so steps 3-6 are batch operations with bit-identical results.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rg_index.cuh"

namespace rg {

enum { kCntWork = 0, kCntNotEnough = 1, kCntOverflow = 2, kCntFatal = 3, kCntWork2 = 4 };
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct SearchParams {
    const float *base;
    const uint32_t *adj;
    const float *queries;
    uint32_t *ids;
    float *dists;
    uint32_t *cmps;
    uint32_t *hops;
    uint32_t *counters;
    uint32_t *overflow_list;   // primary pass appends here; fallback pass reads from here
    uint32_t *ghash;           // global visited-hash slabs (kGlobalHash only)
    uint32_t nq;               // primary: number of queries; fallback: unused (count read from counters)
    uint32_t dim, adj_stride, ep, k, L;
    uint32_t hash_log2, hash_limit;
    uint32_t stage_rows;       // rows per staging buffer (multiple of 8)
    uint32_t stage_bufs;       // 1 = single buffer (more resident warps), 2 = double buffered
    uint32_t row_stride;       // floats between staged rows; row_stride % 32 == 16 -> conflict-free float4 reads
    uint32_t cand_cap;         // capacity of the candidate arrays (>= adj_stride)
    uint32_t chunk_magic;      // ceil(2^32 / (dim/4)) for the cp.async index split
    uint32_t fallback;         // 1 = second pass over overflow_list with the big global table
    // byte offsets inside the per-warp shared-memory slice
    uint32_t smem_per_warp, off_pool, off_cid, off_ckey, off_hash, off_stage, off_mbar;
};

// ---- exact visited set --------------------------------------------------------------------------
__device__ __forceinline__ bool visited_test_and_set(uint32_t *table, uint32_t log2, uint32_t id) {
    const uint32_t mask = (1u << log2) - 1u;
    uint32_t slot = (id * 0x9E3779B1u) >> (32 - log2);
    for (;;) {
        uint32_t old = atomicCAS(table + slot, kEmpty, id);
        if (old == kEmpty) return true;   // first visit
        if (old == id) return false;      // already visited
        slot = (slot + 1) & mask;
    }
}

// ---- distance of 8 rows per warp, 4 lanes per row, reference operation order ---------------------
// Lane t (0..3) of a group owns AVX lanes 4t..4t+3 of the reference's 16-lane accumulator.
template <bool kIP>
__device__ __forceinline__ void main_step(float4 &acc, const float4 v, const float4 q) {  // vmulps + vaddps
    if (kIP) {
        acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, q.x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, q.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, q.z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, q.w));
    } else {
        const float dx = __fsub_rn(v.x, q.x), dy = __fsub_rn(v.y, q.y), dz = __fsub_rn(v.z, q.z),
                    dw = __fsub_rn(v.w, q.w);
        acc.x = __fadd_rn(acc.x, __fmul_rn(dx, dx));
        acc.y = __fadd_rn(acc.y, __fmul_rn(dy, dy));
        acc.z = __fadd_rn(acc.z, __fmul_rn(dz, dz));
        acc.w = __fadd_rn(acc.w, __fmul_rn(dw, dw));
    }
}
template <bool kIP>
__device__ __forceinline__ void fused_step(float4 &m, const float4 v, const float4 q) {  // vfmadd231ps
    if (kIP) {
        m.x = __fmaf_rn(v.x, q.x, m.x);
        m.y = __fmaf_rn(v.y, q.y, m.y);
        m.z = __fmaf_rn(v.z, q.z, m.z);
        m.w = __fmaf_rn(v.w, q.w, m.w);
    } else {
        const float dx = __fsub_rn(v.x, q.x), dy = __fsub_rn(v.y, q.y), dz = __fsub_rn(v.z, q.z),
                    dw = __fsub_rn(v.w, q.w);
        m.x = __fmaf_rn(dx, dx, m.x);
        m.y = __fmaf_rn(dy, dy, m.y);
        m.z = __fmaf_rn(dz, dz, m.z);
        m.w = __fmaf_rn(dw, dw, m.w);
    }
}
// folds 16 -> 8 (AVX lane l+8 lives two CUDA lanes up), the fused 8-wide tail, 8 -> 4, (x0+x1)+(x2+x3)
template <bool kIP>
__device__ __forceinline__ float finish_distance(const float4 acc, bool tail8, const float4 vt, const float4 qt,
                                                  uint32_t t) {
    float4 m;
    m.x = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.x, 2), acc.x);
    m.y = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.y, 2), acc.y);
    m.z = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.z, 2), acc.z);
    m.w = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.w, 2), acc.w);
    if (tail8 && t < 2) fused_step<kIP>(m, vt, qt);
    float4 f;
    f.x = __fadd_rn(__shfl_down_sync(0xffffffffu, m.x, 1), m.x);
    f.y = __fadd_rn(__shfl_down_sync(0xffffffffu, m.y, 1), m.y);
    f.z = __fadd_rn(__shfl_down_sync(0xffffffffu, m.z, 1), m.z);
    f.w = __fadd_rn(__shfl_down_sync(0xffffffffu, m.w, 1), m.w);
    const float r = __fadd_rn(__fadd_rn(f.x, f.y), __fadd_rn(f.z, f.w));
    return kIP ? -r : r;
}

// row staged in shared memory (gather modes 1, 2)
template <bool kIP>
__device__ __forceinline__ float lane_exact_distance(const float4 *__restrict__ rp, const float4 *__restrict__ qp,
                                                      uint32_t n16, bool tail8, uint32_t t) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (uint32_t s = 0; s < n16; ++s) main_step<kIP>(acc, rp[4 * s], qp[4 * s]);
    float4 vt = make_float4(0.f, 0.f, 0.f, 0.f), qt = vt;
    if (tail8 && t < 2) {
        vt = rp[4 * n16];
        qt = qp[4 * n16];
    }
    return finish_distance<kIP>(acc, tail8, vt, qt, t);
}

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {  // read-once row data: keep it out of L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// row read straight from HBM into registers (gather mode 3): kChunk 128-bit loads per lane are in flight
// before the first one is consumed
template <bool kIP, int kChunk>
__device__ __forceinline__ float lane_exact_distance_global(const float4 *__restrict__ gp,
                                                             const float4 *__restrict__ qp, uint32_t n16, bool tail8,
                                                             uint32_t t) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 vt = make_float4(0.f, 0.f, 0.f, 0.f), qt = vt;
    if (tail8 && t < 2) vt = ldg_stream(gp + 4 * n16);
    for (uint32_t s0 = 0; s0 < n16; s0 += kChunk) {
        float4 v[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
            if (s0 + j < n16) v[j] = ldg_stream(gp + 4 * (s0 + j));
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
            if (s0 + j < n16) main_step<kIP>(acc, v[j], qp[4 * (s0 + j)]);
    }
    if (tail8 && t < 2) qt = qp[4 * n16];
    return finish_distance<kIP>(acc, tail8, vt, qt, t);
}

// kGather: 1 = cp.async (LDGSTS 16 B per lane) into shared memory, 2 = TMA bulk copy (one UBLKCP per row) into shared
// memory on an mbarrier, 3 = straight into registers (LDG.128, no staging buffer -> more resident warps)
template <bool kIP, int kGather, bool kGlobalHash>
__global__ void __launch_bounds__(256) rg_search_kernel(const SearchParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t grp = lane >> 2, t = lane & 3;
    unsigned char *ws = smem_raw + size_t(warp) * p.smem_per_warp;
    float *s_query = reinterpret_cast<float *>(ws);
    uint64_t *s_pool = reinterpret_cast<uint64_t *>(ws + p.off_pool);
    uint32_t *s_cid = reinterpret_cast<uint32_t *>(ws + p.off_cid);
    uint64_t *s_ckey = reinterpret_cast<uint64_t *>(ws + p.off_ckey);
    float *s_stage = reinterpret_cast<float *>(ws + p.off_stage);
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(ws + p.off_mbar);
    uint32_t *hash = kGlobalHash
                         ? p.ghash + ((size_t(blockIdx.x) * (blockDim.x >> 5) + warp) << p.hash_log2)
                         : reinterpret_cast<uint32_t *>(ws + p.off_hash);

    const uint32_t dim = p.dim, n16 = dim >> 4;
    const bool tail8 = (dim & 15u) != 0;
    const uint32_t cpr = dim >> 2;  // 16-byte chunks per row
    const uint32_t L = p.L, BR = p.stage_rows, RS = p.row_stride;
    uint32_t phase_bits = 0;  // mbarrier parity of the two staging buffers

    if (kGather == 2) {
        if (lane == 0) {
            mbar_init(&s_mbar[0], 1);
            mbar_init(&s_mbar[1], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }

    // ---- gather helpers: stage candidate rows [c0, c0 + rows) into buffer `buf` ------------------
    auto issue = [&](uint32_t c0, uint32_t rows, uint32_t buf) {
        float *dst0 = s_stage + size_t(buf) * BR * RS;
        if (kGather == 2) {
            if (lane == 0) mbar_arrive_expect_tx(&s_mbar[buf], rows * dim * 4u);
            __syncwarp();
            for (uint32_t r = lane; r < rows; r += 32) {
                const uint32_t id = s_cid[c0 + r];
                bulk_g2s(dst0 + size_t(r) * RS, p.base + size_t(id) * dim, dim * 4u, &s_mbar[buf]);
            }
        } else {
            const uint32_t total = rows * cpr;
            for (uint32_t idx = lane; idx < total; idx += 32) {
                const uint32_t r = __umulhi(idx, p.chunk_magic);
                const uint32_t c = idx - r * cpr;
                const uint32_t id = s_cid[c0 + r];
                cp_async16(dst0 + size_t(r) * RS + 4 * c, p.base + size_t(id) * dim + 4 * c);
            }
            cp_async_commit();
        }
    };
    auto wait_buf = [&](uint32_t buf, bool another_in_flight) {
        if (kGather == 2) {
            mbar_wait(&s_mbar[buf], (phase_bits >> buf) & 1u);
            phase_bits ^= (1u << buf);
        } else {
            if (another_in_flight) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncwarp();
        }
    };
    // scores s_cid[0..ncand) -> s_ckey[0..ncand)
    auto score = [&](uint32_t ncand) {
        if (kGather == 3) {
            const float4 *qp = reinterpret_cast<const float4 *>(s_query) + t;
            for (uint32_t r0 = 0; r0 < ncand; r0 += 8) {
                const uint32_t r = r0 + grp;
                const bool valid = r < ncand;
                const uint32_t id = s_cid[valid ? r : ncand - 1];
                const float4 *gp = reinterpret_cast<const float4 *>(p.base + size_t(id) * dim) + t;
                const float d = lane_exact_distance_global<kIP, 8>(gp, qp, n16, tail8, t);
                if (valid && t == 0) s_ckey[r] = make_key(d, id);
            }
            __syncwarp();
            return;
        }
        const uint32_t nb = (ncand + BR - 1) / BR;
        if (p.stage_bufs == 1) {  // overlap comes from the other resident warps
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t c0 = b * BR;
                const uint32_t rows = min(BR, ncand - c0);
                issue(c0, rows, 0);
                wait_buf(0, false);
                for (uint32_t r0 = 0; r0 < rows; r0 += 8) {
                    const uint32_t r = r0 + grp;
                    const bool valid = r < rows;
                    const uint32_t rr = valid ? r : rows - 1;
                    const float4 *rp = reinterpret_cast<const float4 *>(s_stage + size_t(rr) * RS) + t;
                    const float4 *qp = reinterpret_cast<const float4 *>(s_query) + t;
                    const float d = lane_exact_distance<kIP>(rp, qp, n16, tail8, t);
                    if (valid && t == 0) s_ckey[c0 + r] = make_key(d, s_cid[c0 + r]);
                }
                __syncwarp();
            }
            return;
        }
        issue(0, min(BR, ncand), 0);
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t c0 = b * BR;
            const uint32_t rows = min(BR, ncand - c0);
            const bool more = (b + 1 < nb);
            if (more) issue(c0 + BR, min(BR, ncand - c0 - BR), (b + 1) & 1);
            wait_buf(b & 1, more);
            const float *buf0 = s_stage + size_t(b & 1) * BR * RS;
            for (uint32_t r0 = 0; r0 < rows; r0 += 8) {
                const uint32_t r = r0 + grp;
                const bool valid = r < rows;
                const uint32_t rr = valid ? r : rows - 1;
                const float4 *rp = reinterpret_cast<const float4 *>(buf0 + size_t(rr) * RS) + t;
                const float4 *qp = reinterpret_cast<const float4 *>(s_query) + t;
                const float d = lane_exact_distance<kIP>(rp, qp, n16, tail8, t);
                if (valid && t == 0) s_ckey[c0 + r] = make_key(d, s_cid[c0 + r]);
            }
            __syncwarp();  // all reads of this buffer done before it is refilled two batches later
        }
    };

    uint32_t size = 0, cur = 0;
    uint64_t tail = ~0ull;  // (distance,id) of the last entry once the pool is full, else +inf (warp-uniform)
    // NeighborPriorityQueue::insert (neighbor.h:150-183) for one key, executed by the whole warp
    auto pool_insert = [&](uint64_t key) {
        if (key >= tail) return;  // full pool and worse than its last entry, or the very same (distance,id)
        uint32_t pos = 0;
        bool dup = false;
        for (uint32_t i0 = 0; i0 < size; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint64_t e = (i < size) ? s_pool[i] : ~0ull;
            pos += __popc(__ballot_sync(0xffffffffu, e < key));
            dup |= __any_sync(0xffffffffu, (e & ~1ull) == key);
        }
        if (dup) return;  // "Make sure the same id isn't inserted into the set" (neighbor.h:161)
        const int last = (size < L) ? int(size) : int(L) - 1;  // highest index written by the shift
        for (int hi = last; hi > int(pos); hi -= 32) {
            const int i = hi - int(lane);
            uint64_t v = 0;
            if (i > int(pos)) v = s_pool[i - 1];
            __syncwarp();
            if (i > int(pos)) s_pool[i] = v;
            __syncwarp();
        }
        if (lane == 0) s_pool[pos] = key;
        __syncwarp();
        if (size < L) ++size;
        if (pos < cur) cur = pos;
        if (size == L) tail = s_pool[L - 1] & ~1ull;
    };
    // all scored candidates of a hop: lanes test their keys against the tail in parallel, survivors are inserted
    // one by one (the tail only tightens, so a key rejected here would be rejected by insert() as well)
    auto merge = [&](uint32_t ncand) {
        for (uint32_t c0 = 0; c0 < ncand; c0 += 32) {
            const uint64_t key = (c0 + lane < ncand) ? s_ckey[c0 + lane] : ~0ull;
            uint32_t m = __ballot_sync(0xffffffffu, key < tail);
            while (m) {
                const uint32_t b = __ffs(m) - 1;
                pool_insert(__shfl_sync(0xffffffffu, key, b));
                m &= m - 1;
                m &= __ballot_sync(0xffffffffu, key < tail);
            }
        }
    };

    for (;;) {
        // ---- next query (schedule(dynamic,1)) ---------------------------------------------------
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&p.counters[p.fallback ? kCntWork2 : kCntWork], 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        const uint32_t nwork = p.fallback ? min(p.counters[kCntOverflow], p.nq) : p.nq;
        if (w >= nwork) break;
        const uint32_t qi = p.fallback ? p.overflow_list[w] : w;

        {   // query -> shared memory; clear the visited set
            const float4 *src = reinterpret_cast<const float4 *>(p.queries + size_t(qi) * dim);
            float4 *dst = reinterpret_cast<float4 *>(s_query);
            for (uint32_t i = lane; i < cpr; i += 32) dst[i] = src[i];
            uint4 *h4 = reinterpret_cast<uint4 *>(hash);
            const uint4 e4 = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            for (uint32_t i = lane; i < (1u << (p.hash_log2 - 2)); i += 32) h4[i] = e4;
        }
        __syncwarp();

        size = 0;
        cur = 0;
        tail = ~0ull;
        uint32_t cmps = 0, hops = 0, nvis = 0;
        bool overflow = false;

        // entry point: scored and inserted, NOT marked visited (src/index_bipartite.cpp:2337-2353)
        if (lane == 0) s_cid[0] = p.ep;
        __syncwarp();
        score(1);
        pool_insert(s_ckey[0]);

        while (cur < size) {
            // closest_unexpanded (neighbor.h:185-192)
            const uint64_t ckey = s_pool[cur];
            const uint32_t cur_id = key_id(ckey);
            if (lane == 0) s_pool[cur] = ckey | 1ull;
            __syncwarp();
            {
                uint32_t c = cur + 1;
                uint32_t next = size;
                while (c < size) {
                    const uint32_t i = c + lane;
                    const bool unexp = (i < size) && ((s_pool[i] & 1ull) == 0);
                    const uint32_t m = __ballot_sync(0xffffffffu, unexp);
                    if (m) {
                        next = c + (__ffs(m) - 1);
                        break;
                    }
                    c += 32;
                }
                cur = next;
            }
            ++hops;

            if (nvis + p.adj_stride > p.hash_limit) {  // visited set may fill up: hand over to the big-table pass
                overflow = true;
                break;
            }
            // adjacency row, visited filter, compaction into s_cid (adjacency order preserved)
            const uint32_t *row = p.adj + size_t(cur_id) * p.adj_stride;
            uint32_t ncand = 0;
            uint32_t deg = 0;
            uint32_t wreg[3];  // the whole row (<= 96 words) is requested at once: one DRAM round trip
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) {
                const uint32_t idx = c * 32 + lane;
                wreg[c] = (idx < p.adj_stride) ? __ldg(row + idx) : kEmpty;
            }
            for (uint32_t c0 = 0; c0 < p.adj_stride; c0 += 32) {
                const uint32_t idx = c0 + lane;
                uint32_t word;
                if (c0 == 0) word = wreg[0];
                else if (c0 == 32) word = wreg[1];
                else if (c0 == 64) word = wreg[2];
                else word = (idx < p.adj_stride) ? __ldg(row + idx) : kEmpty;
                if (c0 == 0) deg = __shfl_sync(0xffffffffu, word, 0);
                if (c0 > deg) break;  // warp-uniform
                const bool is_nbr = (idx >= 1) && (idx <= deg);
                bool fresh = false;
                if (is_nbr) fresh = visited_test_and_set(hash, p.hash_log2, word);
                const uint32_t m = __ballot_sync(0xffffffffu, fresh);
                if (fresh) s_cid[ncand + __popc(m & lanemask_lt())] = word;
                ncand += __popc(m);
            }
            __syncwarp();
            cmps += ncand;
            nvis += ncand;
            if (ncand == 0) continue;
            score(ncand);
            // a re-scored entry point lands here too; insert() drops it as a duplicate (neighbor.h:161) or as
            // worse than the tail (neighbor.h:151), exactly like the reference
            merge(ncand);
        }

        if (overflow) {
            if (!p.fallback) {
                if (lane == 0) {
                    const uint32_t pos = atomicAdd(&p.counters[kCntOverflow], 1u);
                    p.overflow_list[pos] = qi;
                }
                continue;
            }
            if (lane == 0) atomicAdd(&p.counters[kCntFatal], 1u);
            size = 0;  // falls through to the "not enough results" fill
        }
        // results (src/index_bipartite.cpp:2408-2419)
        if (size < p.k) {
            if (lane == 0 && !overflow) atomicAdd(&p.counters[kCntNotEnough], 1u);
            for (uint32_t i = lane; i < p.k; i += 32) {
                p.ids[size_t(qi) * p.k + i] = kEmpty;
                p.dists[size_t(qi) * p.k + i] = 0.f;
            }
        } else {
            for (uint32_t i = lane; i < p.k; i += 32) {
                const uint64_t e = s_pool[i];
                p.ids[size_t(qi) * p.k + i] = key_id(e);
                p.dists[size_t(qi) * p.k + i] = key_dist(e);
            }
        }
        if (lane == 0) {
            if (p.cmps) p.cmps[qi] = cmps;
            if (p.hops) p.hops[qi] = hops;
        }
        __syncwarp();
    }
}

// ---- host side: geometry + launch ---------------------------------------------------------------
struct Geometry {
    SearchParams p;
    int warps, ctas_per_sm, gather;
    bool global_hash;
    size_t smem_bytes;
};

static uint32_t round_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// experiment knobs (tuning sweeps only; 0 / unset = automatic)
static int env_int(const char *name, int def) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : def;
}

static uint32_t auto_hash_log2(uint32_t L) {
    // expected worst-case visited nodes per query ~ 1000 + 30 L (max cmps 1371 @L=10 ...
    // 13702 @L=500); outliers take the exact global-table fallback, so this only affects speed.
    const double want = (1000.0 + 30.0 * L) / 0.8;
    uint32_t lg = 10;
    while ((1u << lg) < want && lg < 22) ++lg;
    return lg;
}

static rg_status make_geometry(const rg_index *ix, uint32_t k, uint32_t L, bool fallback, Geometry *g) {
    SearchParams &p = g->p;
    memset(&p, 0, sizeof(p));
    p.dim = ix->dim;
    p.adj_stride = ix->adj_stride;
    p.ep = ix->ep;
    p.k = k;
    p.L = L;
    p.cand_cap = round_up(ix->adj_stride, 32);
    const uint32_t cpr = ix->dim / 4;
    p.chunk_magic = uint32_t((0x100000000ull + cpr - 1) / cpr);
    // smallest rs >= dim with rs % 32 == 16: the two rows a quarter-warp reads with LDS.128 fall in
    // different halves of the 32 banks
    uint32_t rs = (ix->dim % 32 <= 16) ? ix->dim - ix->dim % 32 + 16 : ix->dim - ix->dim % 32 + 48;
    p.row_stride = rs;
    uint32_t br = ix->cfg_stage_rows ? uint32_t(ix->cfg_stage_rows) : 8u;  // 8 rows x 2 buffers measured best (profiles/)
    p.stage_rows = br;
    g->gather = ix->cfg_gather ? ix->cfg_gather : 2;

    uint32_t hl = ix->cfg_hash_log2 ? uint32_t(ix->cfg_hash_log2) : auto_hash_log2(L);
    p.fallback = fallback ? 1u : 0u;
    g->global_hash = fallback || hl > 15 || env_int("RG_K1_GHASH", 0) != 0;
    if (fallback) hl = std::min<uint32_t>(22u, std::max<uint32_t>(16u, hl + 3));
    p.hash_log2 = hl;
    p.hash_limit = uint32_t((uint64_t(1) << hl) * 85 / 100);

    uint32_t off = round_up(ix->dim * 4, 128);
    p.off_pool = off;
    off += round_up((L + 1) * 8, 128);
    p.off_cid = off;
    off += round_up(p.cand_cap * 4, 128);
    p.off_ckey = off;
    off += round_up(p.cand_cap * 8, 128);
    p.off_hash = off;
    if (!g->global_hash) off += (4u << hl);
    p.off_stage = off;
    p.stage_bufs = ix->cfg_stage_bufs ? uint32_t(ix->cfg_stage_bufs) : uint32_t(env_int("RG_K1_BUFS", 2));
    if (g->gather != 3) off += round_up(p.stage_bufs * br * rs * 4, 128);
    p.off_mbar = off;
    off += 128;
    p.smem_per_warp = off;

    const size_t sm_budget = 227 * 1024, cta_overhead = 1024;
    if (size_t(off) > size_t(ix->max_smem_optin))
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq=%u needs %u bytes of shared memory per query (max %d)", L, off,
                        ix->max_smem_optin);
    int best_w = 1, best_c = 1, best_res = 0;
    const int w_opts[] = {8, 4, 2, 1};
    for (int w : w_opts) {
        if (fallback && w > 2) continue;  // few heavy queries, big global tables: keep the slab count small
        if (!fallback && ix->cfg_warps && w != ix->cfg_warps) continue;
        const size_t cta = size_t(w) * off;
        if (cta > size_t(ix->max_smem_optin)) continue;
        int c = int(sm_budget / (cta + cta_overhead));
        c = std::min(c, 2048 / (w * 32));
        c = std::min(c, 32);
        if (ix->cfg_ctas) c = std::min(c, ix->cfg_ctas);
        if (c < 1) continue;
        if (w * c > best_res) {
            best_res = w * c;
            best_w = w;
            best_c = c;
        }
    }
    if (!fallback && ix->cfg_warps && best_res == 0) {
        best_w = ix->cfg_warps;
        best_c = 1;
        if (size_t(best_w) * off > size_t(ix->max_smem_optin))
            return rg::fail(RG_ERR_INVALID_ARGUMENT, "warps_per_cta=%d does not fit in shared memory at L_pq=%u", best_w, L);
    }
    g->warps = best_w;
    g->ctas_per_sm = best_c;
    g->smem_bytes = size_t(best_w) * off;
    return RG_OK;
}

template <bool kIP, int kGather, bool kGlobalHash>
static cudaError_t launch_one(const Geometry &g, int grid, cudaStream_t st) {
    auto kern = rg_search_kernel<kIP, kGather, kGlobalHash>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.smem_bytes));
    if (e != cudaSuccess) return e;
    kern<<<grid, g.warps * 32, g.smem_bytes, st>>>(g.p);
    return cudaGetLastError();
}

static cudaError_t launch(const Geometry &g, bool ip, int grid, cudaStream_t st) {
    if (g.gather == 3) {
        if (g.global_hash) return ip ? launch_one<true, 3, true>(g, grid, st) : launch_one<false, 3, true>(g, grid, st);
        return ip ? launch_one<true, 3, false>(g, grid, st) : launch_one<false, 3, false>(g, grid, st);
    }
    if (g.global_hash) {
        if (g.gather == 2) return ip ? launch_one<true, 2, true>(g, grid, st) : launch_one<false, 2, true>(g, grid, st);
        return ip ? launch_one<true, 1, true>(g, grid, st) : launch_one<false, 1, true>(g, grid, st);
    }
    if (g.gather == 2) return ip ? launch_one<true, 2, false>(g, grid, st) : launch_one<false, 2, false>(g, grid, st);
    return ip ? launch_one<true, 1, false>(g, grid, st) : launch_one<false, 1, false>(g, grid, st);
}

static rg_status ensure(void **ptr, uint64_t *cap, uint64_t want, size_t elem) {
    if (*cap >= want && *ptr) return RG_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    RG_CUDA_OK(cudaMalloc(ptr, want * elem));
    *cap = want;
    return RG_OK;
}

static rg_status search_device(rg_index *ix, const float *d_queries, uint64_t nq, uint32_t k, uint32_t L,
                               uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops,
                               uint32_t *d_status, cudaStream_t st) {
    if (!ix || !d_queries || !d_ids || !d_dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search: null argument");
    if (k == 0 || L == 0 || k > L) return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq must greater or equal than k (k=%u, L_pq=%u)", k, L);
    if (L > 16384) return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq=%u too large (max 16384)", L);
    if (nq >= (1ull << 32)) return rg::fail(RG_ERR_INVALID_ARGUMENT, "too many queries in one batch");
    if (nq == 0) return RG_OK;

    Geometry g1, g2;
    rg_status s = make_geometry(ix, k, L, false, &g1);
    if (s != RG_OK) return s;
    s = make_geometry(ix, k, L, true, &g2);
    if (s != RG_OK) return s;
    const bool ip = ix->metric != RG_METRIC_L2;

    // scratch: overflow list (one slot per query) and global hash slabs
    s = ensure((void **)&ix->d_overflow_list, &ix->overflow_cap, nq, sizeof(uint32_t));
    if (s != RG_OK) return s;
    const int grid1 = int(std::min<uint64_t>((nq + g1.warps - 1) / g1.warps, uint64_t(ix->sm_count) * g1.ctas_per_sm));
    // fallback pass: few, heavy queries; one CTA per SM
    const int grid2 = ix->sm_count;
    uint64_t need_hash = uint64_t(grid2) * g2.warps << g2.p.hash_log2;
    if (g1.global_hash) need_hash = std::max(need_hash, uint64_t(grid1) * g1.warps << g1.p.hash_log2);
    s = ensure((void **)&ix->d_ghash, &ix->ghash_words, need_hash, sizeof(uint32_t));
    if (s != RG_OK) return s;

    for (Geometry *g : {&g1, &g2}) {
        g->p.base = ix->d_base;
        g->p.adj = ix->d_adj;
        g->p.queries = d_queries;
        g->p.ids = d_ids;
        g->p.dists = d_dists;
        g->p.cmps = d_cmps;
        g->p.hops = d_hops;
        g->p.counters = ix->d_counters;
        g->p.overflow_list = ix->d_overflow_list;
        g->p.ghash = ix->d_ghash;
        g->p.nq = uint32_t(nq);
    }
    RG_CUDA_OK(cudaMemsetAsync(ix->d_counters, 0, 8 * sizeof(uint32_t), st));
    RG_CUDA_OK(launch(g1, ip, grid1, st));
    ix->launches++;
    RG_CUDA_OK(launch(g2, ip, grid2, st));  // exits immediately when nothing overflowed
    ix->launches++;
    if (d_status) {
        RG_CUDA_OK(cudaMemcpyAsync(d_status, ix->d_counters + kCntNotEnough, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        RG_CUDA_OK(cudaMemcpyAsync(d_status + 1, ix->d_counters + kCntFatal, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return RG_OK;
}

}  // namespace rg

extern "C" {

rg_status rg_search_batch_device(rg_index *ix, const float *d_queries, uint64_t nq, uint32_t k, uint32_t L,
                                 uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops,
                                 uint32_t *d_status, void *cuda_stream) {
    if (!ix) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_batch_device: null index");
    rg::DeviceGuard guard(ix->device);
    return rg::search_device(ix, d_queries, nq, k, L, d_ids, d_dists, d_cmps, d_hops, d_status,
                             static_cast<cudaStream_t>(cuda_stream));
}

rg_status rg_search_batch(rg_index *ix, const float *queries, uint64_t nq, uint32_t k, uint32_t L, uint32_t *ids,
                          float *dists, uint32_t *cmps, uint32_t *hops) {
    if (!ix || !queries || !ids || !dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_batch: null argument");
    if (nq == 0) return RG_OK;
    rg::DeviceGuard guard(ix->device);
    rg_status s;
    if ((s = rg::ensure((void **)&ix->d_queries, &ix->queries_cap, nq * ix->dim, sizeof(float))) != RG_OK) return s;
    if (ix->res_cap < nq * k) {
        cudaFree(ix->d_ids);
        cudaFree(ix->d_dists);
        ix->d_ids = nullptr;
        ix->d_dists = nullptr;
        ix->res_cap = 0;
        RG_CUDA_OK(cudaMalloc(&ix->d_ids, nq * k * sizeof(uint32_t)));
        RG_CUDA_OK(cudaMalloc(&ix->d_dists, nq * k * sizeof(float)));
        ix->res_cap = nq * k;
    }
    if (ix->stat_cap < nq + 2) {
        cudaFree(ix->d_cmps);
        cudaFree(ix->d_hops);
        ix->d_cmps = ix->d_hops = nullptr;
        ix->stat_cap = 0;
        RG_CUDA_OK(cudaMalloc(&ix->d_cmps, (nq + 2) * sizeof(uint32_t)));
        RG_CUDA_OK(cudaMalloc(&ix->d_hops, (nq + 2) * sizeof(uint32_t)));
        ix->stat_cap = nq + 2;
    }
    cudaStream_t st = ix->stream;
    RG_CUDA_OK(cudaMemcpyAsync(ix->d_queries, queries, nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    uint32_t *d_status = ix->d_cmps + nq;  // two spare words behind the cmps array
    s = rg::search_device(ix, ix->d_queries, nq, k, L, ix->d_ids, ix->d_dists, ix->d_cmps, ix->d_hops, d_status, st);
    if (s != RG_OK) return s;
    uint32_t status[2] = {0, 0};
    RG_CUDA_OK(cudaMemcpyAsync(ids, ix->d_ids, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(dists, ix->d_dists, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (cmps) RG_CUDA_OK(cudaMemcpyAsync(cmps, ix->d_cmps, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (hops) RG_CUDA_OK(cudaMemcpyAsync(hops, ix->d_hops, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(status, d_status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    if (status[1]) return rg::fail(RG_ERR_INTERNAL, "visited set overflow in %u queries (L_pq=%u)", status[1], L);
    if (status[0]) {
        // report the first short query like the reference's message (src/index_bipartite.cpp:2408-2412)
        return rg::fail(RG_ERR_NOT_ENOUGH_RESULTS, "not enough results: fewer than %u pool entries in %u queries, expected: %u", k, status[0], k);
    }
    return RG_OK;
}

}  // extern "C"
