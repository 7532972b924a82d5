// K1 - batched beam search over the projected graph (replaces IndexBipartite::SearchRoarGraph,
// /root/reference src/index_bipartite.cpp:2311-2420, and the OpenMP query loop of
// tests/test_search_roargraph.cpp:203-209).
//
// One CTA of W warps (W = 1..8, default 2) owns one query at a time; a persistent grid pulls query indices from an
// atomic counter (the reference's schedule(dynamic,1)).  Per hop the CTA
//   1. takes the closest unexpanded pool entry             (NeighborPriorityQueue::closest_unexpanded)
//   2. reads that node's fixed-stride adjacency row, one word per thread (neighbour j belongs to warp j % W); the row of
//      the entry that will most likely be expanded NEXT is read in the same round trip: its visited-hash slots are
//      prefetched into L2 and, when the speculation holds, its words are already in registers one hop later
//   3. filters the neighbours through an exact visited set (open-addressing hash, atomicCAS; by default a slab per CTA
//      in global memory - 16-bit quotient entries when the id range allows, else 32-bit keys - optionally shared
//      memory; replaces VisitedList's uint16 tag array)
//   4. every warp gathers the rows of ITS surviving neighbours HBM -> shared memory (TMA bulk copies on an mbarrier, or
//      cp.async) in batches of `stage_rows`
//   5. and scores them 8 rows at a time, 4 lanes per row, in the exact FP32 operation order of the compiled
//      reference distance (16 lane accumulators, unfused main loop, fused tails; distance.h:39-89,179-223);
//      keys that cannot enter the pool (>= its last entry once full) are dropped on the spot
//   6. all threads merge the hop's candidates into the sorted pool IN PLACE: candidates are ranked against the pool
//      (binary search) and against each other, pool entries behind the first insertion point shift right by the number
//      of candidates in front of them (binary search in the sorted candidates), top-down in chunks of one entry per
//      thread - O((L + C) log C / T) per thread, same final state as NeighborPriorityQueue::insert one by one.
// Within a hop the order of insertion does not change the final pool (bounded sorted set under the strict order
// (distance,id)), so steps 2-6 are batch operations with bit-identical results, cmps and hops included.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rg_distance.cuh"
#include "rg_index.cuh"

namespace rg {

enum { kCntWork = 0, kCntNotEnough = 1, kCntOverflow = 2, kCntFatal = 3, kCntWork2 = 4, kCntException = 5 };
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kMaxWarps = 8;
#ifndef RG_K1_MIN_CTAS
#define RG_K1_MIN_CTAS 3  // 256-thread CTAs per SM the register allocation must allow: caps K1 at 80 registers per thread
                          // (64 costs 4-6 % at L_pq = 55, 128 loses a third of the resident queries: profiles/r02_k1_ab_*.txt)
#endif
// shared control words
enum { kCtlWork = 0, kCtlNvis = 1, kCtlHashFull = 2, kCtlHop0 = 4 /* 2 x {ncand, ndup, minlo, curpos} */,
       kCtlFresh = 12 /* ids in the CTA-wide gather list */, kCtlBatch = 13 /* next batch of it */,
       kCtlBeat0 = 14 /* 2 x: a candidate of the hop beat the node speculated on */,
       kCtlPub0 = 16 /* per warp: 1 + length of its gather list once it is complete */, kCtlNext0 = 24 /* per warp: next batch */ };
// visited-set flavours
enum { kHashShared = 0, kHashGlobal32 = 1, kHashGlobal16 = 2, kHashBucket16 = 3, kHashBucket32 = 4 };
// kHashBucket*: no atomics.  The slab is an array of buckets (16 B = 8 x 16-bit quotient entries, or 32 B = 8 x 32-bit
// ids), filled front to back.  Warp w of the CTA owns the contiguous bucket range whose home buckets satisfy
// floor(home * W / buckets) == w and is the only one that ever reads or writes it, so a lookup is ONE vector load of the
// home bucket (ld.global.cg), the decision is taken in registers, simultaneous inserts of one hop are arbitrated inside
// the warp (match.any on the bucket index, rank = slot), and the insert is a plain store nobody waits for.  Because a
// load has no side effect it can be issued a hop AHEAD for the node the search will most likely expand next: when that
// speculation holds (80 % of the hops at L_pq = 500) the visited filter costs no memory round trip at all.
#ifndef RG_K1_SPEC_SECTOR_REGS
#define RG_K1_SPEC_SECTOR_REGS 1  // 16-bit buckets: keep the speculated node's bucket contents in registers (0: L2 prefetch only)
#endif

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
    uint32_t *ghash;           // global visited-hash slabs, one per CTA (global flavours only)
    uint32_t nq;               // primary: number of queries; fallback: unused (count read from counters)
    uint32_t dim, adj_stride, ep, k, L;
    uint32_t hash_log2, hash_limit;
    // 16-bit quotient entries (kHashGlobal16): x = (id * h16_mult) mod 2^h16_bits is a bijection of the id range; the home
    // slot is its top hash_log2 bits, the entry stores the remaining h16_rbits bits and the probe displacement
    uint32_t h16_bits, h16_rbits, h16_dbits, h16_maxd, h16_mult;
    uint32_t nb, nb_per_warp;  // kHashBucket*: buckets per query (any multiple of W) and per warp; h16_* then describe the
                               // bucket-level quotient (h16_rbits = entry bits taken from x)
    uint32_t slab_words;       // kHashBucket*: 32-bit words per slab
    uint32_t stage_rows;       // rows per warp staging buffer (multiple of 8)
    uint32_t row_stride;       // floats between staged rows; row_stride % 32 == 16 -> conflict-free float4 reads
    uint32_t chunk_magic;      // ceil(2^32 / (dim/4)) for the cp.async index split
    uint32_t fallback;         // 1 = second pass over overflow_list with the big global table
    uint32_t l2_hint;          // bit 0: base-row gathers evict_first (bit 1, host side: visited-hash slabs persist in L2)
    uint32_t adj_prefetch;     // bit 0: speculate on the next unexpanded pool entry (adjacency row read ahead, its hash
                               // slots prefetched into L2); bit 1: L2-prefetch the adjacency row of every scored
                               // candidate that beats it (it will be expanded first)
    // build mode (kBuild, SearchProjectionGraphInternal src/index_bipartite.cpp:1279-1350): query w is base row
    // node_lo + w, that node is never scored, the entry point is marked visited, and the EXPANDED nodes are recorded
    uint32_t shared_batches;   // 1 = the hop's unvisited ids go to one CTA-wide list and the warps pull batches of stage_rows from it
    uint32_t steal_batches;    // 1 = per-warp lists; a warp that has finished its own takes batches of the others' (no extra barrier)
    uint32_t early_issue;      // 1 = when the read-ahead prediction holds, filter + first gather of the next hop are issued BEFORE the merge
    uint32_t early_row0;       // ... into warp 0's staging rows from this one up (the rows below hold the merge scratch)
    uint32_t node_lo, exp_cap;
    uint64_t neg_zero2;        // (-0.0f, -0.0f): addend of the packed product in lane_exact_distance_x2 (opaque to the compiler)
    uint64_t *exp_keys;        // [nq][exp_cap] (distance,id) keys in expansion order
    uint32_t *exp_cnt;         // [nq]
    // byte offsets inside the CTA's shared memory
    uint32_t off_pool, off_cand, off_sorted, off_pos, off_fresh, off_ctrl, off_hash, off_warp;
    uint32_t warp_bytes, woff_cid, woff_mine, woff_exc, woff_stage;  // per-warp area: [mbarrier][candidate ids][ids to filter][exception list][row staging]
};

// ---- exact visited set --------------------------------------------------------------------------
// returns 1 = first visit, 0 = already visited, 2 = table cannot take the id (16-bit flavour: displacement field exhausted)
__device__ __forceinline__ uint32_t visited_test_and_set32(uint32_t *table, uint32_t log2, uint32_t id) {
    const uint32_t mask = (1u << log2) - 1u;
    uint32_t slot = (id * 0x9E3779B1u) >> (32 - log2);
    for (;;) {
        uint32_t old = atomicCAS(table + slot, kEmpty, id);
        if (old == kEmpty) return 1u;   // first visit
        if (old == id) return 0u;       // already visited
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ uint32_t hash16_home(const SearchParams &p, uint32_t id, uint32_t *rem) {
    const uint32_t x = (id * p.h16_mult) & ((1u << p.h16_bits) - 1u);
    *rem = x & ((1u << p.h16_rbits) - 1u);
    return x >> p.h16_rbits;
}
// Two 16-bit entries share a 32-bit word and every update is ONE 32-bit atomicCAS on that word (CUDA's 16-bit atomicCAS
// is a software loop - a load, then a 32-bit CAS - i.e. two dependent memory round trips per probe; ncu showed it as the
// top stall of the kernel).  The first touch of a word bets that both halves are still empty: if the CAS succeeds the id
// is inserted in one round trip, and if it fails the returned word tells whether the id is already there (also one round
// trip) or which half is taken (then a second CAS, or the next slot, with the line already in L2).
__device__ __forceinline__ uint32_t visited_test_and_set16(uint32_t *table32, const SearchParams &p, uint32_t id) {
    const uint32_t mask = (1u << p.hash_log2) - 1u;
    uint32_t rem;
    uint32_t slot = hash16_home(p, id, &rem);
    const uint32_t hi = rem << p.h16_dbits;
    uint32_t cur = 0, cur_word = 0xFFFFFFFFu;  // last observed content of word `cur_word`
    for (uint32_t d = 0; d <= p.h16_maxd; ++d) {
        const uint32_t want = hi | d, widx = slot >> 1, sh = (slot & 1u) << 4;
        uint32_t *w = table32 + widx;
        if (widx != cur_word) {
            cur = atomicCAS(w, 0xFFFFFFFFu, ~(0xFFFFu << sh) | (want << sh));
            if (cur == 0xFFFFFFFFu) return 1u;  // both halves were empty: inserted
            cur_word = widx;
        }
        for (;;) {
            const uint32_t e = (cur >> sh) & 0xFFFFu;
            if (e == want) return 0u;           // same displacement -> same home slot, same remainder -> same id
            if (e != 0xFFFFu) break;            // taken by another id: next slot
            const uint32_t old = atomicCAS(w, cur, (cur & ~(0xFFFFu << sh)) | (want << sh));
            if (old == cur) return 1u;
            cur = old;                          // the word changed under us: look again
        }
        slot = (slot + 1) & mask;
    }
    return 2u;
}
__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr)); }

// ---- bucketed visited set without atomics (kHashBucket16 / kHashBucket32) ------------------------------
__device__ __forceinline__ uint4 ld_cg_v4(const void *ptr) {  // L2-only vector load: never served from a stale L1 line
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}
// Home bucket of an id, the warp that owns it and the part of the entry that does not depend on the displacement.  The
// bucket count nb is ANY multiple of W, so a slab is as large as the beam width needs and not the next power of two (which
// decides whether the slabs of all resident queries fit L2).  16-bit flavour: x = id * M mod 2^B is a bijection of the id
// range; with a = x / 2^B in [0, 1) the home bucket is floor(a * nb) and the owner floor(a * W) = home / (nb / W).  The x
// values of one bucket are a contiguous run of at most ceil(2^B / nb) <= 2^rbits numbers, so x mod 2^rbits tells them
// apart: that is what the entry stores - no division anywhere.
template <int kHash>
__device__ __forceinline__ uint32_t bucket_home(const SearchParams &p, uint32_t id, uint32_t W, uint32_t *owner, uint32_t *tag) {
    if (kHash == kHashBucket16) {
        const uint32_t x = (id * p.h16_mult) & ((1u << p.h16_bits) - 1u);
        const uint32_t a = x << (32 - p.h16_bits);
        *tag = (x & ((1u << p.h16_rbits) - 1u)) << p.h16_dbits;
        *owner = __umulhi(a, W);
        return __umulhi(a, p.nb);
    }
    const uint32_t a = id * 0x9E3779B1u;
    *tag = id;
    *owner = __umulhi(a, W);
    return __umulhi(a, p.nb);
}
__device__ __forceinline__ void bucket_scan16(const uint4 s, uint32_t want, bool *found, uint32_t *cnt) {
    const uint32_t w[4] = {s.x, s.y, s.z, s.w};
    bool f = false;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
        f |= (lo == want) | (hi == want);
        c += (lo != 0xFFFFu ? 1u : 0u) + (hi != 0xFFFFu ? 1u : 0u);
    }
    *found = f;
    *cnt = c;
}
__device__ __forceinline__ void bucket_scan32(const uint4 a, const uint4 b, uint32_t want, bool *found, uint32_t *cnt) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    bool f = false;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f |= w[i] == want;
        c += w[i] != kEmpty ? 1u : 0u;
    }
    *found = f;
    *cnt = c;
}
// Warp-collective test-and-set of one id per active lane inside this warp's bucket range [blo, bhi).  Returns 1 = first
// visit, 0 = already visited (or inactive lane), 2 = displacement field exhausted.  `pre` is the content of the lane's home
// bucket when have_pre (loaded a hop ahead; nothing has been inserted since).  Every round: load the current bucket, look
// for the entry, count the used slots; lanes that want to append to the same bucket take consecutive slots in lane order
// (match.any); whoever finds the bucket full moves to the next one (displacement + 1) and sees this round's stores there
// because __syncwarp orders them before the next round's loads.
// An id whose whole probe window (home bucket + displacement range) is full goes to a small per-warp EXCEPTION LIST in
// shared memory (s_exc[0] = count, then ids).  With 13-bit remainders only 3 displacement bits are left, and although a run
// of 7 full buckets is a 1e-8 event per insert at load 0.5, a 10 000-query batch at L_pq = 200 makes 1e8 inserts: the two
// to six queries per batch that hit it used to be re-run from scratch by the big-table pass after the launch (5-18 % of
// the batch time).  A later lookup of such an id walks the same full window, ends here too and finds it in the list.
constexpr uint32_t kExcCap = 3;  // 16 bytes per warp with the count; a fourth such id in one query takes the big-table pass
template <int kHash>
__device__ __forceinline__ uint32_t bucket_test_and_set(unsigned char *slab, const SearchParams &p, uint32_t blo, uint32_t bhi,
                                                        bool active, uint32_t id, bool have_pre, const uint4 pre, uint32_t lane,
                                                        uint32_t *s_exc) {
    uint32_t tag, owner;
    uint32_t bucket = bucket_home<kHash>(p, id, 1u, &owner, &tag), d = 0, res = 0;
    const uint32_t dummy = 0x80000000u | lane;  // ids and bucket indices are < 2^31
    bool pending = active;
    bool need_load = !have_pre;
    uint4 s0 = pre, s1 = make_uint4(0, 0, 0, 0);
    while (__any_sync(0xffffffffu, pending)) {
        bool found = false;
        uint32_t cnt = 0;
        const uint32_t want = kHash == kHashBucket16 ? (tag | d) : id;
        if (kHash == kHashBucket16) {
            if (pending && need_load) s0 = ld_cg_v4(slab + (size_t(bucket) << 4));
            bucket_scan16(s0, want, &found, &cnt);
        } else {
            if (pending) {
                s0 = ld_cg_v4(slab + (size_t(bucket) << 5));
                s1 = ld_cg_v4(slab + (size_t(bucket) << 5) + 16);
            }
            bucket_scan32(s0, s1, want, &found, &cnt);
        }
        if (found) pending = false;
        bool prop = pending && cnt < 8u;
        uint32_t grp = __match_any_sync(0xffffffffu, prop ? bucket : dummy);
        if (__any_sync(0xffffffffu, prop && (grp & (grp - 1u)) != 0u)) {
            // several lanes append to one bucket (rare).  It may be the same id twice in the row: such lanes have walked the
            // same buckets in lockstep, so they meet here; the first lane decides and the others report "visited"
            const uint32_t peers = __match_any_sync(0xffffffffu, prop ? id : dummy);
            if (prop && uint32_t(__ffs(peers)) - 1u != lane) prop = pending = false;
            grp = __match_any_sync(0xffffffffu, prop ? bucket : dummy);
        }
        if (prop) {
            const uint32_t slot = cnt + __popc(grp & lanemask_lt());
            if (slot < 8u) {
                if (kHash == kHashBucket16) __stcg(reinterpret_cast<unsigned short *>(slab) + (size_t(bucket) << 3) + slot, (unsigned short)want);
                else __stcg(reinterpret_cast<uint32_t *>(slab) + (size_t(bucket) << 3) + slot, want);
                res = 1u;
                pending = false;
            }
        }
        if (pending) {  // bucket full
            if (++d > p.h16_maxd) {
                res = 3u;  // probe window exhausted: exception list, below
                pending = false;
            } else {
                bucket = (bucket + 1u == bhi) ? blo : bucket + 1u;
                need_load = true;
            }
        }
        __syncwarp();
    }
    uint32_t exc = __ballot_sync(0xffffffffu, res == 3u);
    while (exc) {  // cold: one lane at a time
        const uint32_t src = __ffs(exc) - 1u;
        const uint32_t eid = __shfl_sync(0xffffffffu, id, src);
        const uint32_t n_exc = s_exc[0];
        const bool hit = __any_sync(0xffffffffu, lane < n_exc && s_exc[1 + lane] == eid);
        if (lane == src) {
            if (hit) res = 0u;
            else if (n_exc < kExcCap) {
                s_exc[1 + n_exc] = eid;
                s_exc[0] = n_exc + 1u;
                res = 1u;
                atomicAdd(&p.counters[kCntException], 1u);  // diagnostics
            } else res = 2u;  // the list is full as well: big-table pass
        }
        __syncwarp();
        exc &= exc - 1u;
    }
    return res;
}

// first index in sorted keys[0..n) whose key (flag bit cleared) is >= key
__device__ __forceinline__ uint32_t lower_bound_key(const uint64_t *keys, uint32_t n, uint64_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((keys[mid] & ~1ull) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// kGather: 1 = cp.async (LDGSTS 16 B per lane), 2 = TMA bulk copy (one UBLKCP per row) on an mbarrier
template <bool kIP, int kGather, int kHash, bool kBuild>
__global__ void __launch_bounds__(kMaxWarps * 32, RG_K1_MIN_CTAS) rg_search_kernel(const SearchParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, T = blockDim.x, W = T >> 5;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t grp = lane >> 2, t = lane & 3;
    float *s_query = reinterpret_cast<float *>(smem_raw);
    uint64_t *P = reinterpret_cast<uint64_t *>(smem_raw + p.off_pool);
    uint64_t *s_cand = reinterpret_cast<uint64_t *>(smem_raw + p.off_cand);
    uint64_t *s_sorted = reinterpret_cast<uint64_t *>(smem_raw + p.off_sorted);
    uint32_t *s_pos = reinterpret_cast<uint32_t *>(smem_raw + p.off_pos);
    uint32_t *s_fresh = reinterpret_cast<uint32_t *>(smem_raw + p.off_fresh);
    volatile uint32_t *s_ctrl = reinterpret_cast<volatile uint32_t *>(smem_raw + p.off_ctrl);
    uint32_t *s_ctrl_nv = reinterpret_cast<uint32_t *>(smem_raw + p.off_ctrl);
    unsigned char *wa = smem_raw + p.off_warp + size_t(warp) * p.warp_bytes;
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(wa);
    uint32_t *s_cid = reinterpret_cast<uint32_t *>(wa + p.woff_cid);
    float *s_stage = reinterpret_cast<float *>(wa + p.woff_stage);
    constexpr bool kBucket = kHash == kHashBucket16 || kHash == kHashBucket32;
    uint32_t *s_mine = reinterpret_cast<uint32_t *>(wa + p.woff_mine);  // bucket flavours: ids of the expanded row this warp owns
    uint32_t *s_exc = reinterpret_cast<uint32_t *>(wa + p.woff_exc);    // bucket flavours: exception list (see bucket_test_and_set)
    uint32_t *hash32 = kHash == kHashShared ? reinterpret_cast<uint32_t *>(smem_raw + p.off_hash)
                       : kBucket            ? p.ghash + size_t(blockIdx.x) * p.slab_words
                                            : p.ghash + (size_t(blockIdx.x) << (kHash == kHashGlobal16 ? p.hash_log2 - 1 : p.hash_log2));
    unsigned short *hash16 = reinterpret_cast<unsigned short *>(hash32);
    unsigned char *slab = reinterpret_cast<unsigned char *>(hash32);
    // bucket flavours: this warp's bucket range
    const uint32_t blo = kBucket ? warp * p.nb_per_warp : 0u;
    const uint32_t bhi = kBucket ? blo + p.nb_per_warp : 0u;

    const uint32_t dim = p.dim, n16 = dim >> 4;
    const bool tail8 = (dim & 15u) != 0;
    const uint32_t cpr = dim >> 2;  // 16-byte chunks per row
    const uint32_t L = p.L, BR = p.stage_rows, RS = p.row_stride;
    uint32_t mb_phase = 0;
    const bool rows_evict_first = (p.l2_hint & 1u) != 0;
    const bool pf_next = (p.adj_prefetch & 1u) != 0, pf_cand = (p.adj_prefetch & 2u) != 0;
    const uint64_t pol_first = l2_policy_evict_first();
    const uint32_t adj_row_bytes = p.adj_stride * 4u;

    if (kGather == 2) {
        if (lane == 0) {
            mbar_init(s_mbar, 1);
            fence_mbar_init();
        }
    }
    __syncthreads();

    auto visit = [&](uint32_t id) -> uint32_t {
        if (kHash == kHashGlobal16) return visited_test_and_set16(hash32, p, id);
        return visited_test_and_set32(hash32, p.hash_log2, id);
    };

    // speculation: adjacency words of the node expected to be expanded next (see step 2 above) and, bucket flavours, the
    // number of its neighbours this warp owns (listed in s_mine) with the content of lane l's home bucket
    uint32_t spec_id = kEmpty, spec_deg = 0, sreg[3] = {kEmpty, kEmpty, kEmpty};
    uint32_t spec_n = 0, self = 0;
    uint4 spec_sec = make_uint4(0, 0, 0, 0);

    // bucket flavours: the ids of an adjacency row (first 96 words in wr, word i = wr[i / 32] of lane i % 32) whose home
    // bucket belongs to this warp, compacted into s_mine in row order
    auto compact_mine = [&](const uint32_t *row, uint32_t deg, const uint32_t (&wr)[3]) -> uint32_t {
        uint32_t n = 0;
        for (uint32_t it = 0; it * 32 < deg; ++it) {
            const uint32_t idx = lane + 32 * it;
            uint32_t word;
            if (it == 0) word = wr[0];
            else if (it == 1) word = wr[1];
            else if (it == 2) word = wr[2];
            else word = (idx < deg) ? __ldg(row + 1 + idx) : kEmpty;
            bool mine = idx < deg && !(kBuild && word == self);
            uint32_t tag, owner = W;
            if (mine) bucket_home<kHash>(p, word, W, &owner, &tag);
            mine = owner == warp;
            const uint32_t m = __ballot_sync(0xffffffffu, mine);
            if (mine) s_mine[n + __popc(m & lanemask_lt())] = word;
            n += __popc(m);
        }
        __syncwarp();
        return n;
    };
    // bucket flavours, while the hop's first rows are in flight: list the speculated node's neighbours this warp owns and
    // request their home buckets (into registers, or into L2)
    auto spec_block = [&]() {
        if (!kBucket || spec_id == kEmpty) return;
        if (spec_deg > 96u) {  // longer rows are not speculated on
            spec_id = kEmpty;
            return;
        }
        spec_n = compact_mine(nullptr, spec_deg, sreg);
        for (uint32_t i = lane; i < spec_n; i += 32) {
            uint32_t tag, owner;
            const uint32_t b = bucket_home<kHash>(p, s_mine[i], 1u, &owner, &tag);
            if (kHash == kHashBucket16) {
                if (RG_K1_SPEC_SECTOR_REGS && i < 32) spec_sec = ld_cg_v4(slab + (size_t(b) << 4));
                else prefetch_l2(slab + (size_t(b) << 4));
            } else {
                prefetch_l2(slab + (size_t(b) << 5));
            }
        }
    };

    // bucket flavours: test-and-set of the n_mine ids in s_mine (this warp's share of the expanded row); the unvisited ones
    // are compacted into s_cid.  from_spec: the list was made a hop ahead and lane l holds the bucket of entry l
    auto filter_mine = [&](uint32_t n_mine, bool from_spec) -> uint32_t {
        uint32_t n_w = 0;
        for (uint32_t r0 = 0; r0 < n_mine; r0 += 32) {
            const bool act = r0 + lane < n_mine;
            const uint32_t id = act ? s_mine[r0 + lane] : 0u;
            const bool pre = RG_K1_SPEC_SECTOR_REGS && kHash == kHashBucket16 && from_spec && r0 == 0;
            uint32_t v = kBucket ? bucket_test_and_set<kBucket ? kHash : kHashBucket16>(slab, p, blo, bhi, act, id, pre, spec_sec, lane, s_exc) : 0u;
            if (v == 2u) {
                s_ctrl[kCtlHashFull] = 1;
                v = 0;
            }
            const bool fresh = v == 1u;
            const uint32_t m = __ballot_sync(0xffffffffu, fresh);
            if (fresh) s_cid[n_w + __popc(m & lanemask_lt())] = id;
            n_w += __popc(m);
        }
        __syncwarp();
        return n_w;
    };

    // issues the gather of rows list[0..rows) into this warp's staging rows row0 .. row0 + rows - 1
    auto issue_rows = [&](const uint32_t *list, uint32_t rows, uint32_t row0) {
        if (kGather == 2) {
            if (lane == 0) mbar_arrive_expect_tx(s_mbar, rows * dim * 4u);
            __syncwarp();
            if (lane < rows) {  // stage_rows <= 32: one row per lane
                float *dst = s_stage + (row0 + lane) * RS;
                const float *src = p.base + size_t(list[lane]) * dim;
                // the gathered rows are touched once: evict_first keeps them from displacing adjacency/hash lines
                if (rows_evict_first) bulk_g2s_hint(dst, src, dim * 4u, s_mbar, pol_first);
                else bulk_g2s(dst, src, dim * 4u, s_mbar);
            }
        } else {
            const uint32_t total = rows * cpr;
            for (uint32_t idx = lane; idx < total; idx += 32) {
                const uint32_t r = __umulhi(idx, p.chunk_magic);
                const uint32_t c = idx - r * cpr;
                cp_async16(s_stage + size_t(row0 + r) * RS + 4 * c, p.base + size_t(list[r]) * dim + 4 * c);
            }
            cp_async_commit();
        }
    };

    // Gathers and scores the rows s_cid[0..n) (this warp's own list, or its batch of the CTA-wide one); keys below `tail`
    // are appended to the CTA-wide candidate list.  The first `pre` rows may already be on their way (staging rows pre_row0..).
    // Candidates that beat `next_key` (the best unexpanded pool entry besides the node being expanded) are expanded before
    // it: their adjacency rows are prefetched into L2 while the rest of the hop is still being scored and merged, and the
    // hop is flagged so that the next one does not start from the speculated node.
    auto gather_and_score = [&](const uint32_t *s_cid, uint32_t n, uint64_t tail, uint32_t ctl, uint64_t next_key, bool do_spec,
                                uint32_t pre, uint32_t pre_row0) {
        for (uint32_t c0 = 0; c0 < n;) {
            uint32_t rows, row0 = 0;
            if (c0 == 0 && pre) {
                rows = pre;
                row0 = pre_row0;
            } else {
                rows = min(BR, n - c0);
                issue_rows(s_cid + c0, rows, 0);
            }
            if (do_spec && c0 == 0) spec_block();
            if (kGather == 2) {
                mbar_wait(s_mbar, mb_phase);
                mb_phase ^= 1u;
            } else {
                cp_async_wait<0>();
                __syncwarp();
            }
            const float *stage = s_stage + size_t(row0) * RS;
            for (uint32_t r0 = 0; r0 < rows; r0 += 8) {
                const uint32_t r = r0 + grp;
                const bool valid = r < rows;
                const uint32_t rr = valid ? r : rows - 1;
                const float4 *rp = reinterpret_cast<const float4 *>(stage + size_t(rr) * RS) + t;
                const float4 *qp = reinterpret_cast<const float4 *>(s_query) + t;
                const float d = lane_exact_distance_x2<kIP>(rp, qp, n16, tail8, t, p.neg_zero2);
                const uint64_t key = make_key(d, s_cid[c0 + rr]);
                // NeighborPriorityQueue::insert rejects keys behind the last entry of a full pool (neighbor.h:151);
                // the tail only tightens during a hop, so dropping them here is exact
                const bool keep = valid && t == 0 && key < tail;
                if (keep && key < next_key) {
                    if (pf_cand) bulk_prefetch_l2(p.adj + size_t(s_cid[c0 + rr]) * p.adj_stride, adj_row_bytes);
                    s_ctrl[kCtlBeat0 + ((ctl - kCtlHop0) >> 2)] = 1;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    const uint32_t leader = __ffs(m) - 1;
                    uint32_t pos0 = 0;
                    if (lane == leader) pos0 = atomicAdd(&s_ctrl_nv[ctl], uint32_t(__popc(m)));
                    pos0 = __shfl_sync(0xffffffffu, pos0, leader);
                    if (keep) s_cand[pos0 + __popc(m & lanemask_lt())] = key;
                }
            }
            __syncwarp();  // all reads of the staging buffer done before the next batch lands in it
            c0 += rows;
        }
    };

    for (;;) {
        // ---- next query (schedule(dynamic,1)) ---------------------------------------------------
        if (tid == 0) {
            s_ctrl[kCtlWork] = atomicAdd(&p.counters[p.fallback ? kCntWork2 : kCntWork], 1u);
            s_ctrl[kCtlNvis] = 0;
            s_ctrl[kCtlHashFull] = 0;
            s_ctrl[kCtlHop0 + 0] = 0;  // ncand
            s_ctrl[kCtlHop0 + 1] = 0;  // ndup
            s_ctrl[kCtlHop0 + 2] = L;  // minlo
            s_ctrl[kCtlHop0 + 3] = L;  // curpos
            s_ctrl[kCtlHop0 + 4] = 0;
            s_ctrl[kCtlHop0 + 5] = 0;
            s_ctrl[kCtlHop0 + 6] = L;
            s_ctrl[kCtlHop0 + 7] = L;
            s_ctrl[kCtlBeat0] = 0;
            s_ctrl[kCtlBeat0 + 1] = 0;
            s_ctrl[kCtlFresh] = 0;
            s_ctrl[kCtlBatch] = 0;
        }
        if (p.steal_batches && tid < 16) s_ctrl[kCtlPub0 + tid] = 0;  // publication words and batch counters of the W <= 8 warps
        if (kBucket && lane == 0) s_exc[0] = 0;
        __syncthreads();
        const uint32_t w = s_ctrl[kCtlWork];
        const uint32_t nwork = p.fallback ? min(p.counters[kCntOverflow], p.nq) : p.nq;
        if (w >= nwork) break;
        const uint32_t qi = p.fallback ? p.overflow_list[w] : w;

        {   // query -> shared memory; clear the visited set
            const float4 *src = reinterpret_cast<const float4 *>(kBuild ? p.base + (size_t(p.node_lo) + qi) * dim
                                                                        : p.queries + size_t(qi) * dim);
            float4 *dst = reinterpret_cast<float4 *>(s_query);
            for (uint32_t i = tid; i < cpr; i += T) dst[i] = src[i];
            uint4 *h4 = reinterpret_cast<uint4 *>(hash32);
            const uint4 e4 = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            const uint32_t n_vec = kBucket ? p.slab_words >> 2 : 1u << (p.hash_log2 - (kHash == kHashGlobal16 ? 3 : 2));
            for (uint32_t i = tid; i < n_vec; i += T) h4[i] = e4;
        }
        __syncthreads();

        uint32_t size = 0, cur = 0, hops = 0, nvis = 0, hp = 0;  // hp: parity of the hop's control words
        uint64_t tail = ~0ull;  // (distance,id) of the last entry once the pool is full, else +inf
        bool have_cur = false, overflow = false;
        spec_id = kEmpty;

        // entry point: scored and inserted, NOT marked visited (src/index_bipartite.cpp:2337-2353); the build-time
        // search does mark it (:1309)
        self = p.node_lo + qi;
        if (kBuild && kBucket) {  // every warp takes part; the owner of the entry point's home bucket inserts it
            uint32_t tag, owner;
            bucket_home<kHash>(p, p.ep, W, &owner, &tag);
            const bool act = lane == 0 && owner == warp;
            bucket_test_and_set<kHash>(slab, p, blo, bhi, act, p.ep, false, spec_sec, lane, s_exc);
        }
        if (warp == 0) {
            if (lane == 0) {
                s_cid[0] = p.ep;
                if (kBuild && !kBucket) visit(p.ep);
            }
            __syncwarp();
            gather_and_score(s_cid, 1, tail, kCtlHop0, ~0ull, false, 0, 0);
        }

        for (;;) {
            __syncthreads();  // the hop's candidates are in s_cand
            const uint32_t ctl = kCtlHop0 + 4 * hp, octl = kCtlHop0 + 4 * (hp ^ 1u);  // this hop's / the other hop's words
            const uint32_t C = s_ctrl[ctl + 0];
            nvis = s_ctrl[kCtlNvis];
            if ((kHash == kHashGlobal16 || kBucket) && s_ctrl[kCtlHashFull]) {  // a displacement field ran out: big-table pass
                overflow = true;
                break;
            }
            // Early expansion (bucket flavours): the previous hop read ahead the node it expected to be expanded next - its
            // neighbours this warp owns are listed in s_mine, their buckets are in registers.  If no candidate of that hop
            // beat it, it IS the next node and keeps its pool position (every new key lands behind it), so the visited
            // filter and the first gather are issued now and the rows travel while the CTA merges the candidates.
            uint32_t n_w = 0, pre_rows = 0;
            const uint32_t pre_row0 = warp == 0 ? p.early_row0 : 0u;
            const bool early = kBucket && p.early_issue && have_cur && spec_id != kEmpty && s_ctrl[kCtlBeat0 + hp] == 0 &&
                               nvis + p.adj_stride <= p.hash_limit;
            if (early) {
                n_w = filter_mine(spec_n, true);
                if (n_w) {
                    if (lane == 0) atomicAdd(&s_ctrl_nv[kCtlNvis], n_w);
                    pre_rows = min(BR - pre_row0, n_w);
                    issue_rows(s_cid, pre_rows, pre_row0);
                }
            }
            uint32_t start;
            if (C == 0) {
                // nothing to insert: flag the expanded entry (closest_unexpanded, neighbor.h:185-192)
                if (tid == 0) {
                    P[cur] |= 1ull;
                    s_ctrl[octl + 0] = 0;
                    s_ctrl[octl + 1] = 0;
                    s_ctrl[octl + 2] = L;
                    s_ctrl[octl + 3] = L;
                    s_ctrl[kCtlBeat0 + (hp ^ 1u)] = 0;
                    s_ctrl[kCtlFresh] = 0;  // every warp has left the previous hop's batch loop
                    s_ctrl[kCtlBatch] = 0;
                }
                if (p.steal_batches && tid < 16) s_ctrl[kCtlPub0 + tid] = 0;
                __syncthreads();
                start = cur + 1;
            } else {
                // (a) position of every candidate among the pool entries; a candidate equal to a pool entry is the
                //     re-scored entry point: "Make sure the same id isn't inserted into the set" (neighbor.h:161)
                uint32_t lb0 = 0;  // pool position of candidate `tid` (kept for (b))
                for (uint32_t j = tid; j < C; j += T) {
                    const uint64_t key = s_cand[j];
                    const uint32_t lo = lower_bound_key(P, size, key);
                    if (j == tid) lb0 = lo;
                    if (lo < size && (P[lo] & ~1ull) == key) {
                        s_cand[j] = ~0ull;
                        atomicAdd(&s_ctrl_nv[ctl + 1], 1u);
                    } else {
                        atomicMin(&s_ctrl_nv[ctl + 2], lo);
                    }
                }
                if (tid == 0) {  // the other hop's control words are free again
                    s_ctrl[octl + 0] = 0;
                    s_ctrl[octl + 1] = 0;
                    s_ctrl[octl + 2] = L;
                    s_ctrl[octl + 3] = L;
                    s_ctrl[kCtlBeat0 + (hp ^ 1u)] = 0;
                    s_ctrl[kCtlFresh] = 0;  // every warp has left the previous hop's batch loop
                    s_ctrl[kCtlBatch] = 0;
                }
                if (p.steal_batches && tid < 16) s_ctrl[kCtlPub0 + tid] = 0;
                __syncthreads();
                const uint32_t Cn = C - s_ctrl[ctl + 1];   // candidates that are really new
                const uint32_t minlo = min(s_ctrl[ctl + 2], size);  // pool entries in front of it do not move
                // (b) candidates sorted by counting; final position = #pool entries + #candidates in front
                for (uint32_t j = tid; j < C; j += T) {
                    const uint64_t key = s_cand[j];
                    if (key == ~0ull) continue;
                    uint32_t r = 0;
                    for (uint32_t i = 0; i < C; ++i) r += (s_cand[i] < key) ? 1u : 0u;
                    s_sorted[r] = key;
                    s_pos[r] = (j == tid ? lb0 : lower_bound_key(P, size, key)) + r;
                }
                if (have_cur && cur < minlo && tid == 0) {  // the expanded entry stays where it is
                    P[cur] |= 1ull;
                    s_ctrl[ctl + 3] = cur;
                }
                __syncthreads();
                // (c) pool entries [minlo, size) shift right by the number of candidates in front of them, in place, in
                //     chunks of 4T entries from the top down; a chunk's new positions are >= its old ones, i.e. inside the
                //     chunk itself (read before the barrier) or above it (already moved).  Candidate r stands in front of
                //     pool entry i iff its pool position q_r = s_pos[r] - r is <= i (keys are distinct), so the shift of
                //     entry i is a binary search over 32-bit words.  A thread moves two PAIRS of neighbouring entries
                //     (i, i+1): the second shift is the first plus the candidates with q_r == i + 1, one search per pair.
                for (uint32_t hi = size; hi > minlo;) {
                    const uint32_t lo_c = (hi - minlo > 4 * T) ? hi - 4 * T : minlo;
                    uint64_t e[4];
                    uint32_t pos[4];
#pragma unroll
                    for (uint32_t u = 0; u < 2; ++u) {
                        const uint32_t i = lo_c + 2 * (tid + u * T);
                        pos[2 * u] = pos[2 * u + 1] = L;
                        e[2 * u] = e[2 * u + 1] = 0;
                        if (i < hi) {
                            e[2 * u] = P[i];
                            uint32_t a = 0, b = Cn;
                            while (a < b) {
                                const uint32_t mid = (a + b) >> 1;
                                if (s_pos[mid] - mid <= i) a = mid + 1;
                                else b = mid;
                            }
                            pos[2 * u] = i + a;
                            if (i + 1 < hi) {
                                e[2 * u + 1] = P[i + 1];
                                while (a < Cn && s_pos[a] - a <= i + 1) ++a;
                                pos[2 * u + 1] = i + 1 + a;
                            }
                            if (have_cur && i == cur) {
                                e[2 * u] |= 1ull;
                                s_ctrl[ctl + 3] = pos[2 * u];
                            }
                            if (have_cur && i + 1 == cur && i + 1 < hi) {
                                e[2 * u + 1] |= 1ull;
                                s_ctrl[ctl + 3] = pos[2 * u + 1];
                            }
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u)
                        if (pos[u] < L) P[pos[u]] = e[u];
                    hi = lo_c;
                }
                for (uint32_t r = tid; r < Cn; r += T) {
                    const uint32_t pos = s_pos[r];
                    if (pos < L) P[pos] = s_sorted[r];
                }
                __syncthreads();
                size = min(L, size + Cn);
                const uint32_t curpos = s_ctrl[ctl + 3];
                start = have_cur ? min(minlo, curpos + 1) : 0u;
            }
            hp ^= 1u;
            {   // first unexpanded entry at or after `start` (everything in front of it is expanded)
                uint32_t c = start, next = size;
                while (c < size) {
                    const uint32_t i = c + lane;
                    const bool unexp = (i < size) && ((P[i] & 1ull) == 0);
                    const uint32_t m = __ballot_sync(0xffffffffu, unexp);
                    if (m) {
                        next = c + (__ffs(m) - 1);
                        break;
                    }
                    c += 32;
                }
                cur = next;
            }
            if (size == L) tail = P[L - 1] & ~1ull;
            if (cur >= size) break;

            // ---- expand P[cur] -------------------------------------------------------------------
            have_cur = true;
            const uint32_t cur_id = key_id(P[cur]);
            if (kBuild && tid == 0 && hops < p.exp_cap) p.exp_keys[size_t(qi) * p.exp_cap + hops] = P[cur] & ~1ull;  // :1318
            ++hops;
            // speculation for the NEXT hop: the best unexpanded entry behind `cur` is expanded next unless a candidate of
            // this hop beats it; its adjacency row is read now (same round trip as the row of `cur`, or the first gather)
            uint64_t next_key = tail;
            auto spec_scan = [&]() {
                spec_id = kEmpty;
                if (!(pf_next || pf_cand)) return;
                uint32_t c = cur + 1, nx = size;
                while (c < size) {
                    const uint32_t i = c + lane;
                    const bool unexp = (i < size) && ((P[i] & 1ull) == 0);
                    const uint32_t m = __ballot_sync(0xffffffffu, unexp);
                    if (m) {
                        nx = c + (__ffs(m) - 1);
                        break;
                    }
                    c += 32;
                }
                if (nx < size) {
                    next_key = P[nx] & ~1ull;
                    if (pf_next) {
                        spec_id = key_id(next_key);
                        const uint32_t *nrow = p.adj + size_t(spec_id) * p.adj_stride;
#pragma unroll
                        for (uint32_t it = 0; it < 3; ++it) {
                            const uint32_t j = kBucket ? lane + 32 * it : (lane + 32 * it) * W + warp;
                            sreg[it] = (j + 1 < p.adj_stride) ? __ldg(nrow + 1 + j) : kEmpty;
                        }
                        spec_deg = __ldg(nrow);
                    }
                }
            };
            if (early) {
                spec_scan();  // the pool is merged now; the filter and the first rows of this hop are already out
            } else {
                if (nvis + p.adj_stride > p.hash_limit) {  // visited set may fill up: hand over to the big-table pass
                    overflow = true;
                    break;
                }
                // adjacency row.  CAS flavours: neighbour j is handled by warp j % W, lane (j / W) % 32; bucket flavours:
                // every warp reads the whole row (word i by lane i % 32) and keeps the neighbours whose home bucket it owns.
                // The first three rounds are requested together with the degree word: one DRAM round trip - or none at
                // all when this node was the one read ahead during the previous hop
                const uint32_t *row = p.adj + size_t(cur_id) * p.adj_stride;
                const bool spec_hit = spec_id == cur_id;
                uint32_t wreg[3], deg;
                if (spec_hit) {
#pragma unroll
                    for (uint32_t it = 0; it < 3; ++it) wreg[it] = sreg[it];
                    deg = spec_deg;
                } else {
#pragma unroll
                    for (uint32_t it = 0; it < 3; ++it) {
                        const uint32_t j = kBucket ? lane + 32 * it : (lane + 32 * it) * W + warp;
                        wreg[it] = (j + 1 < p.adj_stride) ? __ldg(row + 1 + j) : kEmpty;
                    }
                    deg = __ldg(row);
                }
                spec_scan();
                if (kBucket) {
                    // this warp's neighbours are in s_mine: listed a hop ahead (speculation held) or now
                    n_w = filter_mine(spec_hit ? spec_n : compact_mine(row, deg, wreg), spec_hit);
                } else {
                    for (uint32_t it = 0; it * 32 * W < deg; ++it) {
                        const uint32_t j = (lane + 32 * it) * W + warp;
                        uint32_t word;
                        if (it == 0) word = wreg[0];
                        else if (it == 1) word = wreg[1];
                        else if (it == 2) word = wreg[2];
                        else word = (j < deg) ? __ldg(row + 1 + j) : kEmpty;
                        uint32_t v = 0;
                        if (j < deg && !(kBuild && word == self)) v = visit(word);
                        if (kHash == kHashGlobal16 && v == 2u) {
                            s_ctrl[kCtlHashFull] = 1;
                            v = 0;
                        }
                        const bool fresh = v == 1u;
                        const uint32_t m = __ballot_sync(0xffffffffu, fresh);
                        if (fresh) s_cid[n_w + __popc(m & lanemask_lt())] = word;
                        n_w += __popc(m);
                    }
                    __syncwarp();
                }
                if (!kBucket && pf_next && spec_id != kEmpty && kHash != kHashShared) {
                    // pull the visited-hash slots the speculated node's neighbours map to into L2: one hop from now their
                    // atomicCAS probes are L2 hits instead of HBM round trips on the query's dependent chain
#pragma unroll
                    for (uint32_t it = 0; it < 3; ++it) {
                        const uint32_t j = (lane + 32 * it) * W + warp;
                        if (j < spec_deg) {
                            if (kHash == kHashGlobal16) {
                                uint32_t rem;
                                prefetch_l2(hash16 + hash16_home(p, sreg[it], &rem));
                            } else {
                                prefetch_l2(hash32 + ((sreg[it] * 0x9E3779B1u) >> (32 - p.hash_log2)));
                            }
                        }
                    }
                }
            }
            // a re-scored entry point lands in the lists below too; the merge drops it as a duplicate (neighbor.h:161) or the
            // tail test drops it (neighbor.h:151), exactly like the reference
            if (p.shared_batches) {
                // one CTA-wide list, batches of BR rows pulled by whichever warp is free: no warp waits at the hop barrier
                // because its share of the unvisited neighbours was larger, and only the hop's last batch is partial
                uint32_t pos0 = 0;
                if (lane == 0 && n_w) {
                    pos0 = atomicAdd(&s_ctrl_nv[kCtlFresh], n_w);
                    atomicAdd(&s_ctrl_nv[kCtlNvis], n_w);
                }
                pos0 = __shfl_sync(0xffffffffu, pos0, 0);
                for (uint32_t i = lane; i < n_w; i += 32) s_fresh[pos0 + i] = s_cid[i];
                __syncthreads();
                const uint32_t F = s_ctrl[kCtlFresh];
                bool first = true;
                for (;;) {
                    uint32_t b = 0;
                    if (lane == 0) b = atomicAdd(&s_ctrl_nv[kCtlBatch], 1u);
                    b = __shfl_sync(0xffffffffu, b, 0) * BR;
                    if (b >= F) break;
                    gather_and_score(s_fresh + b, min(BR, F - b), tail, kCtlHop0 + 4 * hp, next_key, first, 0, 0);
                    first = false;
                }
                if (first) spec_block();
            } else if (p.steal_batches) {
                // per-warp lists with work stealing: every list is handed out in batches of BR rows through a counter; a warp
                // that has emptied its own list takes batches of the other warps' lists (once those are published), so
                // nobody waits at the hop barrier because its share of the unvisited neighbours was a batch shorter
                if (lane == 0) {
                    if (n_w) atomicAdd(&s_ctrl_nv[kCtlNvis], n_w);
                    __threadfence_block();  // the list (written before the __syncwarp above) before its length
                    s_ctrl[kCtlPub0 + warp] = n_w + 1u;
                }
                bool first = true;
                for (uint32_t v = 0; v < W; ++v) {
                    const uint32_t vw = warp + v < W ? warp + v : warp + v - W;
                    uint32_t n_v = n_w;
                    if (v) {
                        uint32_t pub = 0;
                        if (lane == 0) pub = s_ctrl[kCtlPub0 + vw];
                        pub = __shfl_sync(0xffffffffu, pub, 0);
                        if (pub <= BR + 1u) continue;  // not published yet, or a single batch that its owner is working on
                        n_v = pub - 1u;
                    }
                    const uint32_t *list = reinterpret_cast<const uint32_t *>(smem_raw + p.off_warp + size_t(vw) * p.warp_bytes + p.woff_cid);
                    for (;;) {
                        uint32_t b = 0;
                        if (lane == 0) b = atomicAdd(&s_ctrl_nv[kCtlNext0 + vw], 1u);
                        b = __shfl_sync(0xffffffffu, b, 0) * BR;
                        if (b >= n_v) break;
                        gather_and_score(list + b, min(BR, n_v - b), tail, kCtlHop0 + 4 * hp, next_key, first, 0, 0);
                        first = false;
                    }
                }
                if (first) spec_block();
            } else if (n_w) {
                if (!early && lane == 0) atomicAdd(&s_ctrl_nv[kCtlNvis], n_w);
                gather_and_score(s_cid, n_w, tail, kCtlHop0 + 4 * hp, next_key, true, pre_rows, pre_row0);
            } else {
                spec_block();
            }
        }

        if (overflow) {
            if (!p.fallback) {
                if (tid == 0) {
                    const uint32_t pos = atomicAdd(&p.counters[kCntOverflow], 1u);
                    p.overflow_list[pos] = qi;
                }
                __syncthreads();
                continue;
            }
            if (tid == 0) atomicAdd(&p.counters[kCntFatal], 1u);
            size = 0;  // falls through to the "not enough results" fill
        }
        if (kBuild) {
            if (tid == 0) p.exp_cnt[qi] = overflow ? 0u : min(hops, p.exp_cap);
            __syncthreads();
            continue;
        }
        // results (src/index_bipartite.cpp:2408-2419)
        if (size < p.k) {
            if (tid == 0 && !overflow) atomicAdd(&p.counters[kCntNotEnough], 1u);
            for (uint32_t i = tid; i < p.k; i += T) {
                p.ids[size_t(qi) * p.k + i] = kEmpty;
                p.dists[size_t(qi) * p.k + i] = 0.f;
            }
        } else {
            for (uint32_t i = tid; i < p.k; i += T) {
                const uint64_t e = P[i];
                p.ids[size_t(qi) * p.k + i] = key_id(e);
                p.dists[size_t(qi) * p.k + i] = key_dist(e);
            }
        }
        if (tid == 0) {
            if (p.cmps) p.cmps[qi] = nvis;
            if (p.hops) p.hops[qi] = hops;
        }
        __syncthreads();
    }
}

// ---- host side: geometry + launch ---------------------------------------------------------------
typedef void (*SearchKernel)(const SearchParams);

struct Geometry {
    SearchParams p;
    int warps, gather, hash_kind;
    size_t smem_bytes;
    uint64_t slab_bytes;  // global visited-hash bytes per CTA (0: shared memory)
    SearchKernel fn;
    int ctas_per_sm;
};

static uint32_t round_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

static uint32_t auto_hash_log2(uint32_t L, bool global_space) {
    // visited nodes per query ~ 1000 + 30 L on the 100K probe set (max cmps 1371 @L=10 ... 13702 @L=500)
    // and ~1.7x that at 10M (mean 3180 @L=60).  A slab in global memory is sized for a load factor <= 0.4 at that
    // estimate (short probe chains, overflow pass practically never needed); a shared-memory table for <= 0.8.
    // Outliers take the exact big-table fallback pass, so this only affects speed.
    const double want = (1000.0 + 30.0 * L) / (global_space ? 0.4 : 0.8);
    uint32_t lg = 10;
    while ((1u << lg) < want && lg < 22) ++lg;
    return lg;
}

template <bool kIP, int kGather, int kHash>
static SearchKernel pick_build(bool build) {
    if (build) return rg_search_kernel<kIP, 2, kHash == kHashShared ? kHashGlobal32 : kHash, true>;
    return rg_search_kernel<kIP, kGather, kHash, false>;
}
template <bool kIP, int kGather>
static SearchKernel pick_hash(int hash_kind, bool build) {
    if (hash_kind == kHashBucket16) return pick_build<kIP, 2, kHashBucket16>(build);  // bucket flavours: TMA gather only
    if (hash_kind == kHashBucket32) return pick_build<kIP, 2, kHashBucket32>(build);
    if (hash_kind == kHashGlobal16) return pick_build<kIP, kGather, kHashGlobal16>(build);
    if (hash_kind == kHashGlobal32) return pick_build<kIP, kGather, kHashGlobal32>(build);
    return pick_build<kIP, kGather, kHashShared>(build);
}
static SearchKernel pick_kernel(bool ip, int gather, int hash_kind, bool build) {
    if (gather == 2 || build) return ip ? pick_hash<true, 2>(hash_kind, build) : pick_hash<false, 2>(hash_kind, build);
    return ip ? pick_hash<true, 1>(hash_kind, build) : pick_hash<false, 1>(hash_kind, build);
}

static bool persisting_window_fits(const rg_index *ix, uint64_t bytes);

static rg_status make_geometry(const rg_index *ix, uint32_t k, uint32_t L, bool fallback, bool build, Geometry *g,
                               int warps_override = 0, bool allow16 = true, int space_override = -1) {
    SearchParams &p = g->p;
    const int space = space_override >= 0 ? space_override : ix->cfg_hash_space;
    memset(&p, 0, sizeof(p));
    p.neg_zero2 = kNegZero2;
    p.dim = ix->dim;
    p.adj_stride = ix->adj_stride;
    p.ep = ix->ep;
    p.k = k;
    p.L = L;
    const uint32_t cpr = ix->dim / 4;
    p.chunk_magic = uint32_t((0x100000000ull + cpr - 1) / cpr);
    // smallest rs >= dim with rs % 32 == 16: the two rows a quarter-warp reads with LDS.128 fall in
    // different halves of the 32 banks
    p.row_stride = (ix->dim % 32 <= 16) ? ix->dim - ix->dim % 32 + 16 : ix->dim - ix->dim % 32 + 48;
    p.stage_rows = ix->cfg_stage_rows ? uint32_t(ix->cfg_stage_rows) : 8u;
    g->gather = ix->cfg_gather ? ix->cfg_gather : 2;
    if (build) g->gather = 2;
    g->warps = ix->cfg_warps ? ix->cfg_warps : 2;  // measured best on B200 (profiles/r01_k1_v2_sweep.txt)
    if (warps_override) g->warps = warps_override;
    const uint32_t W = uint32_t(g->warps);

    uint32_t hl = ix->cfg_hash_log2 ? uint32_t(ix->cfg_hash_log2) : auto_hash_log2(L, build || space != 1);
    p.fallback = fallback ? 1u : 0u;
    p.l2_hint = uint32_t(ix->cfg_l2_hint);
    p.adj_prefetch = uint32_t(ix->cfg_adj_prefetch) & 3u;
    const int batch_mode = ix->cfg_batch_mode;
    // visited set: a slab per CTA in global memory unless shared memory was asked for (hash_space 1).
    //   hash_space 0 (auto) / 4: buckets without atomics - 16-bit quotient entries (8 per 16-byte bucket) when the id range
    //                            leaves >= 3 displacement bits at no more than twice the slots, else 32-bit ids (8 per 32 B)
    //   hash_space 5: 32-bit buckets;  2 / 3: round 1/2's atomicCAS tables (32-bit keys / 16-bit quotient entries)
    const bool global_hash = fallback || build || hl > 15 || space != 1;
    if (fallback) hl = std::min<uint32_t>(22u, std::max<uint32_t>(16u, hl + 3));
    g->hash_kind = global_hash ? kHashGlobal32 : kHashShared;
    uint32_t id_bits = 1;
    while ((1ull << id_bits) < ix->n && id_bits < 31) ++id_bits;
    const bool want_bucket = global_hash && !fallback && (space == 0 || space == 4 || space == 5);
    p.hash_limit = 0;
    uint64_t bucket_slab_bytes = 0;
    if (want_bucket) {
        // slots wanted: load <= 0.4 at the usual number of visited nodes (auto_hash_log2 rounds this up to a power of two
        // for the atomicCAS tables; the buckets take it as it is)
        // (target loads of 0.5 / 0.6 - smaller slabs - measured no faster, and 0.6 runs into the load limit:
        // profiles/r02_k1_table_sizing.txt)
        const double want = ix->cfg_hash_log2 ? double(1u << ix->cfg_hash_log2) : (1000.0 + 30.0 * L) / 0.4;
        const uint32_t nbw = std::max<uint32_t>(uint32_t((want + 7.0) / 8.0), 8u);           // 8 entries per bucket
        const uint32_t min_nb16 = id_bits > 13 ? 1u << (id_bits - 13) : 1u;                   // entry keeps <= 13 bits of x
        const uint32_t nb16 = std::max(nbw, min_nb16);
        if (space != 5 && allow16 && nb16 <= 2 * nbw && nb16 <= (1u << 19)) {
            g->hash_kind = kHashBucket16;
            p.nb = round_up(nb16, W);
            p.h16_bits = id_bits;
            const uint64_t run = ((1ull << id_bits) + p.nb - 1) / p.nb;  // x values per bucket
            p.h16_rbits = 0;
            while ((1ull << p.h16_rbits) < run) ++p.h16_rbits;
            p.h16_dbits = std::min<uint32_t>(16u - p.h16_rbits, 8u);
            p.h16_maxd = (1u << p.h16_dbits) - 2u;  // all-ones is kept for the empty entry
            p.h16_mult = (uint32_t(double(1ull << p.h16_bits) * 0.6180339887498949) | 1u) & uint32_t((1ull << p.h16_bits) - 1);
            bucket_slab_bytes = uint64_t(p.nb) * 16;
        } else {
            g->hash_kind = kHashBucket32;
            p.nb = round_up(nbw, W);
            p.h16_maxd = 30;
            bucket_slab_bytes = uint64_t(p.nb) * 32;
        }
        p.nb_per_warp = p.nb / W;
        p.slab_words = uint32_t(bucket_slab_bytes / 4);
        // a warp's bucket range must stay longer than the longest probe walk
        p.h16_maxd = std::min<uint32_t>(p.h16_maxd, std::max<uint32_t>(1u, p.nb_per_warp - 1u));
        // The load limit only keeps probe walks short; what a bucket table cannot take shows up as an exhausted displacement
        // field (the insert reports it and the query goes to the big-table pass).  That pass re-runs a query from scratch
        // after the primary launch - four queries of 10 000 cost 18 % at L_pq = 200 - so heavy queries may fill the table to
        // 90 % before they are handed over (at 70 % the 4 heaviest of 10 000 were: 1.4x the mean number of visited nodes).
        p.hash_limit = uint32_t(uint64_t(p.nb) * 8 * 90 / 100);
    } else if (global_hash && !fallback && space != 2 && allow16) {
        const uint32_t h16 = std::max<uint32_t>(hl, id_bits > 10 ? id_bits - 10 : 0);  // >= 6 displacement bits
        if (h16 <= hl + 1 && h16 <= 22) {
            g->hash_kind = kHashGlobal16;
            hl = h16;
            p.h16_bits = std::max(id_bits, hl);
            p.h16_rbits = p.h16_bits - hl;
            p.h16_dbits = std::min<uint32_t>(16u - p.h16_rbits, 12u);
            p.h16_maxd = (1u << p.h16_dbits) - 2u;
            // odd multiplier ~ 2^bits / golden ratio: a bijection of [0, 2^bits) whose top bits mix well
            p.h16_mult = (uint32_t(double(1ull << p.h16_bits) * 0.6180339887498949) | 1u) & uint32_t((1ull << p.h16_bits) - 1);
        }
    }
    p.hash_log2 = hl;
    if (!p.hash_limit) p.hash_limit = uint32_t((uint64_t(1) << hl) * 85 / 100);
    g->slab_bytes = g->hash_kind == kHashShared ? 0
                    : bucket_slab_bytes     ? bucket_slab_bytes
                                            : (uint64_t(g->hash_kind == kHashGlobal16 ? 2 : 4) << hl);
    const bool bucket = g->hash_kind == kHashBucket16 || g->hash_kind == kHashBucket32;
    // one CTA-wide gather list with dynamic batches: measured no better than per-warp lists (profiles/r02_k1_sweep_buckets.txt), opt-in
    p.shared_batches = batch_mode == 2 && W > 1 ? 1u : 0u;
    p.steal_batches = batch_mode == 3 && W > 1 ? 1u : 0u;
    p.early_issue = bucket && (ix->cfg_adj_prefetch & 1) && (ix->cfg_adj_prefetch & 4) && !p.shared_batches && !p.steal_batches ? 1u : 0u;

    // Shared-memory layout, packed to 16 bytes (TMA bulk destinations and LDS.128 need no more): the CTA count per SM is
    // decided by it (12 CTAs of two warps fit up to L_pq ~ 170 at D = 200, 11 at 500).  The merge scratch (sorted candidates
    // and their positions) aliases warp 0's row staging buffer: a merge only runs between hops, when no gather is in flight.
    uint32_t off = round_up(ix->dim * 4, 16);
    p.off_pool = off;
    off += round_up((L + 1) * 8, 16);
    p.off_cand = off;
    off += round_up(ix->adj_stride * 8, 16);
    p.off_ctrl = off;
    off += p.steal_batches ? 128 : 64;  // words 16..31 are the stealing mode's publication words and batch counters
    p.off_fresh = off;
    if (p.shared_batches) off += round_up(ix->adj_stride * 4, 16);
    p.off_hash = off;
    if (g->hash_kind == kHashShared) off += (4u << hl);
    p.off_warp = off;
    // CAS flavours: a warp filters every W-th neighbour; bucket flavours: any share of the row may hash into its range
    const uint32_t cid_cap = bucket ? round_up(ix->adj_stride, 4) : round_up((ix->adj_stride - 1 + W - 1) / W, 4);
    p.woff_cid = 16;
    p.woff_mine = p.woff_cid + cid_cap * 4;
    p.woff_exc = p.woff_mine + (bucket ? cid_cap * 4 : 0);
    p.woff_stage = round_up(p.woff_exc + (bucket ? (kExcCap + 1) * 4 : 0), 16);
    // the padding behind the last staged row is not needed
    const uint32_t stage_bytes = std::max<uint32_t>(round_up(((p.stage_rows - 1) * p.row_stride + ix->dim) * 4, 16),
                                                    round_up(ix->adj_stride * 8, 16) + round_up(ix->adj_stride * 4, 16));
    p.warp_bytes = p.woff_stage + stage_bytes;
    p.off_sorted = p.off_warp + p.woff_stage;                       // aliases warp 0's staging rows
    {   // early issue: warp 0's first rows land above the merge scratch
        const uint32_t scratch = round_up(ix->adj_stride * 8, 16) + round_up(ix->adj_stride * 4, 16);
        p.early_row0 = (scratch + p.row_stride * 4 - 1) / (p.row_stride * 4);
        if (p.early_row0 >= p.stage_rows) p.early_issue = 0;
    }
    p.off_pos = p.off_sorted + round_up(ix->adj_stride * 8, 16);
    off += W * p.warp_bytes;
    g->smem_bytes = off;
    if (size_t(off) > size_t(ix->max_smem_optin))
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq=%u needs %u bytes of shared memory per query (max %d)", L, off,
                        ix->max_smem_optin);
    g->fn = pick_kernel(ix->metric != RG_METRIC_L2, g->gather, g->hash_kind, build);
    cudaError_t e = cudaFuncSetAttribute(g->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g->smem_bytes));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->ctas_per_sm, g->fn, g->warps * 32, g->smem_bytes);
    if (e != cudaSuccess) return rg::fail(RG_ERR_CUDA, "K1 launch configuration failed: %s", cudaGetErrorString(e));
    if (g->ctas_per_sm < 1) return rg::fail(RG_ERR_INTERNAL, "K1 does not fit on an SM (L_pq=%u)", L);
    if (ix->cfg_ctas) g->ctas_per_sm = std::min(g->ctas_per_sm, ix->cfg_ctas);
    return RG_OK;
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

// Launches g.fn with an access-policy window that marks [ptr, ptr + bytes) as persisting in L2.  Only used when the whole
// window fits the device's persisting set-aside: pinning a random fraction of slabs that are larger than L2 anyway (the
// L_pjpq = 500 build searches: 256 KB per CTA) only takes cache away from the adjacency rows - measured 38 s -> 55 s
// for the 10M connectivity-enhancement searches.  The set-aside is a device-wide limit; it follows the window size.
static bool persisting_window_fits(const rg_index *ix, uint64_t bytes) {
    int max_persist = 0, max_window = 0;
    if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ix->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ix->device) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return bytes > 0 && bytes <= uint64_t(max_persist) && bytes <= uint64_t(max_window);
}

// Device-wide persisting-L2 set-aside, shared by every index on the device.  A launch WITHOUT a window must give it back:
// lines that an earlier launch marked persisting stay pinned, and the set-aside stays carved out of L2, until they are
// reset - an L_pq = 200 batch right after an L_pq = 100 batch otherwise runs with 57 MB of L2 holding dead lines.
static std::mutex g_persist_mu;
static uint64_t g_persist_bytes[64] = {0};

static rg_status set_persisting_limit(rg_index *ix, uint64_t bytes) {
    std::lock_guard<std::mutex> lock(g_persist_mu);
    uint64_t &cur = g_persist_bytes[ix->device & 63];
    if (cur != bytes) {
        if (bytes == 0) RG_CUDA_OK(cudaCtxResetPersistingL2Cache());
        RG_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(bytes)));
        cur = bytes;
    }
    ix->persist_bytes = bytes;
    return RG_OK;
}

// hit_ratio < 1: the window is larger than the set-aside and a random fraction of its lines is pinned
static rg_status launch_with_persisting_window(rg_index *ix, const Geometry &g, int grid, void *ptr, uint64_t bytes,
                                               uint64_t limit_bytes, float hit_ratio, cudaStream_t st) {
    rg_status s = set_persisting_limit(ix, limit_bytes);
    if (s != RG_OK) return s;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(g.warps * 32));
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    memset(&attr, 0, sizeof(attr));
    attr.id = cudaLaunchAttributeAccessPolicyWindow;
    attr.val.accessPolicyWindow.base_ptr = ptr;
    attr.val.accessPolicyWindow.num_bytes = size_t(bytes);
    attr.val.accessPolicyWindow.hitRatio = hit_ratio;
    attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    RG_CUDA_OK(cudaLaunchKernelEx(&cfg, g.fn, g.p));
    return RG_OK;
}

// d_exp_keys != nullptr selects the build-time variant: the queries are base rows node_lo .. node_lo + nq - 1
static rg_status search_device_impl(rg_index *ix, const float *d_queries, uint64_t nq, uint32_t k, uint32_t L,
                                    uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops,
                                    uint32_t *d_status, uint32_t node_lo, uint64_t *d_exp_keys, uint32_t *d_exp_cnt,
                                    uint32_t exp_cap, cudaStream_t st) {
    const bool build = d_exp_keys != nullptr;
    if (!ix || (!build && (!d_queries || !d_ids || !d_dists))) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search: null argument");
    if (k == 0 || L == 0 || k > L) return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq must greater or equal than k (k=%u, L_pq=%u)", k, L);
    if (L > 8192) return rg::fail(RG_ERR_INVALID_ARGUMENT, "L_pq=%u too large (max 8192)", L);
    if (nq >= (1ull << 32)) return rg::fail(RG_ERR_INVALID_ARGUMENT, "too many queries in one batch");
    if (nq == 0) return RG_OK;

    Geometry g1, g2;
    rg_status s = make_geometry(ix, k, L, true, build, &g2);
    if (s != RG_OK) return s;
    // Primary pass, automatic mode (same-box sweeps in profiles/r02_k1_sweep_*.txt): two warps per query; atomicCAS on 32-bit
    // keys while the slabs of all resident queries fit the persisting part of L2 (L_pq <= ~75 at 10M rows: 0.89 of the HBM
    // peak at L_pq = 55 against 0.85 for the buckets, whose warps each read the whole adjacency row); beyond that the
    // bucketed visited set without atomics (0.67 vs 0.60 at L_pq = 200, 0.56 vs 0.51 at 500).
    const bool auto_hash = ix->cfg_hash_space == 0;
    s = make_geometry(ix, k, L, false, build, &g1);
    if (s != RG_OK) return s;
    if (auto_hash && !build) {
        Geometry g;
        const rg_status s2 = make_geometry(ix, k, L, false, build, &g, 0, true, 2);
        if (s2 == RG_OK && g.slab_bytes &&
            persisting_window_fits(ix, std::min<uint64_t>(nq, uint64_t(ix->sm_count) * g.ctas_per_sm) * g.slab_bytes))
            g1 = g;
    }

    // scratch: overflow list (one slot per query) and global hash slabs (one per CTA).  The fallback pass re-runs the few
    // queries whose visited set outgrew the primary table with big 32-bit tables; its grid is kept small (64 CTAs) so that
    // the scratch it needs (4 MB per CTA at L_pq = 500, 16 MB at the maximum table) stays modest on a device that already
    // holds a 100M-row index.
    s = ensure((void **)&ix->d_overflow_list, &ix->overflow_cap, nq, sizeof(uint32_t));
    if (s != RG_OK) return s;
    const int grid1 = int(std::min<uint64_t>(nq, uint64_t(ix->sm_count) * g1.ctas_per_sm));
    const int grid2 = int(std::min<uint64_t>(nq, 64));
    const uint64_t need_bytes = std::max(uint64_t(grid2) * g2.slab_bytes, uint64_t(grid1) * g1.slab_bytes);
    s = ensure((void **)&ix->d_ghash, &ix->ghash_words, (need_bytes + 3) / 4, sizeof(uint32_t));
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
        g->p.node_lo = node_lo;
        g->p.exp_keys = d_exp_keys;
        g->p.exp_cnt = d_exp_cnt;
        g->p.exp_cap = exp_cap;
    }
    RG_CUDA_OK(cudaMemsetAsync(ix->d_counters, 0, 8 * sizeof(uint32_t), st));
    const uint64_t slab_bytes = uint64_t(grid1) * g1.slab_bytes;
    if (g1.slab_bytes && (ix->cfg_l2_hint & 2) && persisting_window_fits(ix, slab_bytes)) {
        // pin the visited-hash slabs of the resident CTAs in the persisting part of L2 (atomics take no cache hint)
        s = launch_with_persisting_window(ix, g1, grid1, ix->d_ghash, slab_bytes, slab_bytes, 1.0f, st);
        if (s != RG_OK) return s;
    } else {
        s = set_persisting_limit(ix, 0);
        if (s != RG_OK) return s;
        g1.fn<<<grid1, g1.warps * 32, g1.smem_bytes, st>>>(g1.p);
    }
    RG_CUDA_OK(cudaGetLastError());
    ix->launches++;
    g2.fn<<<grid2, g2.warps * 32, g2.smem_bytes, st>>>(g2.p);  // exits immediately when nothing overflowed
    RG_CUDA_OK(cudaGetLastError());
    ix->launches++;
    if (d_status) {
        RG_CUDA_OK(cudaMemcpyAsync(d_status, ix->d_counters + kCntNotEnough, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        RG_CUDA_OK(cudaMemcpyAsync(d_status + 1, ix->d_counters + kCntFatal, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return RG_OK;
}

static rg_status search_device(rg_index *ix, const float *d_queries, uint64_t nq, uint32_t k, uint32_t L,
                               uint32_t *d_ids, float *d_dists, uint32_t *d_cmps, uint32_t *d_hops,
                               uint32_t *d_status, cudaStream_t st) {
    return search_device_impl(ix, d_queries, nq, k, L, d_ids, d_dists, d_cmps, d_hops, d_status, 0, nullptr, nullptr, 0, st);
}

// Build-time beam searches (connectivity enhancement): expanded nodes of base rows [node_lo, node_lo + count)
rg_status search_expanded_device(rg_index *ix, uint32_t node_lo, uint64_t count, uint32_t L, uint64_t *d_exp_keys,
                                 uint32_t *d_exp_cnt, uint32_t exp_cap, cudaStream_t st) {
    return search_device_impl(ix, nullptr, count, 1, L, nullptr, nullptr, nullptr, nullptr, nullptr, node_lo, d_exp_keys,
                              d_exp_cnt, exp_cap, st);
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

rg_status rg_search_expanded_device(rg_index *ix, uint32_t node_lo, uint64_t count, uint32_t L, uint64_t *d_exp_keys,
                                    uint32_t *d_exp_cnt, uint32_t exp_cap, void *cuda_stream) {
    if (!ix || !d_exp_keys || !d_exp_cnt || exp_cap == 0)
        return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_expanded_device: null argument");
    if (uint64_t(node_lo) + count > ix->n) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_expanded_device: rows out of range");
    rg::DeviceGuard guard(ix->device);
    return rg::search_expanded_device(ix, node_lo, count, L, d_exp_keys, d_exp_cnt, exp_cap, static_cast<cudaStream_t>(cuda_stream));
}

// diagnostics: ids of the last batch that went to a warp's exception list (bucketed visited set, probe window exhausted)
uint32_t rg_search_last_exception_count(rg_index *ix) {
    if (!ix) return 0;
    rg::DeviceGuard guard(ix->device);
    uint32_t v = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpy(&v, ix->d_counters + rg::kCntException, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return v;
}

// diagnostics: queries of the last batch whose visited set outgrew the primary table and were redone by the big-table pass
uint32_t rg_search_last_overflow_count(rg_index *ix) {
    if (!ix) return 0;
    rg::DeviceGuard guard(ix->device);
    uint32_t v = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpy(&v, ix->d_counters + rg::kCntOverflow, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return v;
}

// Device-visible alias of a caller buffer when ALL of it is page-locked host memory (cudaHostAlloc / cudaHostRegister; under
// unified addressing every such allocation is mapped), else nullptr.  First and last byte must belong to host
// registrations whose device aliases are contiguous: a buffer that only shares its first page with a registered neighbour
// (two small heap vectors on one page) must take the staged path.
static void *mapped_alias(const void *p, uint64_t bytes) {
    if (!bytes) return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, static_cast<const char *>(p) + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
    if (static_cast<char *>(a1.devicePointer) - static_cast<char *>(a0.devicePointer) != ptrdiff_t(bytes - 1)) return nullptr;
    return a0.devicePointer;
}

static rg_status report_status(const uint32_t status[2], uint32_t k, uint32_t L) {
    if (status[1]) return rg::fail(RG_ERR_INTERNAL, "visited set overflow in %u queries (L_pq=%u)", status[1], L);
    if (status[0]) {
        // report the first short query like the reference's message (src/index_bipartite.cpp:2408-2412)
        return rg::fail(RG_ERR_NOT_ENOUGH_RESULTS, "not enough results: fewer than %u pool entries in %u queries, expected: %u", k, status[0], k);
    }
    return RG_OK;
}

rg_status rg_search_batch(rg_index *ix, const float *queries, uint64_t nq, uint32_t k, uint32_t L, uint32_t *ids,
                          float *dists, uint32_t *cmps, uint32_t *hops) {
    if (!ix || !queries || !ids || !dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_search_batch: null argument");
    if (nq == 0) return RG_OK;
    rg::DeviceGuard guard(ix->device);
    rg_status s;
    cudaStream_t st = ix->stream;
    uint32_t status[2] = {0, 0};
    if (ix->stat_cap < nq + 2) {
        cudaFree(ix->d_cmps);
        cudaFree(ix->d_hops);
        ix->d_cmps = ix->d_hops = nullptr;
        ix->stat_cap = 0;
        RG_CUDA_OK(cudaMalloc(&ix->d_cmps, (nq + 2) * sizeof(uint32_t)));
        RG_CUDA_OK(cudaMalloc(&ix->d_hops, (nq + 2) * sizeof(uint32_t)));
        ix->stat_cap = nq + 2;
    }
    uint32_t *d_status = ix->d_cmps + nq;  // two spare words behind the cmps array

    // Zero-copy path: when every caller buffer is page-locked host memory the kernel reads each query (D floats, once,
    // into shared memory) and writes each result row straight through the mapped pointers, so the host<->device
    // transfers ride inside the search instead of in front of and behind it (8 MB in / 0.8 MB out per 10 000 queries
    // at D=200, k=10: 2 GB/s against >50 GB/s of PCIe, hidden behind the HBM-bound gathers of the other resident queries).
    if (ix->cfg_zero_copy) {
        const float *zq = static_cast<const float *>(mapped_alias(queries, nq * ix->dim * sizeof(float)));
        uint32_t *zi = static_cast<uint32_t *>(mapped_alias(ids, nq * k * sizeof(uint32_t)));
        float *zd = static_cast<float *>(mapped_alias(dists, nq * k * sizeof(float)));
        uint32_t *zc = cmps ? static_cast<uint32_t *>(mapped_alias(cmps, nq * sizeof(uint32_t))) : nullptr;
        uint32_t *zh = hops ? static_cast<uint32_t *>(mapped_alias(hops, nq * sizeof(uint32_t))) : nullptr;
        if (zq && zi && zd && (!cmps || zc) && (!hops || zh)) {
            s = rg::search_device(ix, zq, nq, k, L, zi, zd, zc, zh, d_status, st);
            if (s != RG_OK) return s;
            RG_CUDA_OK(cudaMemcpyAsync(status, d_status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            RG_CUDA_OK(cudaStreamSynchronize(st));
            return report_status(status, k, L);
        }
    }

    // Staged path (pageable caller buffers): H2D copy, search on index-owned device buffers, D2H copies.
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
    RG_CUDA_OK(cudaMemcpyAsync(ix->d_queries, queries, nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    s = rg::search_device(ix, ix->d_queries, nq, k, L, ix->d_ids, ix->d_dists, ix->d_cmps, ix->d_hops, d_status, st);
    if (s != RG_OK) return s;
    RG_CUDA_OK(cudaMemcpyAsync(ids, ix->d_ids, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(dists, ix->d_dists, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (cmps) RG_CUDA_OK(cudaMemcpyAsync(cmps, ix->d_cmps, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (hops) RG_CUDA_OK(cudaMemcpyAsync(hops, ix->d_hops, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaMemcpyAsync(status, d_status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    return report_status(status, k, L);
}

}  // extern "C"
