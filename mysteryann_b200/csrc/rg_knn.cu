// K2/K3/K4 - build-time exact kNN (replaces exact_knn + the part merge of the DiskANN tool the reference uses,
// /root/reference thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:126-248, 396-448).
//
// The reference materialises a 1024 x npoints FP32 score block in host RAM with MKL sgemm and scans it with a heap
// per query.  Here the scores never leave the SM:
//   K2  tcgen05 GEMM (kind::f16, FP16 operands, FP32 accumulate in TMEM) over tiles of 128 queries x 128 base rows;
//       operands arrive by TMA (cp.async.bulk.tensor, 128B swizzle) in a ring of 64-wide K slabs; the epilogue reads
//       the accumulators with tcgen05.ld and keeps only the scores that beat the query's current threshold
//       (appended to a small per-query candidate list) - a threshold-filtered top-k' instead of a heap scan.
//       The base is visited in blocks of doubling size; after each block K2s (one warp per query) finds the k'-th
//       best score of the list by bisection, compacts the list to the entries at or below it and tightens the threshold.
//   K3  FP32 re-rank of the k' survivors in the reference's lane-ordered arithmetic + a per-query CERTIFICATE:
//       every point that is not a survivor has approx score >= tau (the worst survivor); the FP16 rounding error of a
//       score is bounded by eps = 2^-10 |q| max|b| (x2 for L2), so if exact_K < tau - eps the top-K is provably
//       complete.  Queries that fail the certificate (or whose list overflowed) are redone by an exact FP32 scan.
//   K4  k-way merge of per-shard lists (multi-GPU: base sharded, lists exchanged with NCCL by the caller).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "rg_knn.cuh"

static thread_local uint64_t g_knn_stats[3] = {0, 0, 0};  // per calling thread (one host thread per GPU in the tools)

namespace rg {
namespace knn {

void set_last_stats(const uint64_t stats[3]) { memcpy(g_knn_stats, stats, sizeof(g_knn_stats)); }

constexpr int kTileM = 128;      // queries per CTA (its 128 TMEM lanes); the CTA pair's MMA covers 2 x 128
constexpr int kHalfN = 128;      // base rows each CTA of the pair stages per tile
constexpr int kPairN = 256;      // base rows per MMA tile (TMEM columns per accumulator)
constexpr int kTileN = kHalfN;   // TMA box rows
constexpr int kSlabK = 64;       // FP16 elements per K slab = 128 bytes = one 128B-swizzle row
constexpr int kSlabBytes = kTileN * kSlabK * 2;  // 16 KB
constexpr int kUmmaK = 16;
constexpr int kMaxSlabs = 8;     // dim <= 512
constexpr int kEpiWarps = 16;    // epilogue warps: warp w reads TMEM lanes 32*(w%4).., column quarter w/4 (64 columns)
constexpr int kThreads = (kEpiWarps + 2) * 32;  // + warp 16 TMA producer, warp 17 MMA issuer + TMEM owner
constexpr uint32_t kCap = 1024;  // candidate list capacity per query
constexpr uint32_t kStageSlots = 4;  // per-lane shared-memory slots for candidates found by the GEMM epilogue
constexpr int kChunkTiles = 64;  // base tiles (256 rows) per work unit (the query tile stays resident for a whole unit)

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMA)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // relaxed: nothing written by this thread is read through the barrier (the TMEM loads have completed in
    // tcgen05.wait::ld; tcgen05.fence::before_thread_sync orders them before the arrive); a release at cluster
    // scope would cost a memory fence per tile and warp
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA tile load of a CTA pair: data lands in the executing CTA's shared memory, the transaction bytes are
// signalled on `bar_cluster_addr`, a barrier of the pair's leader CTA (cute SM100_TMA_2SM_LOAD_2D).
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// TMEM of a CTA pair: the same warp of both CTAs allocates / frees (cute::TMEM::Allocator2Sm)
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// One MMA of the CTA pair, issued by one thread of the leader: D[256 x N] (+)= A[256 x 16] * B[N x 16]^T, FP16 operands
// from both CTAs' shared memory (same offsets), FP32 accumulators in both CTAs' TMEM.  Descriptors are given as
// (low word, shared high word): low = smem address >> 4 (advance 2 per 32-byte K step).
__device__ __forceinline__ void umma2_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (count 1) on the barrier at this offset in BOTH CTAs when all prior MMAs of this thread retire
__device__ __forceinline__ void umma2_commit(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
        "h"(uint16_t(3))
        : "memory");
}
// 32 lanes x 32 consecutive FP32 columns of the accumulator -> 32 registers per thread (asynchronous: tmem_ld_wait())
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// The wait that makes the registers of outstanding tcgen05.ld's valid, tied to those registers: the "+r" operands make
// every later use of v depend on this statement, so the compiler cannot hoist pure arithmetic on the loaded values above
// the wait (a volatile asm with a memory clobber alone only orders memory operations and other volatile asms).
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), K-major, swizzled: low word = start address >> 4,
// high word = SBO (8 rows x row bytes) >> 4 | version 1 << 14 | layout << 29 (2 = 128B swizzle: 64 halves per row,
// 4 = 64B: 32 halves, 6 = 32B: 16 halves).  LBO is ignored for swizzled K-major operands.
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=F32 (1<<4), A=B=F16 (0), K-major both,
// N>>3 at bit 17, M>>4 at bit 24 (M = 256: the pair's two 128-row halves).
constexpr uint32_t kIdesc2 = (1u << 4) | (uint32_t(kPairN >> 3) << 17) | (uint32_t((2 * kTileM) >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------------
// FP32 -> FP16 slab layout [nslab][rows_pad][64], scaled by a power of two; also squared norms / max |x|
// ---------------------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float *__restrict__ x, uint64_t count, uint32_t *out_bits) {
    float m = 0.f;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += uint64_t(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// one warp per row: columns [0, n_full*64) go to the full-slab array [n_full][rows_pad][64], the remaining
// columns (zero padded to tail_k) to the tail array [rows_pad][tail_k]
__global__ void to_half_slabs_kernel(const float *__restrict__ x, uint64_t rows, uint32_t dim, uint64_t rows_pad,
                                     uint32_t n_full, uint32_t tail_k, float scale, __half *__restrict__ out_full,
                                     __half *__restrict__ out_tail, float *__restrict__ norms, uint32_t *max_norm_bits) {
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t kfull = n_full * kSlabK;
    for (uint64_t r = warp; r < rows_pad; r += nwarps) {
        float ss = 0.f;
        for (uint32_t c = lane; c < kfull + tail_k; c += 32) {
            float v = (r < rows && c < dim) ? x[r * dim + c] : 0.f;
            ss += v * v;
            const __half h = __float2half_rn(v * scale);
            if (c < kfull) out_full[(uint64_t(c / kSlabK) * rows_pad + r) * kSlabK + (c % kSlabK)] = h;
            else out_tail[r * tail_k + (c - kfull)] = h;
        }
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0 && r < rows) {
            if (norms) norms[r] = ss;
            if (max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(ss));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2: GEMM + threshold filter.  CTA PAIR (cluster of 2, tcgen05 cta_group::2): one MMA covers 256 queries (128 TMEM
// lanes in each CTA) x 256 base rows (each CTA stages 128 of them), so every SM reads 4 KB of A + 4 KB of B from its
// shared memory per 128 tensor cycles (64 B/clk) instead of 128 B/clk for a 128x128 single-CTA tile, and a B slab
// fetched from L2 feeds 256 queries.  The query tile stays resident in shared memory for a whole work unit
// (chunk_tiles base tiles) and is double buffered across units when the budget allows.
// ---------------------------------------------------------------------------------------------------------------
struct GemmParams {
    uint32_t n_full;       // full 64-wide K slabs (128B swizzle)
    uint32_t tail_k;       // 0, 16 (32B swizzle) or 32 (64B swizzle): width of the last, narrower K slab
    uint32_t a_bytes;      // shared-memory bytes of one resident query tile (128 rows, all its slabs), 1024-aligned
    uint32_t a_bufs;       // 1 or 2 query-tile buffers
    uint32_t nq;           // valid queries in this batch
    uint32_t m_tiles;      // ceil(nq / 256)
    uint32_t chunk_tiles;  // base tiles (of 256 rows) per work unit
    uint64_t q_rows_pad;   // padded row count of the query arrays
    uint64_t b_rows_pad;   // padded row count of the base arrays
    uint64_t row_lo;       // first base row of this block (multiple of 256)
    uint32_t n_tiles;      // base tiles in this block
    uint64_t n_valid;      // number of real base rows (ids >= n_valid are padding)
    uint64_t id_base;      // global id of base row 0
    int l2;                // 1: score = |b|^2 - 2<q,b> ; 0: score = -<q,b>
    float inv_scale;       // 1 / (scale_q * scale_b): accumulator -> true <q,b>
    const float *bnorm;    // |b|^2 per base row (L2 only)
    const float *thr;      // per-query threshold tau (candidates must be < tau)
    uint64_t *cand;        // [nq][kCap] keys: ordered(score)<<32 | local row id
    uint32_t *cand_count;  // [nq]
    uint32_t n_stages;     // B slab ring depth
};

// filter value of accumulator column i: IP the raw accumulator, L2 the accumulator minus |b|^2 * scale/2
template <bool kL2>
__device__ __forceinline__ float filt(uint32_t raw, const float *bn, int i, float half_scale) {
    return kL2 ? fmaf(-half_scale, bn[i], __uint_as_float(raw)) : __uint_as_float(raw);
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // FMNMX3
// 3-ary max tree over the 32 filter values of one accumulator slice (17 instructions); the inner nodes are kept so the
// rare hit can be located by walking down the tree instead of testing all 32 columns
struct MaxTree {
    float a[11];  // a[i] = max of columns 3i..3i+2 (a[10]: columns 30, 31)
    float b[4];   // b[g] = max of a[3g..3g+2] (b[3]: a[9], a[10])
    float m;
};
template <bool kL2>
__device__ __forceinline__ void build_tree(const uint32_t (&v)[32], const float *bn, float half_scale, MaxTree &t) {
#pragma unroll
    for (int i = 0; i < 10; ++i)
        t.a[i] = max3(filt<kL2>(v[3 * i], bn, 3 * i, half_scale), filt<kL2>(v[3 * i + 1], bn, 3 * i + 1, half_scale),
                      filt<kL2>(v[3 * i + 2], bn, 3 * i + 2, half_scale));
    t.a[10] = fmaxf(filt<kL2>(v[30], bn, 30, half_scale), filt<kL2>(v[31], bn, 31, half_scale));
#pragma unroll
    for (int g = 0; g < 3; ++g) t.b[g] = max3(t.a[3 * g], t.a[3 * g + 1], t.a[3 * g + 2]);
    t.b[3] = fmaxf(t.a[9], t.a[10]);
    t.m = fmaxf(max3(t.b[0], t.b[1], t.b[2]), t.b[3]);
}
// Dense path (first blocks, where a large share of the scores still beats the threshold): the lane's hits of this
// accumulator slice are counted first, ONE returning atomic reserves room for all of them, then predicated stores
// write the keys.  A returning atomic per hit (~700 cycles, and one divergent program position per column) made the
// small early blocks cost ~0.4 ms each regardless of their size.
template <bool kL2>
__device__ __forceinline__ void bulk_append(const uint32_t (&v)[32], const float *bn, float half_scale, float theta,
                                         float inv_scale, uint64_t row_first, uint64_t n_valid, uint32_t n_hits,
                                         uint32_t *count, uint64_t *my_cand) {
    uint32_t pos = atomicAdd(count, n_hits);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint64_t row = row_first + i;
        if (filt<kL2>(v[i], bn, i, half_scale) > theta && row < n_valid) {
            const float a = __uint_as_float(v[i]);
            const float sc = kL2 ? fmaf(-2.f * inv_scale, a, bn[i]) : -(a * inv_scale);
            if (pos < kCap) my_cand[pos] = (uint64_t(float_to_ordered(sc)) << 32) | uint32_t(row);
            ++pos;
        }
    }
}
template <bool kL2>
__device__ __forceinline__ uint32_t count_hits(const uint32_t (&v)[32], const float *bn, float half_scale, float theta,
                                               uint64_t row_first, uint64_t n_valid) {
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) n += (filt<kL2>(v[i], bn, i, half_scale) > theta && row_first + i < n_valid) ? 1u : 0u;
    return n;
}
// Sparse path of one lane (at most kStageSlots - n_st hits, checked by the caller): walk the max tree down to the columns
// that beat the threshold and STAGE them in the lane's shared-memory slots st[j * 32] (j < kStageSlots).  Nothing here
// waits on global memory: a returning atomic in one of the pair's 16 epilogue warps would delay the accumulator hand-back
// of the whole pair.
template <bool kL2>
__device__ __forceinline__ uint32_t stage_hits(const MaxTree &t, const uint32_t (&v)[32], const float *bn, float half_scale,
                                               float theta, float inv_scale, uint64_t row_first, uint64_t n_valid, uint64_t *st,
                                               uint32_t n_st) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (t.b[g] > theta) {
#pragma unroll
            for (int tt = 3 * g; tt < 3 * g + 3 && tt < 11; ++tt) {
                if (t.a[tt] > theta) {
#pragma unroll
                    for (int i = 3 * tt; i < 3 * tt + 3 && i < 32; ++i) {
                        if (filt<kL2>(v[i], bn, i, half_scale) > theta) {
                            const uint64_t row = row_first + i;
                            const float a = __uint_as_float(v[i]);
                            const float sc = kL2 ? fmaf(-2.f * inv_scale, a, bn[i]) : -(a * inv_scale);
                            if (row < n_valid && n_st < kStageSlots) {
                                st[n_st * 32] = (uint64_t(float_to_ordered(sc)) << 32) | uint32_t(row);
                                ++n_st;
                            }
                        }
                    }
                }
            }
        }
    }
    return n_st;
}
// hits of one 32-column accumulator slice of this lane: few -> staged in shared memory, many -> one bulk reservation
template <bool kL2>
__device__ __forceinline__ uint32_t take_hits(const MaxTree &t, const uint32_t (&v)[32], const float *bn, float half_scale,
                                              float theta, float inv_scale, uint64_t row_first, uint64_t n_valid, uint64_t *st,
                                              uint32_t n_st, uint32_t *count, uint64_t *my_cand) {
    const uint32_t n_hits = count_hits<kL2>(v, bn, half_scale, theta, row_first, n_valid);
    if (n_hits + n_st <= kStageSlots)
        return stage_hits<kL2>(t, v, bn, half_scale, theta, inv_scale, row_first, n_valid, st, n_st);
    bulk_append<kL2>(v, bn, half_scale, theta, inv_scale, row_first, n_valid, n_hits, count, my_cand);
    return n_st;
}
// Sparse blocks (everything after the first few): no hit counting at all - a warp gets here when ANY of its lanes has a
// hit (~1 lane in 100), and the exact 32-column count of the dense path would then cost the whole warp several times the
// max tree (measured: 512K-row block 5.35 ms -> 6.66 ms).  Hits are staged while there is room; the rare extra hit takes
// one returning atomic.
__device__ __noinline__ void direct_append(uint32_t *count, uint64_t *my_cand, uint64_t key) {
    const uint32_t pos = atomicAdd(count, 1u);
    if (pos < kCap) my_cand[pos] = key;
}
template <bool kL2>
__device__ __forceinline__ uint32_t take_hits_sparse(const MaxTree &t, const uint32_t (&v)[32], const float *bn, float half_scale,
                                                     float theta, float inv_scale, uint64_t row_first, uint64_t n_valid,
                                                     uint64_t *st, uint32_t n_st, uint32_t *count, uint64_t *my_cand) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (t.b[g] > theta) {
#pragma unroll
            for (int tt = 3 * g; tt < 3 * g + 3 && tt < 11; ++tt) {
                if (t.a[tt] > theta) {
#pragma unroll
                    for (int i = 3 * tt; i < 3 * tt + 3 && i < 32; ++i) {
                        if (filt<kL2>(v[i], bn, i, half_scale) > theta) {
                            const uint64_t row = row_first + i;
                            const float a = __uint_as_float(v[i]);
                            const float sc = kL2 ? fmaf(-2.f * inv_scale, a, bn[i]) : -(a * inv_scale);
                            if (row < n_valid) {
                                const uint64_t key = (uint64_t(float_to_ordered(sc)) << 32) | uint32_t(row);
                                if (n_st < kStageSlots) {
                                    st[n_st * 32] = key;
                                    ++n_st;
                                } else {
                                    direct_append(count, my_cand, key);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return n_st;
}
// one reservation per lane for everything it staged, then the copies
__device__ __noinline__ void stage_flush(const uint64_t *st, uint32_t n_st, uint32_t *count, uint64_t *my_cand) {
    if (n_st) {
        const uint32_t base = atomicAdd(count, n_st);
        for (uint32_t j = 0; j < n_st; ++j)
            if (base + j < kCap) my_cand[base + j] = st[j * 32];
    }
}

// kDense: the epilogue may meet many hits per accumulator slice (the first blocks of a query batch, where the threshold is
// still loose) and reserves list room in bulk; otherwise hits are rare and are staged one by one.
template <bool kL2, bool kDense>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
knn_gemm_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                       const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_bt,
                       const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // [A: a_bufs x a_bytes][B ring: n_stages x 16 KB][barriers][tmem slot][bnorm 2 x 256 floats]; identical in both CTAs
    unsigned char *smem_a = smem;
    unsigned char *smem_b = smem + size_t(p.a_bufs) * p.a_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_b + size_t(p.n_stages) * kSlabBytes);
    uint64_t *full = bars;                      // [n_stages]  used in the leader CTA only (both CTAs' TMA signal it)
    uint64_t *empty = bars + 16;                // [n_stages]  one per CTA, released by the leader's multicast commit
    uint64_t *a_full = bars + 32;               // [2]         leader only
    uint64_t *a_empty = bars + 34;              // [2]         per CTA
    uint64_t *tmem_full = bars + 36;            // [2]         per CTA (multicast commit)
    uint64_t *tmem_empty = bars + 38;           // [2]         leader only: 2 x kEpiWarps arrivals (both CTAs' epilogues)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);
    float *s_bn = reinterpret_cast<float *>(bars + 48);  // [2][256]
    uint64_t *s_stage = reinterpret_cast<uint64_t *>(s_bn + 2 * kPairN);  // [kEpiWarps][kStageSlots][32]

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();    // 0 = leader (issues every MMA), 1 = peer
    const uint32_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const uint32_t tail_bytes = kHalfN * p.tail_k * 2;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a_full[a], 1);
            mbar_init(&a_empty[a], 1);
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 2 * kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc2(tmem_slot, 512);   // the same warp of both CTAs allocates the pair's columns
    tc_fence_before();
    cluster_sync();                                            // barriers of both CTAs initialised before any remote use
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t chunks = (p.n_tiles + p.chunk_tiles - 1) / p.chunk_tiles;
    const uint32_t units = chunks * p.m_tiles;  // chunk-major: concurrently running pairs share the B chunk in L2

    if (warp == kEpiWarps) {
        // ===================== TMA producer (both CTAs; each loads its 128 query rows and its 128 base rows) ======
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, abuf = 0, a_ph = 0;  // a_ph: bit b = phase of a_empty[b]
            const uint32_t full_leader = mapa_u32(smem_u32(full), 0);
            const uint32_t a_full_leader = mapa_u32(smem_u32(a_full), 0);
            const uint32_t a_tx = p.n_full * kSlabBytes + tail_bytes;
            for (uint32_t u = cid; u < units; u += ncl) {
                const uint32_t c = u / p.m_tiles, m = u % p.m_tiles;
                mbar_wait(&a_empty[abuf], ((a_ph >> abuf) & 1u) ^ 1u);  // MMAs that read this buffer have retired
                a_ph ^= 1u << abuf;
                if (rank == 0) mbar_arrive_expect_tx(&a_full[abuf], 2 * a_tx);
                unsigned char *ah = smem_a + size_t(abuf) * p.a_bytes;
                const uint64_t qrow = uint64_t(m) * (2 * kTileM) + rank * kTileM;
                for (uint32_t j = 0; j < p.n_full; ++j)
                    tma_load_2d_pair(ah + size_t(j) * kSlabBytes, &map_q, 0, int(j * p.q_rows_pad + qrow), a_full_leader + abuf * 8);
                if (p.tail_k) tma_load_2d_pair(ah + size_t(p.n_full) * kSlabBytes, &map_qt, 0, int(qrow), a_full_leader + abuf * 8);
                const uint32_t t0 = c * p.chunk_tiles, t1 = min(p.n_tiles, t0 + p.chunk_tiles);
                for (uint32_t t = t0; t < t1; ++t) {
                    const uint64_t row0 = p.row_lo + uint64_t(t) * kPairN + rank * kHalfN;
                    for (uint32_t j = 0; j < p.n_full; ++j) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * kSlabBytes);
                        tma_load_2d_pair(smem_b + size_t(stage) * kSlabBytes, &map_b, 0, int(j * p.b_rows_pad + row0),
                                         full_leader + stage * 8);
                        if (++stage == p.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (p.tail_k) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * tail_bytes);
                        tma_load_2d_pair(smem_b + size_t(stage) * kSlabBytes, &map_bt, 0, int(row0), full_leader + stage * 8);
                        if (++stage == p.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                if (++abuf == p.a_bufs) abuf = 0;
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0, abuf = 0, a_ph = 0, acc = 0, acc_phase = 0;
            // descriptor words: low = (smem address >> 4) & 0x3FFF, high = SBO>>4 | version 1 <<14 | layout <<29
            const uint32_t hi_full = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t hi_tail = ((8u * p.tail_k * 2u) >> 4) | (1u << 14) | ((p.tail_k == 32 ? 4u : 6u) << 29);
            const uint32_t b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
            for (uint32_t u = cid; u < units; u += ncl) {
                const uint32_t c = u / p.m_tiles;
                mbar_wait(&a_full[abuf], (a_ph >> abuf) & 1u);
                a_ph ^= 1u << abuf;
                const uint32_t a_lo0 = (smem_u32(smem_a + size_t(abuf) * p.a_bytes) & 0x3FFFFu) >> 4;
                const uint32_t t0 = c * p.chunk_tiles, t1 = min(p.n_tiles, t0 + p.chunk_tiles);
                for (uint32_t t = t0; t < t1; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // both epilogues drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * kPairN;
                    for (uint32_t j = 0; j < p.n_full; ++j) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + stage * (kSlabBytes >> 4);
                        const uint32_t a_lo = a_lo0 + j * (kSlabBytes >> 4);
#pragma unroll
                        for (uint32_t k = 0; k < kSlabK / kUmmaK; ++k)
                            umma2_f16_lohi(tmem_d, a_lo + 2 * k, b_lo + 2 * k, hi_full, kIdesc2, (j | k) != 0 ? 1u : 0u);
                        umma2_commit(&empty[stage]);  // both CTAs' slab halves may be overwritten once these MMAs retire
                        if (++stage == p.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (p.tail_k) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + stage * (kSlabBytes >> 4);
                        const uint32_t a_lo = a_lo0 + p.n_full * (kSlabBytes >> 4);
                        const uint32_t first = p.n_full ? 1u : 0u;
                        umma2_f16_lohi(tmem_d, a_lo, b_lo, hi_tail, kIdesc2, first);
                        if (p.tail_k == 32) umma2_f16_lohi(tmem_d, a_lo + 2, b_lo + 2, hi_tail, kIdesc2, 1u);
                        umma2_commit(&empty[stage]);
                        if (++stage == p.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma2_commit(&tmem_full[acc]);
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
                umma2_commit(&a_empty[abuf]);
                if (++abuf == p.a_bufs) abuf = 0;
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> threshold filter =====================
        // warp w reads TMEM lanes 32*(w%4).. (its 32 queries) and the column quarter w/4 (64 base rows) of each tile:
        // 4 warps per scheduler keep the TMEM loads and the max trees of different warps overlapped
        uint32_t acc = 0, acc_phase = 0;
        const uint32_t quarter = warp & 3, part = warp >> 2;
        const uint32_t c0 = part * 64;
        const uint32_t row_in_pair = rank * kTileM + quarter * 32 + lane;
        const uint32_t tmem_empty_leader = mapa_u32(smem_u32(tmem_empty), 0);
        uint64_t *st = s_stage + size_t(warp) * (kStageSlots * 32) + lane;
        uint32_t n_st = 0;  // candidates staged by this lane
        for (uint32_t u = cid; u < units; u += ncl) {
            const uint32_t c = u / p.m_tiles, m = u % p.m_tiles;
            const uint32_t q = m * (2 * kTileM) + row_in_pair;
            const bool q_valid = q < p.nq;
            const float tau = q_valid ? p.thr[q] : -INFINITY;
            const float scale = 1.f / p.inv_scale;            // power of two
            const float theta_raw = -tau * scale;             // IP threshold on the raw accumulator
            const float half_scale = 0.5f * scale;            // L2
            const float theta_l2 = -tau * half_scale;
            uint64_t *my_cand = p.cand + uint64_t(q_valid ? q : 0) * kCap;
            const uint32_t t0 = c * p.chunk_tiles, t1 = min(p.n_tiles, t0 + p.chunk_tiles);
            for (uint32_t t = t0; t < t1; ++t) {
                const uint64_t row0 = p.row_lo + uint64_t(t) * kPairN;
                if (kL2) {
                    if (threadIdx.x < kPairN) {               // thread <-> column
                        const uint64_t r = row0 + threadIdx.x;
                        s_bn[acc * kPairN + threadIdx.x] = (r < p.n_valid) ? p.bnorm[r] : INFINITY;
                    }
                    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiWarps * 32) : "memory");
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kPairN + c0;
                const float *bn = s_bn + acc * kPairN + c0;
                {
                    // two 32-column accumulator slices in flight, one wait
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(taddr, v0);
                    tmem_ld32_nowait(taddr + 32, v1);
                    tmem_ld_wait(v0);  // one hardware wait covers both loads; the second statement only adds the
                    tmem_ld_wait(v1);  // register dependency for v1 (tcgen05.wait::ld with nothing outstanding is free)
                    // "does any of my accumulators beat the threshold": a 3-input max tree (FMNMX3) over the raw, scaled
                    // values and ONE compare; the slices are only indexed with constants so they stay in registers.
                    //   IP: -<q,b> < tau            <=>  acc > -tau * scale
                    //   L2: |b|^2 - 2<q,b> < tau    <=>  acc - |b|^2 * scale/2 > -tau * scale/2
                    const float theta = kL2 ? theta_l2 : theta_raw;
                    MaxTree t;
                    build_tree<kL2>(v0, bn, half_scale, t);
                    if (t.m > theta)  // rare per lane: a few candidates per query per block
                        n_st = kDense ? take_hits<kL2>(t, v0, bn, half_scale, theta, p.inv_scale, row0 + c0, p.n_valid, st, n_st,
                                                       p.cand_count + q, my_cand)
                                      : take_hits_sparse<kL2>(t, v0, bn, half_scale, theta, p.inv_scale, row0 + c0, p.n_valid, st,
                                                              n_st, p.cand_count + q, my_cand);
                    build_tree<kL2>(v1, bn + 32, half_scale, t);
                    if (t.m > theta)
                        n_st = kDense ? take_hits<kL2>(t, v1, bn + 32, half_scale, theta, p.inv_scale, row0 + c0 + 32, p.n_valid, st,
                                                       n_st, p.cand_count + q, my_cand)
                                      : take_hits_sparse<kL2>(t, v1, bn + 32, half_scale, theta, p.inv_scale, row0 + c0 + 32,
                                                              p.n_valid, st, n_st, p.cand_count + q, my_cand);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                if (n_st > kStageSlots / 2) {  // after the accumulator went back: keep room for the next tile's hits
                    stage_flush(st, n_st, p.cand_count + q, my_cand);
                    n_st = 0;
                }
            }
            if (__any_sync(0xffffffffu, n_st > 0)) {  // the lane's query changes with the unit
                stage_flush(st, n_st, p.cand_count + q, my_cand);
                n_st = 0;
            }
        }
    }
    tc_fence_before();
    cluster_sync();   // the peer's shared memory and TMEM stay alive until the leader's last MMA has been consumed
    if (warp == kEpiWarps + 1) tmem_dealloc2(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// warp-level bitonic sort of n <= P (power of two) keys in shared memory, ascending
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_bitonic_sort(uint64_t *s, uint32_t P, uint32_t lane) {
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < P; i += 32) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const uint64_t a = s[i], b = s[x];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        s[i] = b;
                        s[x] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}
__device__ __forceinline__ uint32_t next_pow2(uint32_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

// K2s: one warp per query: keep the kprime best candidates, tighten tau, flag overflow.  No sorting (K3 re-scores and
// sorts the survivors anyway): the kprime-th smallest score is found by bisection on its ordered 32-bit pattern with the
// scores held in registers (32 probes, one redux.sync each), then the list is compacted in place to the entries at or
// below it.  Equal scores stay together, so a list keeps a few more than kprime entries only on exact FP32 ties.
__global__ void __launch_bounds__(128) knn_select_kernel(uint64_t *cand, uint32_t *cand_count, float *thr, uint32_t *overflow,
                                                          uint32_t nq, uint32_t kprime) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    uint32_t n = cand_count[q];
    if (n > kCap) {  // more survivors than the list holds: the exact scan will redo this query
        if (lane == 0) overflow[q] = 1;
        n = kCap;
    }
    if (n <= kprime) return;  // nothing to drop (tau stays)
    uint64_t *list = cand + uint64_t(q) * kCap;
    const uint32_t *hi_words = reinterpret_cast<const uint32_t *>(list) + 1;  // ordered(score) of entry i = hi_words[2 i]
    constexpr uint32_t kPerLane = kCap / 32;
    uint32_t sc[kPerLane];
    const uint32_t nj = (n + 31) >> 5;
#pragma unroll
    for (uint32_t j = 0; j < kPerLane; ++j) {
        const uint32_t i = j * 32 + lane;
        sc[j] = (j < nj && i < n) ? hi_words[2 * i] : 0xFFFFFFFFu;  // padding is above every probe
    }
    uint32_t lo = 0, hi = 0xFFFFFFFEu;  // smallest v with #{score <= v} >= kprime
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t c = 0;
#pragma unroll
        for (uint32_t j = 0; j < kPerLane; ++j) {
            if (j >= nj) break;  // warp-uniform
            c += (sc[j] <= mid) ? 1u : 0u;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= kprime) hi = mid;
        else lo = mid + 1;
    }
    const uint32_t v = lo;
    // in-place compaction: round r reads entries [32r, 32r+32) and writes at most that many entries at positions
    // <= 32r + 31, all of which were read in this or an earlier round (the ballot orders the reads before the writes)
    uint32_t kept = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t key = (i < n) ? list[i] : ~0ull;
        const bool keep = (i < n) && uint32_t(key >> 32) <= v;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) list[kept + __popc(m & lanemask_lt())] = key;
        kept += __popc(m);
    }
    if (lane == 0) {
        cand_count[q] = kept;
        thr[q] = ordered_to_float(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: exact FP32 re-rank (reference lane order, same arithmetic as the search kernel) + certificate
// ---------------------------------------------------------------------------------------------------------------
// 4 lanes per row, 8 rows per warp step; returns the lane-ordered sum (IP: dot, L2: squared distance) in lane 4g.
// The row is streamed in chunks of 8 steps (8 x 16 B per lane, all loads issued before the first use), so a warp keeps
// 8 rows x 512 B in flight instead of one 64-byte segment per row: K3 is a gather of k' random rows per query and was
// latency-bound with one load per step (6.4 ms per 131072-query batch against a ~2.5 ms HBM floor).  The arithmetic and
// its order are unchanged.
template <bool kIP>
__device__ __forceinline__ float exact_score(const float *__restrict__ a, const float *__restrict__ b, uint32_t dim,
                                             uint32_t t) {
    const float4 *ap = reinterpret_cast<const float4 *>(a) + t;
    const float4 *bp = reinterpret_cast<const float4 *>(b) + t;
    const uint32_t n16 = dim >> 4;
    const bool tail8 = (dim & 15u) != 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t s0 = 0; s0 < n16; s0 += 8) {
        float4 vv[8], qq[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
            if (s0 + j < n16) vv[j] = __ldg(ap + 4 * (s0 + j));
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
            if (s0 + j < n16) qq[j] = __ldg(bp + 4 * (s0 + j));
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            if (s0 + j < n16) {
                const float4 v = vv[j], q = qq[j];
                if (kIP) {
                    acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, q.x));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, q.y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, q.z));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, q.w));
                } else {
                    const float dx = __fsub_rn(v.x, q.x), dy = __fsub_rn(v.y, q.y), dz = __fsub_rn(v.z, q.z), dw = __fsub_rn(v.w, q.w);
                    acc.x = __fadd_rn(acc.x, __fmul_rn(dx, dx));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(dy, dy));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(dz, dz));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(dw, dw));
                }
            }
        }
    }
    float4 m;
    m.x = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.x, 2), acc.x);
    m.y = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.y, 2), acc.y);
    m.z = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.z, 2), acc.z);
    m.w = __fadd_rn(__shfl_down_sync(0xffffffffu, acc.w, 2), acc.w);
    if (tail8 && t < 2) {
        const float4 v = ap[4 * n16], q = bp[4 * n16];
        if (kIP) {
            m.x = __fmaf_rn(v.x, q.x, m.x);
            m.y = __fmaf_rn(v.y, q.y, m.y);
            m.z = __fmaf_rn(v.z, q.z, m.z);
            m.w = __fmaf_rn(v.w, q.w, m.w);
        } else {
            const float dx = __fsub_rn(v.x, q.x), dy = __fsub_rn(v.y, q.y), dz = __fsub_rn(v.z, q.z), dw = __fsub_rn(v.w, q.w);
            m.x = __fmaf_rn(dx, dx, m.x);
            m.y = __fmaf_rn(dy, dy, m.y);
            m.z = __fmaf_rn(dz, dz, m.z);
            m.w = __fmaf_rn(dw, dw, m.w);
        }
    }
    float4 f;
    f.x = __fadd_rn(__shfl_down_sync(0xffffffffu, m.x, 1), m.x);
    f.y = __fadd_rn(__shfl_down_sync(0xffffffffu, m.y, 1), m.y);
    f.z = __fadd_rn(__shfl_down_sync(0xffffffffu, m.z, 1), m.z);
    f.w = __fadd_rn(__shfl_down_sync(0xffffffffu, m.w, 1), m.w);
    return __fadd_rn(__fadd_rn(f.x, f.y), __fadd_rn(f.z, f.w));
}

// one warp per query.  out_ids/out_dists: [nq][K]; need_exact[q] = 1 when the certificate fails.
template <bool kIP>
__global__ void __launch_bounds__(128) knn_rerank_kernel(const float *__restrict__ base, const float *__restrict__ queries,
                                                          uint32_t dim, uint64_t id_base, const uint64_t *__restrict__ cand,
                                                          const uint32_t *__restrict__ cand_count, const float *__restrict__ thr,
                                                          const uint32_t *__restrict__ overflow, float eps_factor,
                                                          float max_bnorm2, uint32_t nq, uint32_t K, uint64_t n_base,
                                                          uint32_t *__restrict__ out_ids, float *__restrict__ out_dists,
                                                          uint32_t *__restrict__ need_exact) {
    extern __shared__ __align__(16) unsigned char sm[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 2, t = lane & 3;
    uint64_t *s = reinterpret_cast<uint64_t *>(sm) + size_t(warp) * kCap;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    const uint32_t n = min(cand_count[q], kCap);
    const float *qv = queries + uint64_t(q) * dim;
    const uint64_t *list = cand + uint64_t(q) * kCap;
    const uint32_t P = max(32u, next_pow2(n));
    for (uint32_t i = lane; i < P; i += 32) s[i] = ~0ull;
    __syncwarp();
    for (uint32_t r0 = 0; r0 < n; r0 += 8) {
        const uint32_t r = r0 + grp;
        const bool valid = r < n;
        const uint32_t row = uint32_t(list[valid ? r : n - 1]);
        const float e = exact_score<kIP>(base + uint64_t(row) * dim, qv, dim, t);
        if (valid && t == 0) s[r] = make_key(kIP ? -e : e, row) >> 1;  // ordered(score)<<31 | row  (no flag bit needed)
    }
    __syncwarp();
    warp_bitonic_sort(s, P, lane);
    // |q| for the error bound
    const float qq = __shfl_sync(0xffffffffu, exact_score<true>(qv, qv, dim, t), 0);
    const float qn = sqrtf(qq);
    bool complete = (n_base <= n);  // every base row was a survivor
    if (!complete && n >= K) {
        // tau bounds the APPROXIMATE score of every non-survivor from below.  IP: score = -<q,b>.  L2: the filter's
        // score is |b|^2 - 2<q,b>, i.e. the true squared distance minus |q|^2.
        const float tau = thr[q] + (kIP ? 0.f : qq);
        const float eps = eps_factor * qn + 1e-5f * (qq + max_bnorm2);  // FP16 rounding bound + FP32 evaluation slack
        const float eK = ordered_to_float(uint32_t(s[K - 1] >> 31));
        // n >= K survivors whose K-th exact score is clear of every non-survivor's lower bound: the top K is complete
        complete = (eK < tau - eps) && overflow[q] == 0;
    }
    for (uint32_t i = lane; i < K; i += 32) {
        if (i < n) {
            const uint64_t k = s[i];
            const float sc = ordered_to_float(uint32_t(k >> 31));
            out_ids[uint64_t(q) * K + i] = uint32_t(id_base + (k & 0x7fffffffull));
            out_dists[uint64_t(q) * K + i] = kIP ? -sc : sc;  // mips distances are written as +ip (compute_groundtruth.cpp:438-441)
        } else {
            out_ids[uint64_t(q) * K + i] = 0xFFFFFFFFu;
            out_dists[uint64_t(q) * K + i] = 0.f;
        }
    }
    if (lane == 0) need_exact[q] = complete ? 0u : 1u;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact scan for the (rare) queries without a certificate: one CTA per flagged query, FP32 lane-ordered scores,
// per-warp bounded sorted lists merged at the end.  K <= 128.
// ---------------------------------------------------------------------------------------------------------------
template <bool kIP>
__global__ void __launch_bounds__(256) knn_exact_scan_kernel(const float *__restrict__ base, uint64_t n, uint64_t id_base,
                                                              const float *__restrict__ queries, uint32_t dim,
                                                              const uint32_t *__restrict__ flagged, uint32_t n_flagged,
                                                              uint32_t K, uint32_t *__restrict__ out_ids,
                                                              float *__restrict__ out_dists) {
    extern __shared__ __align__(16) unsigned char sm[];
    // per warp: 256 keys (K best + 128 staging slots, sorted together when the staging area fills)
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 2, t = lane & 3;
    const uint32_t nwarps = blockDim.x >> 5;
    uint64_t *s = reinterpret_cast<uint64_t *>(sm) + size_t(warp) * 256;
    for (uint32_t f = blockIdx.x; f < n_flagged; f += gridDim.x) {
        const uint32_t q = flagged[f];
        const float *qv = queries + uint64_t(q) * dim;
        for (uint32_t i = lane; i < 256; i += 32) s[i] = ~0ull;
        __syncwarp();
        uint32_t fill = 0;            // staging entries in s[128 .. 128+fill)
        uint64_t worst = ~0ull;       // K-th best so far of this warp
        for (uint64_t r0 = uint64_t(warp) * 8; r0 < n; r0 += uint64_t(nwarps) * 8) {
            const uint64_t r = r0 + grp;
            const bool valid = r < n;
            const float e = exact_score<kIP>(base + (valid ? r : n - 1) * dim, qv, dim, t);
            const uint64_t key = make_key(kIP ? -e : e, uint32_t(r)) >> 1;
            const bool want = valid && t == 0 && key < worst;
            const uint32_t mask = __ballot_sync(0xffffffffu, want);
            if (want) s[128 + fill + __popc(mask & lanemask_lt())] = key;
            fill += __popc(mask);
            __syncwarp();
            if (fill > 120) {
                warp_bitonic_sort(s, 256, lane);
                for (uint32_t i = 128 + lane; i < 256; i += 32) s[i] = ~0ull;
                __syncwarp();
                worst = s[K - 1];
                fill = 0;
            }
        }
        warp_bitonic_sort(s, 256, lane);
        __syncthreads();
        // merge the per-warp lists: warp 0 gathers the K best of each warp and sorts (nwarps*K <= 1024)
        if (warp == 0) {
            uint64_t *all = reinterpret_cast<uint64_t *>(sm) + size_t(nwarps) * 256;
            const uint32_t tot = nwarps * 128;
            for (uint32_t i = lane; i < 1024; i += 32) {
                const uint32_t w = i / 128, j = i % 128;
                all[i] = (i < tot && j < K) ? (reinterpret_cast<uint64_t *>(sm))[size_t(w) * 256 + j] : ~0ull;
            }
            __syncwarp();
            warp_bitonic_sort(all, 1024, lane);
            for (uint32_t i = lane; i < K; i += 32) {
                const uint64_t k = all[i];
                if (k != ~0ull) {
                    const float sc = ordered_to_float(uint32_t(k >> 31));
                    out_ids[uint64_t(q) * K + i] = uint32_t(id_base + (k & 0x7fffffffull));
                    out_dists[uint64_t(q) * K + i] = kIP ? -sc : sc;
                } else {
                    out_ids[uint64_t(q) * K + i] = 0xFFFFFFFFu;
                    out_dists[uint64_t(q) * K + i] = 0.f;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void fill_f32_kernel(float *p, uint64_t n, float v) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// K4: merge G sorted lists per query (output convention: ids + dists, IP dists as +ip) into the global top-K
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) knn_merge_kernel(const uint32_t *__restrict__ part_ids, const float *__restrict__ part_dists,
                                                         uint32_t G, uint64_t nq, uint32_t K, int ip,
                                                         uint32_t *__restrict__ out_ids, float *__restrict__ out_dists) {
    extern __shared__ __align__(16) unsigned char sm[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *s = reinterpret_cast<uint64_t *>(sm) + size_t(warp) * 1024;
    const uint64_t q = uint64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    const uint32_t tot = G * K;
    const uint32_t P = max(32u, next_pow2(tot));
    for (uint32_t i = lane; i < P; i += 32) {
        uint64_t key = ~0ull;
        if (i < tot) {
            const uint32_t g = i / K, j = i % K;
            const uint32_t id = part_ids[(uint64_t(g) * nq + q) * K + j];
            const float d = part_dists[(uint64_t(g) * nq + q) * K + j];
            if (id != 0xFFFFFFFFu) key = (uint64_t(float_to_ordered(ip ? -d : d)) << 32) | id;
        }
        s[i] = key;
    }
    __syncwarp();
    warp_bitonic_sort(s, P, lane);
    for (uint32_t i = lane; i < K; i += 32) {
        const uint64_t k = (i < P) ? s[i] : ~0ull;
        if (k != ~0ull) {
            const float sc = ordered_to_float(uint32_t(k >> 32));
            out_ids[q * K + i] = uint32_t(k);
            out_dists[q * K + i] = ip ? -sc : sc;
        } else {
            out_ids[q * K + i] = 0xFFFFFFFFu;
            out_dists[q * K + i] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// array [outer][width] halves viewed as a 2-D tensor {width, outer}; box {width, 128}; swizzle = row bytes
// (full slabs: width 64 -> 128B, outer = n_full*rows_pad; tail: width 16/32 -> 32B/64B, outer = rows_pad)
static rg_status make_slab_map(CUtensorMap *map, const __half *ptr, uint32_t width, uint64_t outer) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(RG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {cuuint64_t(width), cuuint64_t(outer)};
    cuuint64_t strides[1] = {cuuint64_t(width) * sizeof(__half)};
    cuuint32_t box[2] = {cuuint32_t(width), cuuint32_t(kTileN)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = width == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : width == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
    return RG_OK;
}

static float pow2_scale(float maxabs) {  // power of two that brings max|x| into (128, 256]
    if (!(maxabs > 0.f) || !std::isfinite(maxabs)) return 1.f;
    int e;
    std::frexp(maxabs, &e);  // maxabs = m * 2^e, m in [0.5, 1)
    return std::ldexp(1.f, 8 - e);
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

// Per-device scratch cache.  rg_knn_exact_device used to cudaMalloc / cudaFree ~4.5 GB (10M-row shard) on every call, which
// serialises with everything else on the device and showed up as 0.25-0.7 s outliers between back-to-back calls; the
// buffers are now kept (grow-only) until rg_knn_release_scratch().  One call at a time per device (the mutex is held for
// the whole call; the tools run one host thread per GPU).
struct DeviceScratch {
    std::mutex mu;
    enum { kSlots = 24 };
    void *ptr[kSlots] = {nullptr};
    uint64_t cap[kSlots] = {0};
    // FP16 copy of the last base shard converted on this device (valid while the caller keeps the rows unchanged: only
    // reused inside one rg_knn_exact_sharded call, which runs K2 once per query segment against the same shard)
    const float *prep_base = nullptr;
    uint64_t prep_n = 0;
    uint32_t prep_dim = 0;
    float prep_scale = 0.f, prep_max_bnorm2 = 0.f;
    template <typename T>
    cudaError_t get(int slot, T **out, uint64_t count) {
        const uint64_t bytes = std::max<uint64_t>(count, 1) * sizeof(T);
        if (cap[slot] < bytes) {
            if (ptr[slot]) cudaFree(ptr[slot]);
            ptr[slot] = nullptr;
            cap[slot] = 0;
            cudaError_t e = cudaMalloc(&ptr[slot], bytes);
            if (e != cudaSuccess) return e;
            cap[slot] = bytes;
        }
        *out = static_cast<T *>(ptr[slot]);
        return cudaSuccess;
    }
    void release() {
        for (int i = 0; i < kSlots; ++i) {
            if (ptr[i]) cudaFree(ptr[i]);
            ptr[i] = nullptr;
            cap[i] = 0;
        }
        prep_base = nullptr;
    }
};
static DeviceScratch &device_scratch(int dev) {
    static DeviceScratch all[64];
    return all[dev & 63];
}
void release_scratch(int dev) {
    DeviceScratch &d = device_scratch(dev);
    std::lock_guard<std::mutex> lock(d.mu);
    d.release();
}

__global__ void gather_rows_kernel(const float *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n_idx, uint32_t dim,
                                   float *__restrict__ dst) {
    const uint32_t c4 = dim >> 2;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < uint64_t(n_idx) * c4; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t r = uint32_t(i / c4), c = uint32_t(i % c4);
        reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(src + uint64_t(idx[r]) * dim)[c];
    }
}
// rows of (src_ids, src_d) [n_idx][K] -> rows idx[r] of (dst_ids, dst_d)
__global__ void scatter_results_kernel(const uint32_t *__restrict__ src_ids, const float *__restrict__ src_d,
                                       const uint32_t *__restrict__ idx, uint32_t n_idx, uint32_t K,
                                       uint32_t *__restrict__ dst_ids, float *__restrict__ dst_d) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < uint64_t(n_idx) * K; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t r = uint32_t(i / K), c = uint32_t(i % K);
        dst_ids[uint64_t(idx[r]) * K + c] = src_ids[i];
        dst_d[uint64_t(idx[r]) * K + c] = src_d[i];
    }
}
// appends q_off + i for every flagged i (the order is irrelevant)
__global__ void append_flags_kernel(const uint32_t *__restrict__ flags, uint32_t n, uint32_t q_off, uint32_t *__restrict__ list,
                                    uint32_t *__restrict__ count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) list[atomicAdd(count, 1u)] = q_off + i;
}

static uint64_t env_u64(const char *name, uint64_t dflt, uint64_t lo, uint64_t hi) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    const long long v = atoll(e);
    return uint64_t(v < (long long)lo ? lo : (v > (long long)hi ? hi : v));
}

// Threshold schedule.  The base is visited in blocks of growing size; after each block K2s tightens the per-query threshold
// tau and compacts the candidate list to the entries at or below it.
//   conservative: tau = the k'-th best score seen so far.  Always certifiable, but the expected number of list insertions
//                 at row i is k'/i: with k' = 150 the epilogue's hit path stays busy until ~1M rows have been seen, and a
//                 1.25M-row shard (C4 on 8 GPUs) never gets out of it (round 1: 663 TFLOP/s per GPU against 1090 on one).
//   optimistic  : tau = the r-th best score seen so far with r << k' while few rows have been seen.  Rows in storage order
//                 are treated as a sample: the r-th best of `seen` rows has expected rank r n / seen in the whole
//                 shard, c_eff = r n / (seen k') times deeper than the k'-th best that is needed in the end.  r is the
//                 smallest of 4, 8, 16, 32, ... whose c_eff still clears a safety factor (8, 4, 2.5, then 2: the rank
//                 estimate of an r-th order statistic has a relative spread of 1/sqrt(r)), capped at k'.  Insertions
//                 drop from k'/i to r/i per row - a 256-row first block, then a dozen hits per block until ~16K rows -
//                 so every block after the first few runs at the sparse epilogue's speed, and the last select always
//                 uses k' so that K3 re-scores k' survivors, not more.  tau only ever decreases and every non-survivor
//                 was rejected against a tau >= the final one, so the K3 certificate (exact_K < tau - eps) is exactly
//                 as strong as before; what changes is that a query whose estimate was too tight ends with too few
//                 survivors and FAILS the certificate (expected: a few per 10^4 on rows in arbitrary order).  Those
//                 queries are re-run with the conservative schedule (and only what fails that goes to the exact FP32
//                 scan), so the result is exact for any row order - the order only decides how many queries need the
//                 second pass.
struct PassOutput {
    uint32_t *ids;
    float *dists;
    uint32_t *flag_list;   // pass-local query indices that failed the certificate
    uint32_t *flag_count;
};

rg_status knn_device(const float *d_base, uint64_t n, uint64_t id_base, const float *d_queries, uint64_t nq, uint32_t dim,
                     int metric, uint32_t K, uint32_t *d_ids, float *d_dists, cudaStream_t st, uint64_t *stats, bool reuse_base) {
    if (!d_base || !d_queries || !d_ids || !d_dists) return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: null argument");
    if (dim == 0 || dim % 8 != 0 || dim > kMaxSlabs * kSlabK)
        return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: dim must be a multiple of 8 in [8, %d]", kMaxSlabs * kSlabK);
    if (K == 0 || K > 128) return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: K must be in [1, 128]");
    if (n == 0 || n >= (1ull << 31)) return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: shard size must be in [1, 2^31)");
    if (nq >= (1ull << 32)) return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: too many queries in one call");
    if (metric != RG_METRIC_L2 && metric != RG_METRIC_INNER_PRODUCT && metric != RG_METRIC_COSINE)
        return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: unsupported metric %d", metric);
    if (nq == 0) return RG_OK;
    const bool ip = metric != RG_METRIC_L2;
    // K split: full 64-wide slabs + one narrower tail slab (16 or 32 wide) when that saves tensor work
    uint32_t n_full = dim / kSlabK, tail_k = 0;
    {
        const uint32_t rem = dim % kSlabK;
        if (rem > 32) n_full += 1;          // 33..63 left: pad to a full slab
        else if (rem > 16) tail_k = 32;
        else if (rem > 0) tail_k = 16;
    }
    const uint32_t tail_bytes = kTileN * tail_k * 2;
    const uint32_t a_bytes = (n_full * kSlabBytes + tail_bytes + 1023) / 1024 * 1024;
    const uint64_t b_rows_pad = (n + kPairN - 1) / kPairN * kPairN;
    const uint32_t kprime = std::min<uint32_t>(256, std::max<uint32_t>(K + K / 2, K + 32));
    // queries per batch: the candidate lists take 8 KB per query (1 GB at 131072); larger batches give the small early
    // blocks 4x the M tiles of round 1's 32768 and quarter the number of per-batch launches
    const uint64_t q_batch = std::min<uint64_t>(env_u64("RG_KNN_QBATCH", 131072, 256, 1u << 20), (nq + 255) / 256 * 256);
    const uint64_t q_rows_pad = (q_batch + 2 * kTileM - 1) / (2 * kTileM) * (2 * kTileM);
    const bool optimistic = env_u64("RG_KNN_OPTIMISTIC", 1, 0, 1) != 0;
    const uint64_t r_min = env_u64("RG_KNN_RMIN", 4, 1, 256);
    const uint64_t first_rows = env_u64("RG_KNN_FIRST_ROWS", 256, 256, kCap) / kPairN * kPairN;
    int dev = 0, sms = 0, smem_max = 0;
    RG_CUDA_OK(cudaGetDevice(&dev));
    RG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RG_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    const bool trace = std::getenv("RG_KNN_TRACE") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rg_knn] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
        t_start = now;
    };
    DeviceScratch &sc = device_scratch(dev);
    std::lock_guard<std::mutex> scratch_lock(sc.mu);
    __half *b16 = nullptr, *q16 = nullptr, *b16t = nullptr, *q16t = nullptr;
    float *bnorm = nullptr, *thr = nullptr;
    uint32_t *scal = nullptr, *cand_count = nullptr, *overflow = nullptr, *need_exact = nullptr, *flag_list = nullptr;
    uint64_t *cand = nullptr;
    RG_CUDA_OK(sc.get(0, &b16, uint64_t(n_full) * b_rows_pad * kSlabK));
    RG_CUDA_OK(sc.get(1, &q16, uint64_t(n_full) * q_rows_pad * kSlabK));
    RG_CUDA_OK(sc.get(2, &b16t, b_rows_pad * std::max<uint32_t>(tail_k, 16)));
    RG_CUDA_OK(sc.get(3, &q16t, q_rows_pad * std::max<uint32_t>(tail_k, 16)));
    RG_CUDA_OK(sc.get(4, &bnorm, n));
    RG_CUDA_OK(sc.get(5, &thr, q_batch));
    RG_CUDA_OK(sc.get(6, &scal, 8));
    RG_CUDA_OK(sc.get(7, &cand_count, q_batch));
    RG_CUDA_OK(sc.get(8, &overflow, q_batch));
    RG_CUDA_OK(sc.get(9, &need_exact, q_batch));
    RG_CUDA_OK(sc.get(10, &flag_list, nq));
    RG_CUDA_OK(sc.get(11, &cand, q_batch * kCap));

    lap("scratch");
    // scal[0] = max|b| bits, scal[1] = max |b|^2 bits, scal[2] = max|q| bits, scal[3] / scal[4] = #flagged (pass 1 / pass 2)
    const bool have_base = reuse_base && sc.prep_base == d_base && sc.prep_n == n && sc.prep_dim == dim;
    sc.prep_base = nullptr;  // invalid while the buffers are being rewritten
    RG_CUDA_OK(cudaMemsetAsync(scal, 0, 8 * sizeof(uint32_t), st));
    if (!have_base) absmax_kernel<<<sms * 8, 256, 0, st>>>(d_base, n * dim, scal + 0);
    absmax_kernel<<<sms * 8, 256, 0, st>>>(d_queries, nq * dim, scal + 2);
    uint32_t h_scal[8];
    RG_CUDA_OK(cudaMemcpyAsync(h_scal, scal, sizeof(h_scal), cudaMemcpyDeviceToHost, st));
    RG_CUDA_OK(cudaStreamSynchronize(st));
    float max_b, max_q, max_bnorm2;
    memcpy(&max_b, &h_scal[0], 4);
    memcpy(&max_q, &h_scal[2], 4);
    const float scale_b = have_base ? sc.prep_scale : pow2_scale(max_b), scale_q = pow2_scale(max_q);
    if (have_base) {
        max_bnorm2 = sc.prep_max_bnorm2;
    } else {
        to_half_slabs_kernel<<<sms * 8, 256, 0, st>>>(d_base, n, dim, b_rows_pad, n_full, tail_k, scale_b, b16, b16t, bnorm, scal + 1);
        RG_CUDA_OK(cudaMemcpyAsync(h_scal, scal, sizeof(h_scal), cudaMemcpyDeviceToHost, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        memcpy(&max_bnorm2, &h_scal[1], 4);
    }
    sc.prep_base = d_base;
    sc.prep_n = n;
    sc.prep_dim = dim;
    sc.prep_scale = scale_b;
    sc.prep_max_bnorm2 = max_bnorm2;
    // FP16 rounding: |fl(x) - x| <= 2^-11 |x| per operand -> |<q,b>~ - <q,b>| <= (2^-10 + 2^-22) |q| |b|; 2% slack
    // covers FP32 accumulation; L2 scores carry the factor 2 of -2<q,b>.
    const float eps_factor = 1.02f * std::ldexp(1.f, -10) * std::sqrt(max_bnorm2) * (ip ? 1.f : 2.f);

    lap("base conversion");
    CUtensorMap map_q, map_b, map_qt, map_bt;
    rg_status s = RG_OK;
    if (n_full) {
        if ((s = make_slab_map(&map_b, b16, kSlabK, uint64_t(n_full) * b_rows_pad)) != RG_OK) return s;
        if ((s = make_slab_map(&map_q, q16, kSlabK, uint64_t(n_full) * q_rows_pad)) != RG_OK) return s;
    }
    if ((s = make_slab_map(&map_bt, b16t, tail_k ? tail_k : 16, b_rows_pad)) != RG_OK) return s;
    if ((s = make_slab_map(&map_qt, q16t, tail_k ? tail_k : 16, q_rows_pad)) != RG_OK) return s;
    if (!n_full) {  // never dereferenced by the kernel, but must be valid kernel arguments
        map_b = map_bt;
        map_q = map_qt;
    }

    // shared-memory budget: the resident query tile is double buffered across work units when the B ring still
    // gets >= 4 stages; the ring takes the rest
    const size_t misc = 48 * 8 + 2 * kPairN * sizeof(float) + size_t(kEpiWarps) * kStageSlots * 32 * 8 + 1024;
    uint32_t a_bufs = 2;
    if (size_t(smem_max) < 2 * size_t(a_bytes) + 4 * size_t(kSlabBytes) + misc) a_bufs = 1;
    if (size_t(smem_max) < size_t(a_bufs) * a_bytes + 2 * size_t(kSlabBytes) + misc)
        return fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: dim %u leaves no room for the operand ring", dim);
    uint32_t n_stages = uint32_t((size_t(smem_max) - misc - size_t(a_bufs) * a_bytes) / kSlabBytes);
    n_stages = std::min<uint32_t>(n_stages, 16);
    const size_t gemm_smem = size_t(a_bufs) * a_bytes + size_t(n_stages) * kSlabBytes + misc;
    RG_CUDA_OK(cudaFuncSetAttribute(knn_gemm_filter_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm_smem)));
    RG_CUDA_OK(cudaFuncSetAttribute(knn_gemm_filter_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm_smem)));
    RG_CUDA_OK(cudaFuncSetAttribute(knn_gemm_filter_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm_smem)));
    RG_CUDA_OK(cudaFuncSetAttribute(knn_gemm_filter_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gemm_smem)));
    // persistent grid: as many CTA pairs as the device can co-schedule (one CTA per SM, pairs on one TPC); the
    // occupancy query costs tens of milliseconds, so its answer is cached per (device, shared-memory size)
    uint32_t max_pairs = uint32_t(sms) / 2;
    {
        static std::mutex mu;
        static std::vector<std::pair<uint64_t, uint32_t>> cache;
        const uint64_t key = (uint64_t(dev) << 32) | uint64_t(gemm_smem);
        std::lock_guard<std::mutex> lock(mu);
        bool hit = false;
        for (auto &e : cache)
            if (e.first == key) {
                max_pairs = e.second;
                hit = true;
            }
        if (!hit) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(unsigned(sms) / 2 * 2);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = gemm_smem;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2;
            attr.val.clusterDim.y = 1;
            attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            int active = 0;
            if (cudaOccupancyMaxActiveClusters(&active, knn_gemm_filter_kernel<false, true>, &cfg) == cudaSuccess && active > 0)
                max_pairs = std::min<uint32_t>(max_pairs, uint32_t(active));
            else
                (void)cudaGetLastError();
            cache.emplace_back(key, max_pairs);
        }
    }
    RG_CUDA_OK(cudaFuncSetAttribute(knn_rerank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kCap * 8));
    RG_CUDA_OK(cudaFuncSetAttribute(knn_rerank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kCap * 8));
    const size_t scan_smem = (8 * 256 + 1024) * 8;
    RG_CUDA_OK(cudaFuncSetAttribute(knn_exact_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(scan_smem)));
    RG_CUDA_OK(cudaFuncSetAttribute(knn_exact_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(scan_smem)));
    lap("tensor maps + attributes");

    // RG_KNN_GROWTH = g (2..8): a block is (g-1) x everything seen so far.  RG_KNN_DENSE_ROWS: blocks that start before this
    // many rows use the dense-hit epilogue (measured per block at 10M x 32768, profiles/r01_launches_knn_*)
    const uint64_t growth_opt = env_u64("RG_KNN_GROWTH", 0, 0, 8);
    const uint64_t dense_rows = env_u64("RG_KNN_DENSE_ROWS", 4096, 0, 1ull << 40);
    uint64_t launches = 5;

    // One pass over `pq` [pnq][dim] (device): results to out.ids/out.dists [pnq][K]; queries without a certificate are
    // appended to out.flag_list.  No host synchronisation inside.
    auto run_pass = [&](const float *pq, uint64_t pnq, bool opt, const PassOutput &out) -> rg_status {
        const uint64_t growth = growth_opt >= 2 ? growth_opt : (opt ? 4 : 2);
        for (uint64_t q0 = 0; q0 < pnq; q0 += q_batch) {
            const uint32_t bq = uint32_t(std::min<uint64_t>(q_batch, pnq - q0));
            const float *dq = pq + q0 * dim;
            to_half_slabs_kernel<<<sms * 4, 256, 0, st>>>(dq, bq, dim, q_rows_pad, n_full, tail_k, scale_q, q16, q16t, nullptr, nullptr);
            fill_f32_kernel<<<64, 256, 0, st>>>(thr, bq, INFINITY);
            RG_CUDA_OK(cudaMemsetAsync(cand_count, 0, bq * sizeof(uint32_t), st));
            RG_CUDA_OK(cudaMemsetAsync(overflow, 0, bq * sizeof(uint32_t), st));
            launches += 2;
            GemmParams gp;
            memset(&gp, 0, sizeof(gp));
            gp.n_full = n_full;
            gp.tail_k = tail_k;
            gp.a_bytes = a_bytes;
            gp.a_bufs = a_bufs;
            gp.nq = bq;
            gp.m_tiles = (bq + 2 * kTileM - 1) / (2 * kTileM);
            gp.chunk_tiles = kChunkTiles;
            gp.q_rows_pad = q_rows_pad;
            gp.b_rows_pad = b_rows_pad;
            gp.n_valid = n;
            gp.id_base = id_base;
            gp.l2 = ip ? 0 : 1;
            gp.inv_scale = 1.f / (scale_q * scale_b);
            gp.bnorm = bnorm;
            gp.thr = thr;
            gp.cand = cand;
            gp.cand_count = cand_count;
            gp.n_stages = n_stages;
            // first block (256 rows optimistic, 1024 conservative), then (growth-1) x everything seen so far (tau = +inf in
            // the first block: every score of its rows is kept, so the list can never overflow there)
            uint64_t lo = 0, len = opt ? first_rows : uint64_t(kCap);
            while (lo < b_rows_pad) {
                const uint64_t hi = std::min(b_rows_pad, lo + len);
                gp.row_lo = lo;
                gp.n_tiles = uint32_t((hi - lo) / kPairN);
                const uint32_t units = ((gp.n_tiles + kChunkTiles - 1) / kChunkTiles) * gp.m_tiles;
                const uint32_t grid = 2 * std::min<uint32_t>(units, max_pairs);
                // the threshold after `lo` rows lets ~r / lo of the scores through: dense handling only while that is percents
                const bool dense = lo < dense_rows;
                if (ip) {
                    if (dense) knn_gemm_filter_kernel<false, true><<<grid, kThreads, gemm_smem, st>>>(map_q, map_qt, map_b, map_bt, gp);
                    else knn_gemm_filter_kernel<false, false><<<grid, kThreads, gemm_smem, st>>>(map_q, map_qt, map_b, map_bt, gp);
                } else {
                    if (dense) knn_gemm_filter_kernel<true, true><<<grid, kThreads, gemm_smem, st>>>(map_q, map_qt, map_b, map_bt, gp);
                    else knn_gemm_filter_kernel<true, false><<<grid, kThreads, gemm_smem, st>>>(map_q, map_qt, map_b, map_bt, gp);
                }
                // rank of the new threshold among the scores kept so far (see "Threshold schedule" above); the last
                // block always ends with the conservative k' so that K3 re-scores k' survivors, not c k'
                uint32_t rank = kprime;
                if (opt && hi < b_rows_pad) {
                    const double seen = double(std::min<uint64_t>(hi, n));
                    uint64_t r = r_min;
                    for (; r < kprime; r *= 2) {
                        const double c_eff = double(r) * double(n) / (seen * kprime);
                        const double c_req = r < 8 ? 8.0 : (r < 16 ? 4.0 : (r < 32 ? 2.5 : 2.0));
                        if (c_eff >= c_req) break;
                    }
                    rank = uint32_t(std::min<uint64_t>(kprime, r));
                }
                knn_select_kernel<<<(bq + 3) / 4, 128, 0, st>>>(cand, cand_count, thr, overflow, bq, rank);
                launches += 2;
                lo = hi;
                len = std::max<uint64_t>(len, lo * (growth - 1));
            }
            lap("  batch: gemm + select");
            if (ip)
                knn_rerank_kernel<true><<<(bq + 3) / 4, 128, 4 * kCap * 8, st>>>(d_base, dq, dim, id_base, cand, cand_count, thr, overflow,
                                                                           eps_factor, max_bnorm2, bq, K, n, out.ids + q0 * K,
                                                                           out.dists + q0 * K, need_exact);
            else
                knn_rerank_kernel<false><<<(bq + 3) / 4, 128, 4 * kCap * 8, st>>>(d_base, dq, dim, id_base, cand, cand_count, thr, overflow,
                                                                            eps_factor, max_bnorm2, bq, K, n, out.ids + q0 * K,
                                                                            out.dists + q0 * K, need_exact);
            append_flags_kernel<<<(bq + 255) / 256, 256, 0, st>>>(need_exact, bq, uint32_t(q0), out.flag_list, out.flag_count);
            launches += 2;
            RG_CUDA_OK(cudaGetLastError());
            lap("  batch: rerank");
        }
        return RG_OK;
    };
    auto read_count = [&](const uint32_t *d_count, uint32_t *h) -> rg_status {
        RG_CUDA_OK(cudaMemcpyAsync(h, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        return RG_OK;
    };
    auto exact_scan = [&](const float *pq, const uint32_t *list, uint32_t count, uint32_t *ids, float *dists) {
        const uint32_t grid = std::min<uint32_t>(count, uint32_t(sms) * 2);
        if (ip) knn_exact_scan_kernel<true><<<grid, 256, scan_smem, st>>>(d_base, n, id_base, pq, dim, list, count, K, ids, dists);
        else knn_exact_scan_kernel<false><<<grid, 256, scan_smem, st>>>(d_base, n, id_base, pq, dim, list, count, K, ids, dists);
        launches += 1;
    };

    uint64_t redo_total = 0, scan_total = 0;
    PassOutput main_out = {d_ids, d_dists, flag_list, scal + 3};
    if ((s = run_pass(d_queries, nq, optimistic, main_out)) != RG_OK) return s;
    uint32_t n_flagged = 0;
    if ((s = read_count(scal + 3, &n_flagged)) != RG_OK) return s;  // the only synchronisation after the set-up
    if (n_flagged && !optimistic) {
        scan_total = n_flagged;
        exact_scan(d_queries, flag_list, n_flagged, d_ids, d_dists);
    } else if (n_flagged) {
        // second pass, conservative schedule, over the gathered flagged queries; its own failures take the exact scan
        redo_total = n_flagged;
        float *rq = nullptr, *rd = nullptr;
        uint32_t *ri = nullptr, *rflag = nullptr;
        RG_CUDA_OK(sc.get(12, &rq, uint64_t(n_flagged) * dim));
        RG_CUDA_OK(sc.get(13, &ri, uint64_t(n_flagged) * K));
        RG_CUDA_OK(sc.get(14, &rd, uint64_t(n_flagged) * K));
        RG_CUDA_OK(sc.get(15, &rflag, n_flagged));
        gather_rows_kernel<<<sms * 4, 256, 0, st>>>(d_queries, flag_list, n_flagged, dim, rq);
        PassOutput redo_out = {ri, rd, rflag, scal + 4};
        if ((s = run_pass(rq, n_flagged, false, redo_out)) != RG_OK) return s;
        uint32_t n_scan = 0;
        if ((s = read_count(scal + 4, &n_scan)) != RG_OK) return s;
        if (n_scan) {
            scan_total = n_scan;
            exact_scan(rq, rflag, n_scan, ri, rd);
        }
        scatter_results_kernel<<<sms * 4, 256, 0, st>>>(ri, rd, flag_list, n_flagged, K, d_ids, d_dists);
        launches += 2;
    }
    RG_CUDA_OK(cudaGetLastError());
    RG_CUDA_OK(cudaStreamSynchronize(st));
    lap("redo + exact scan");
    if (stats) {
        stats[0] = launches;
        stats[1] = scan_total;
        stats[2] = redo_total;
    }
    return RG_OK;
}

}  // namespace knn
}  // namespace rg

extern "C" {

rg_status rg_knn_exact_device(const float *d_base, uint64_t n, uint64_t id_base, const float *d_queries, uint64_t nq,
                              uint32_t dim, int metric, uint32_t K, uint32_t *d_ids, float *d_dists, int device,
                              void *cuda_stream) {
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    return rg::knn::knn_device(d_base, n, id_base, d_queries, nq, dim, metric, K, d_ids, d_dists,
                               static_cast<cudaStream_t>(cuda_stream), g_knn_stats, false);
}

rg_status rg_knn_exact(const float *base, uint64_t n, uint64_t id_base, const float *queries, uint64_t nq, uint32_t dim,
                       int metric, uint32_t K, uint32_t *ids, float *dists, int device) {
    if (!base || !queries || !ids || !dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_exact: null argument");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg::knn::Scratch sc;
    float *d_base = nullptr, *d_q = nullptr, *d_d = nullptr;
    uint32_t *d_i = nullptr;
    RG_CUDA_OK(sc.alloc(&d_base, n * dim));
    RG_CUDA_OK(sc.alloc(&d_q, nq * dim));
    RG_CUDA_OK(sc.alloc(&d_i, nq * K));
    RG_CUDA_OK(sc.alloc(&d_d, nq * K));
    RG_CUDA_OK(cudaMemcpy(d_base, base, n * dim * sizeof(float), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMemcpy(d_q, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice));
    rg_status s = rg::knn::knn_device(d_base, n, id_base, d_q, nq, dim, metric, K, d_i, d_d, nullptr, g_knn_stats, false);
    if (s != RG_OK) return s;
    RG_CUDA_OK(cudaMemcpy(ids, d_i, nq * K * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    RG_CUDA_OK(cudaMemcpy(dists, d_d, nq * K * sizeof(float), cudaMemcpyDeviceToHost));
    return RG_OK;
}

rg_status rg_knn_merge_device(const uint32_t *d_part_ids, const float *d_part_dists, uint32_t G, uint64_t nq, uint32_t K,
                              int metric, uint32_t *d_ids, float *d_dists, int device, void *cuda_stream) {
    if (!d_part_ids || !d_part_dists || !d_ids || !d_dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_merge_device: null argument");
    if (G == 0 || K == 0 || uint64_t(G) * K > 1024) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_merge_device: need 1 <= G*K <= 1024");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    if (nq == 0) return RG_OK;
    rg::DeviceGuard guard(device);
    RG_CUDA_OK(cudaFuncSetAttribute(rg::knn::knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 1024 * 8));
    rg::knn::knn_merge_kernel<<<unsigned((nq + 3) / 4), 128, 4 * 1024 * 8, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_part_ids, d_part_dists, G, nq, K, metric != RG_METRIC_L2 ? 1 : 0, d_ids, d_dists);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

rg_status rg_knn_merge(const uint32_t *part_ids, const float *part_dists, uint32_t G, uint64_t nq, uint32_t K, int metric,
                       uint32_t *ids, float *dists, int device) {
    if (!part_ids || !part_dists || !ids || !dists) return rg::fail(RG_ERR_INVALID_ARGUMENT, "rg_knn_merge: null argument");
    if (rg_device_count() <= 0) return rg::fail(RG_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    if (nq == 0) return RG_OK;
    rg::DeviceGuard guard(device);
    if (!guard.ok) return rg::fail(RG_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    rg::knn::Scratch sc;
    uint32_t *d_pi = nullptr, *d_i = nullptr;
    float *d_pd = nullptr, *d_d = nullptr;
    const uint64_t part = uint64_t(G) * nq * K;
    RG_CUDA_OK(sc.alloc(&d_pi, part));
    RG_CUDA_OK(sc.alloc(&d_pd, part));
    RG_CUDA_OK(sc.alloc(&d_i, nq * K));
    RG_CUDA_OK(sc.alloc(&d_d, nq * K));
    RG_CUDA_OK(cudaMemcpy(d_pi, part_ids, part * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMemcpy(d_pd, part_dists, part * sizeof(float), cudaMemcpyHostToDevice));
    rg_status s = rg_knn_merge_device(d_pi, d_pd, G, nq, K, metric, d_i, d_d, device, nullptr);
    if (s != RG_OK) return s;
    RG_CUDA_OK(cudaMemcpy(ids, d_i, nq * K * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    RG_CUDA_OK(cudaMemcpy(dists, d_d, nq * K * sizeof(float), cudaMemcpyDeviceToHost));
    return RG_OK;
}

// launches / queries that needed the exact scan in the last rg_knn_exact* call of the calling thread
void rg_knn_last_stats(uint64_t *launches, uint64_t *exact_scans) {
    if (launches) *launches = g_knn_stats[0];
    if (exact_scans) *exact_scans = g_knn_stats[1];
}
uint64_t rg_knn_last_second_pass_count(void) { return g_knn_stats[2]; }

rg_status rg_knn_release_scratch(int device) {
    if (device >= 0) {
        rg::DeviceGuard guard(device);
        rg::knn::release_scratch(device);
        return RG_OK;
    }
    const int n = rg_device_count();
    for (int d = 0; d < n; ++d) {
        rg::DeviceGuard guard(d);
        rg::knn::release_scratch(d);
    }
    return RG_OK;
}
}
