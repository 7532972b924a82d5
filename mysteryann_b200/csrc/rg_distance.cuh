// FP32 distances in the exact operation order of the compiled reference (include/efanna2e/distance.h:39-89,
// 179-223 as g++ -Ofast emits them: 16 lane accumulators, unfused vmulps+vaddps main loop, fused 8-wide tail,
// folds 16->8->4, (x0+x1)+(x2+x3)).  Four CUDA lanes score one row; lane t owns AVX lanes 4t..4t+3.
// Used by K1 (rg_search.cu) and the occlusion prunes of the graph build (rg_build.cu).
#pragma once
#include "rg_common.cuh"

namespace rg {

// ---- distance of 8 rows per warp, 4 lanes per row, reference operation order ---------------------
// Lane t (0..3) of a group owns AVX lanes 4t..4t+3 of the reference's 16-lane accumulator.
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

// ---- main loop on packed FP32 (sm_100 FFMA2 / FADD2: two IEEE-rounded operations per instruction) --------------------
// The reference's main loop rounds the product and the sum separately (vmulps, vaddps), and ptxas contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (it does not contract the scalar .rn forms), so the product is written as
// fma(a, b, -0.0): RN(a*b + (-0)) == RN(a*b) for every input including signed zeros, and the -0.0 pair `nz` comes from the
// kernel parameters, where the compiler cannot see its value and therefore cannot fold the fma away.  Each element sees
// exactly the scalar sequence acc = acc + RN(v*q) (L2: d = v - q; acc = acc + RN(d*d)), so the result is bit-identical
// to the reference's; against scalar FMUL + FADD the loop issues half the FP instructions (measured: +1 % on K1 at
// D = 200, +15 % at D = 512, profiles/r02_k1_sweep_packed_fp32.txt, r02_bench_c3_k10.txt).
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
constexpr uint64_t kNegZero2 = 0x8000000080000000ull;  // host side: SearchParams::neg_zero2, PruneParams::neg_zero2

template <bool kIP>
__device__ __forceinline__ float lane_exact_distance_x2(const float4 *__restrict__ rp, const float4 *__restrict__ qp,
                                                         uint32_t n16, bool tail8, uint32_t t, uint64_t nz) {
    uint64_t a01 = 0, a23 = 0;  // (+0, +0)
    // one LDS.128 per operand lands in two aligned register pairs: (x, y) and (z, w) as they are needed
    const ulonglong2 *rp2 = reinterpret_cast<const ulonglong2 *>(rp), *qp2 = reinterpret_cast<const ulonglong2 *>(qp);
#pragma unroll 4
    for (uint32_t s = 0; s < n16; ++s) {
        const ulonglong2 v = rp2[4 * s], q = qp2[4 * s];
        if (kIP) {
            a01 = add2(a01, fma2(v.x, q.x, nz));
            a23 = add2(a23, fma2(v.y, q.y, nz));
        } else {
            const uint64_t d01 = sub2(v.x, q.x), d23 = sub2(v.y, q.y);
            a01 = add2(a01, fma2(d01, d01, nz));
            a23 = add2(a23, fma2(d23, d23, nz));
        }
    }
    float4 acc;
    unpack2(a01, acc.x, acc.y);
    unpack2(a23, acc.z, acc.w);
    float4 vt = make_float4(0.f, 0.f, 0.f, 0.f), qt = vt;
    if (tail8 && t < 2) {
        vt = rp[4 * n16];
        qt = qp[4 * n16];
    }
    return finish_distance<kIP>(acc, tail8, vt, qt, t);
}

}  // namespace rg
